#!/bin/bash
# round 2, visit O: A/B of the compile-time variants of k_fast (CTAs per SM x Cholesky form), full GPU suite on the default build
tag=r02o
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15
line() { python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', d['ms_per_step'], 'frac', round(d['roofline']['frac'],4))"; }
for v in default m3rt1 m3rt0 m4rt0; do
  if [ $v = default ]; then unset IWBC_B200_LIB; else export IWBC_B200_LIB=$PWD/build/variants/libiwbc_$v.so; fi
  for w in talos_ds_16384 icub_ds_16384 talos_walk_131072; do
    timeout 600 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu --no-secondary 2>/dev/null | line "$v $w"
  done
done | tee gpurun_out/${tag}_variants.txt
