#!/bin/bash
# round 2, visit R: A/B of k_fast micro-optimisations (QR steps unrolled, index tables, 16-byte loads)
tag=r02r
mkdir -p gpurun_out
line() { python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', d['ms_per_step'], 'frac', round(d['roofline']['frac'],4))"; }
for v in default qr0 base qr1only default; do
  if [ $v = default ]; then unset IWBC_B200_LIB; else export IWBC_B200_LIB=$PWD/build/variants/libiwbc_$v.so; fi
  for w in talos_ds_16384 icub_ds_16384 talos_walk_131072; do
    timeout 600 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu --no-secondary 2>/dev/null | line "$v $w"
  done
done | tee gpurun_out/${tag}_variants.txt
unset IWBC_B200_LIB
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_behaviors.py tests/test_determinism_gpu.py -m gpu -q 2>&1 | tail -5
