#!/bin/bash
# round 2, visit Y: A/B of flattened block records + L2 prefetch + hoisted tile indices against the previous commit's build
tag=r02y
mkdir -p gpurun_out
line() { python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', d['ms_per_step'], 'e2e_ms', round(d['e2e']['ms_per_step'],4))"; }
for v in default head pf0 default head pf0; do
  if [ $v = default ]; then unset IWBC_B200_LIB; else export IWBC_B200_LIB=$PWD/build/variants/libiwbc_$v.so; fi
  for w in talos_ds_16384 icub_ds_16384 talos_walk_131072; do
    timeout 600 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu --no-secondary 2>/dev/null | line "$v $w"
  done
done | tee gpurun_out/${tag}_variants.txt
unset IWBC_B200_LIB
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/${tag}_pytest.txt
