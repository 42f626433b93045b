#!/bin/bash
# round 2, visit T: A/B of the host-folded constants in k_fast, chunk plan of the host-buffer call after the 4-CTA change
tag=r02t
mkdir -p gpurun_out
line() { python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', d['ms_per_step'], 'e2e_ms', round(d['e2e']['ms_per_step'],4))"; }
for v in default bf0 default bf0; do
  if [ $v = default ]; then unset IWBC_B200_LIB; else export IWBC_B200_LIB=$PWD/build/variants/libiwbc_$v.so; fi
  for w in talos_ds_16384 icub_ds_16384 talos_walk_131072; do
    timeout 600 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu --no-secondary 2>/dev/null | line "$v $w"
  done
done | tee gpurun_out/${tag}_variants.txt
unset IWBC_B200_LIB
for p in 4,10 2,10 2,6 3,8 1,6 2,4 6,12; do
  IWBC_PIPE=$p timeout 600 python bench.py --workload talos_ds_16384 --steps 20 --warmup 3 --no-cpu --no-secondary 2>/dev/null | line "pipe=$p talos_ds"
done | tee gpurun_out/${tag}_pipe.txt
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_behaviors.py -m gpu -q 2>&1 | tail -3
