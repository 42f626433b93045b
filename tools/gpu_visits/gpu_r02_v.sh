#!/bin/bash
# round 2, visit V: iCub walking roll-out sensitivity, A/B of the single-thread diagonal factor
tag=r02v
mkdir -p gpurun_out
timeout 600 python tools/diag_icub_walk.py 2>&1 | tail -25 | tee gpurun_out/${tag}_icub_walk_growth.txt
IWBC_B200_LIB=$PWD/build/variants/libiwbc_bf0.so timeout 600 python tools/diag_icub_walk.py 2>&1 | tail -25 | tee gpurun_out/${tag}_icub_walk_growth_bf0.txt
line() { python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', d['ms_per_step'], 'e2e_ms', round(d['e2e']['ms_per_step'],4))"; }
for v in default diag0 default diag0; do
  if [ $v = default ]; then unset IWBC_B200_LIB; else export IWBC_B200_LIB=$PWD/build/variants/libiwbc_$v.so; fi
  for w in talos_ds_16384 icub_ds_16384 talos_walk_131072; do
    timeout 600 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu --no-secondary 2>/dev/null | line "$v $w"
  done
done | tee gpurun_out/${tag}_variants.txt
unset IWBC_B200_LIB
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_determinism_gpu.py tests/test_torque_task.py -m gpu -q 2>&1 | tail -4
