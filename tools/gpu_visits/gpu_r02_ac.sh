#!/bin/bash
# round 2, visit AC: A/B of the one-step-ahead pivot norm in the QR, arm chunk plan, GPU suite
tag=r02ac
mkdir -p gpurun_out
line() { python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', d['ms_per_step'], 'e2e_ms', round(d['e2e']['ms_per_step'],4))"; }
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_behaviors.py tests/test_determinism_gpu.py -m gpu -q 2>&1 | tail -4 | tee gpurun_out/${tag}_pytest.txt
for v in default sg0 default sg0; do
  if [ $v = default ]; then unset IWBC_B200_LIB; else export IWBC_B200_LIB=$PWD/build/variants/libiwbc_$v.so; fi
  for w in talos_ds_16384 icub_ds_16384 talos_walk_131072; do
    timeout 600 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu --no-secondary 2>/dev/null | line "$v $w"
  done
done | tee gpurun_out/${tag}_variants.txt
unset IWBC_B200_LIB
for w in franka_65536 tiago_65536; do timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu --no-secondary 2>/dev/null | line "$w"; done
