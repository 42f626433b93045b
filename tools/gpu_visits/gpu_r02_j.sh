#!/bin/bash
# round 2, visit J: occupancy of the thread-per-instance arm kernel
tag=r02j
mkdir -p gpurun_out
for m in 3 4 5 6; do
  for w in franka_65536 tiago_65536; do
    IWBC_ARM_TPI=1 IWBC_TPI_MINB=$m timeout 300 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu --no-secondary 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w tpi minb=$m', round(d['value']), 'e2e', round(d['e2e']['value']), 'opt', d['config']['optimal_fraction'])"
  done
done
IWBC_ARM_TPI=1 IWBC_TPI_MINB=4 timeout 300 python bench.py --workload mixed_icub_tiago_16384 --steps 20 --warmup 3 --no-cpu --no-secondary 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('mixed tpi', round(d['value']), 'e2e', round(d['e2e']['value']))"
IWBC_ARM_TPI=1 IWBC_TPI_MINB=4 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_behaviors.py tests/test_collision.py tests/test_gpu_handles.py -m gpu -q 2>&1 | tail -4
