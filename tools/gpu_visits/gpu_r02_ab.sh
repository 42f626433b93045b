#!/bin/bash
# round 2, visit AB: chunk plan of the host-buffer call for the thread-per-instance arm kernel
tag=r02ab
mkdir -p gpurun_out
line() { python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['value']), 'e2e', round(d['e2e']['value']), 'e2e_ms', round(d['e2e']['ms_per_step'],4), 'p50', round(d['e2e']['p50_cycle_latency_ms'],3))"; }
for w in franka_65536 tiago_65536; do
for p in 4,10 3,3 3,4 3,6 4,4 4,6 3,5; do
  IWBC_PIPE=$p timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu --no-secondary 2>/dev/null | line "pipe=$p $w"
done; done | tee gpurun_out/${tag}_pipe_arms.txt
