#!/bin/bash
# round 2, visit AQ: chunk plan of the host-buffer call for the continuation-heavy workloads (iCub, walking)
tag=r02aq
mkdir -p gpurun_out
line() { python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['value']), 'e2e', round(d['e2e']['value']), 'e2e_ms', round(d['e2e']['ms_per_step'],4))"; }
for w in icub_ds_16384 talos_walk_131072; do
for p in 2,6 3,8 4,10 2,10 4,16 6,12; do
  IWBC_PIPE=$p timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu --no-secondary 2>/dev/null | line "pipe=$p $w"
done; done | tee gpurun_out/${tag}_pipe_slowheavy.txt
