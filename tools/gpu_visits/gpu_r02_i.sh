#!/bin/bash
# round 2, visit I: thread-per-instance experiment for the arms (generic code with a group of one thread, local-memory working set)
tag=r02i
mkdir -p gpurun_out
IWBC_ARM_TPI=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "franka_batch or tiago or franka_full or franka_known" 2>&1 | tail -4
for t in 0 1; do
  for w in franka_65536 tiago_65536; do
    IWBC_ARM_TPI=$t timeout 300 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu --no-secondary 2>gpurun_out/${tag}_$w_$t.err | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w tpi=$t', round(d['value']), 'e2e', round(d['e2e']['value']), 'opt', d['config']['optimal_fraction'])"
  done
done
IWBC_ARM_TPI=1 timeout 600 ncu --set full --clock-control none -k regex:'k_cycle_tpi' --launch-skip 3 -c 1 -f -o /tmp/tpi_tiago python bench.py --workload tiago_65536 --steps 2 --warmup 3 --no-cpu --no-secondary > /dev/null 2>&1
ncu -i /tmp/tpi_tiago.ncu-rep --page raw --csv 2>/dev/null | python -c "
import sys,csv
rows=list(csv.reader(sys.stdin)); h=rows[0]
keys=['gpu__time_duration.sum','launch__registers_per_thread','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','smsp__thread_inst_executed_per_inst_executed.ratio','dram__bytes_read.sum','dram__bytes_write.sum','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__inst_executed.sum']
for r in rows[2:]:
    d=dict(zip(h,r))
    for k in keys: print(k, d.get(k))
" > gpurun_out/${tag}_ncu_tpi_tiago.txt 2>&1; cat gpurun_out/${tag}_ncu_tpi_tiago.txt
