#!/bin/bash
# round 2, visit K: thread-per-instance arm kernel as the default for large batches -- tests, batch-size crossover against the
# warp-per-instance kernel, ncu evidence for both arms
tag=r02k
mkdir -p gpurun_out /tmp/ncu
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
line() { python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', d['ms_per_step'], 'opt', d['config'].get('optimal_fraction'))"; }
for w in franka_65536 tiago_65536; do
  for B in 1024 2048 4096 8192 16384 32768 65536 262144; do
    for m in 0 1000000000; do
      IWBC_ARM_TPI_MIN_BATCH=$m timeout 300 python bench.py --workload $w --batch $B --steps 20 --warmup 3 --no-cpu --no-secondary 2>/dev/null | line "$w B=$B tpi_min=$m"
    done
  done
done | tee gpurun_out/${tag}_crossover.txt
for w in franka_65536 tiago_65536 mixed_icub_tiago_16384; do
  timeout 300 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu --no-secondary 2>gpurun_out/${tag}_$w.err > gpurun_out/${tag}_bench_$w.json; line $w < gpurun_out/${tag}_bench_$w.json
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_$w.csv python bench.py --workload $w --steps 2 --warmup 3 --no-cpu --no-secondary > gpurun_out/${tag}_ncu_launch_$w.log 2>&1
done
full() { # name kernel-regex skip count command...
  name=$1; regex=$2; skip=$3; count=$4; shift 4
  timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"$regex" --launch-skip $skip -c $count -f -o /tmp/ncu/${tag}_$name "$@" > gpurun_out/${tag}_ncu_full_$name.log 2>&1
  python tools/ncu_summary.py /tmp/ncu/${tag}_$name.ncu-rep > gpurun_out/${tag}_ncu_$name.txt 2> gpurun_out/${tag}_ncu_$name.err
  head -30 gpurun_out/${tag}_ncu_$name.txt
}
full franka_tpi 'k_cycle_tpi' 3 1 python bench.py --workload franka_65536 --steps 2 --warmup 3 --no-cpu --no-secondary
full tiago_tpi 'k_cycle_tpi' 3 1 python bench.py --workload tiago_65536 --steps 2 --warmup 3 --no-cpu --no-secondary
timeout 600 python bench.py > gpurun_out/${tag}_bench_default.json 2> gpurun_out/${tag}_bench_default.err; tail -c 1500 gpurun_out/${tag}_bench_default.json
