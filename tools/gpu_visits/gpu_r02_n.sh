#!/bin/bash
# round 2, visit N: register-tile Cholesky in k_fast, one pass per launch in the thread-per-instance kernel
tag=r02n
mkdir -p gpurun_out /tmp/ncu
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
line() { python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', d['ms_per_step'], 'opt', d['config'].get('optimal_fraction'), 'frac', d['roofline']['frac'])"; }
for w in talos_ds_16384 icub_ds_16384 talos_walk_131072 talos_fast_ds_16384; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu --no-secondary 2>gpurun_out/${tag}_$w.err > gpurun_out/${tag}_bench_$w.json; line $w < gpurun_out/${tag}_bench_$w.json
done
for B in 65536 131072 262144; do
  timeout 300 python bench.py --workload tiago_65536 --batch $B --steps 20 --warmup 3 --no-cpu --no-secondary 2>/dev/null | line "tiago B=$B"
done
timeout 300 python tools/phase_profile.py talos_ds_16384 16384 > gpurun_out/${tag}_phase_talos_ds.json 2>/dev/null
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"k_fast" --launch-skip 3 -c 1 -f -o /tmp/ncu/${tag}_talos_ds python bench.py --steps 2 --warmup 3 --no-cpu --no-secondary > gpurun_out/${tag}_ncu_full.log 2>&1
python tools/ncu_summary.py /tmp/ncu/${tag}_talos_ds.ncu-rep > gpurun_out/${tag}_ncu_talos_ds.txt 2> gpurun_out/${tag}_ncu_talos_ds.err
head -30 gpurun_out/${tag}_ncu_talos_ds.txt
