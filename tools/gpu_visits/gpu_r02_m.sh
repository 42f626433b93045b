#!/bin/bash
# round 2, visit M: failing roll-out test in detail, line-level ncu of k_fast, resident-CTA sweep of the thread-per-instance kernel
tag=r02m
mkdir -p gpurun_out /tmp/ncu
timeout 900 python -m pytest tests/test_determinism_gpu.py -m gpu -x -q -s 2>&1 | tail -40 > gpurun_out/${tag}_determinism.txt; tail -30 gpurun_out/${tag}_determinism.txt
line() { python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', d['ms_per_step'])"; }
for w in tiago_65536 franka_65536; do
for c in 1 2 3 4; do
for B in 65536 262144; do
  IWBC_TPI_CTAS=$c timeout 300 python bench.py --workload $w --batch $B --steps 10 --warmup 3 --no-cpu --no-secondary 2>/dev/null | line "$w ctas=$c B=$B"
done; done; done | tee gpurun_out/${tag}_tpi_ctas.txt
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"k_fast" --launch-skip 3 -c 1 -f -o /tmp/ncu/${tag}_talos_ds python bench.py --steps 2 --warmup 3 --no-cpu --no-secondary > gpurun_out/${tag}_ncu_full.log 2>&1
python tools/ncu_summary.py /tmp/ncu/${tag}_talos_ds.ncu-rep > gpurun_out/${tag}_ncu_talos_ds.txt 2> gpurun_out/${tag}_ncu_talos_ds.err
tail -50 gpurun_out/${tag}_ncu_talos_ds.txt
