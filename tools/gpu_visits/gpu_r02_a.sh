#!/bin/bash
# round 2, visit A: parity tests (incl. the reference-scale determinism re-host), measured worst-case parity table, baseline bench lines
tag=r02a
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/${tag}_smi.txt 2>&1
nproc >> gpurun_out/${tag}_smi.txt
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=12 ) > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${tag}_pytest_gpu.log
tail -30 gpurun_out/${tag}_pytest_gpu.log
timeout 900 python tools/parity_report.py > gpurun_out/${tag}_parity_report.jsonl 2> gpurun_out/${tag}_parity_report.err
for w in talos_ds_16384 franka_65536 tiago_65536 icub_ds_16384 talos_walk_131072; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu > gpurun_out/${tag}_bench_$w.json 2> gpurun_out/${tag}_bench_$w.err
  tail -c 300 gpurun_out/${tag}_bench_$w.json
done
