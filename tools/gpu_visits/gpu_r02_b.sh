#!/bin/bash
# round 2, visit B: per-handle device state (no __constant__ ownership) -> parity suite, concurrency, bench lines against visit A
tag=r02b
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --durations=8 ) > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${tag}_pytest_gpu.log
tail -25 gpurun_out/${tag}_pytest_gpu.log
for w in talos_ds_16384 franka_65536 tiago_65536 icub_ds_16384 talos_walk_131072; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu --no-secondary > gpurun_out/${tag}_bench_$w.json 2> gpurun_out/${tag}_bench_$w.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${tag}_bench_$w.json").read().strip().splitlines()[-1])
    print("$w", round(d["value"]), "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"],4), "launches", d["gpu_launches"])
except Exception as e:
    print("$w failed", e)
PY
done
