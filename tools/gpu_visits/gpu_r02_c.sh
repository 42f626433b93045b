#!/bin/bash
# round 2, visit C: stabilizer pre-step, weight-vector variants in the fast kernel; squat divergence diagnostic
tag=r02c
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_stabilizer.py tests/test_gpu_parity.py tests/test_gpu_handles.py tests/test_safety.py -m gpu -q -s --durations=5 ) > gpurun_out/${tag}_pytest_new.log 2>&1; echo "pytest(new) rc=$?"
tail -30 gpurun_out/${tag}_pytest_new.log
timeout 600 python tools/diag_squat.py > gpurun_out/${tag}_diag_squat.log 2>&1; tail -15 gpurun_out/${tag}_diag_squat.log
( time timeout 1200 python -m pytest tests -m gpu -q --durations=6 ) > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/${tag}_pytest_gpu.log
for w in talos_ds_16384 icub_ds_16384 talos_walk_131072; do
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
