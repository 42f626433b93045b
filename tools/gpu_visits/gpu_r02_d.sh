#!/bin/bash
# round 2, visit D: fixes after visit C; arm-shape experiment (sub-warp groups for Franka / Tiago)
tag=r02d
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_stabilizer.py tests/test_determinism_gpu.py -m gpu -q -s ) > gpurun_out/${tag}_pytest_new.log 2>&1; echo "pytest(new) rc=$?"
grep -E "stabilizer|determinism|passed|failed|Error" gpurun_out/${tag}_pytest_new.log | head -20
bench() { # name env workload
  timeout 600 env $2 python bench.py --workload $3 --steps 20 --warmup 3 --no-cpu --no-secondary > gpurun_out/${tag}_bench_$1.json 2> gpurun_out/${tag}_bench_$1.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${tag}_bench_$1.json").read().strip().splitlines()[-1])
    print("$1", round(d["value"]), "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"],4), "launches", d["gpu_launches"])
except Exception as e:
    print("$1 failed", e)
PY
}
bench talos_ds X=1 talos_ds_16384
bench talos_walk X=1 talos_walk_131072
bench icub X=1 icub_ds_16384
for shape in 32,4 16,8 16,12 8,16 8,24; do
  bench franka_$shape IWBC_ARM_SHAPE=$shape franka_65536
  bench tiago_$shape IWBC_ARM_SHAPE=$shape tiago_65536
done
IWBC_ARM_SHAPE=8,24 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "franka_batch or tiago or dense_qp" 2>&1 | tail -3
IWBC_ARM_SHAPE=16,8 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "franka_batch or tiago or dense_qp" 2>&1 | tail -3
