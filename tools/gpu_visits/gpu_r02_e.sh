#!/bin/bash
# round 2, visit E: occupancy experiment for the arm kernel (min CTAs per SM), phase profiles of the arms, new bench line, multi-shard call
tag=r02e
mkdir -p gpurun_out
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
for m in 3 4 5 6; do
  bench franka_minb$m IWBC_ARM_MINB=$m franka_65536
  bench tiago_minb$m IWBC_ARM_MINB=$m tiago_65536
done
for w in franka_65536 tiago_65536; do
  timeout 300 python tools/phase_profile.py $w 65536 > gpurun_out/${tag}_phase_$w.json 2> gpurun_out/${tag}_phase_$w.err; cat gpurun_out/${tag}_phase_$w.json | cut -c1-600
done
bench talos_ds X=1 talos_ds_16384
bench talos_walk X=1 talos_walk_131072
bench mixed X=1 mixed_icub_tiago_16384
( time timeout 1200 python -m pytest tests -m gpu -q --durations=6 ) > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/${tag}_pytest_gpu.log
( time timeout 900 python bench.py --steps 20 --warmup 3 ) > gpurun_out/${tag}_bench_default.json 2> gpurun_out/${tag}_bench_default.err
tail -c 1500 gpurun_out/${tag}_bench_default.json; tail -5 gpurun_out/${tag}_bench_default.err
