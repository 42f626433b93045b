#!/bin/bash
# round 2, visit F: arm kernel at 4 CTAs/SM (inlined phases, overlaid scratch), parity of the generic path, ncu evidence for profiles/
tag=r02f
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
bench franka X=1 franka_65536
bench tiago X=1 tiago_65536
bench franka_minb3 IWBC_ARM_MINB=3 franka_65536
bench tiago_minb3 IWBC_ARM_MINB=3 tiago_65536
bench mixed X=1 mixed_icub_tiago_16384
bench talos_ds X=1 talos_ds_16384
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_handles.py tests/test_collision.py tests/test_cpp_shim.py -m gpu -q -k "not determinism" 2>&1 | tail -5
IWBC_NO_FAST=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "config3 or moving_base or icub" 2>&1 | tail -3
for w in franka_65536 tiago_65536; do
  timeout 300 python tools/phase_profile.py $w 65536 > gpurun_out/${tag}_phase_$w.json 2> gpurun_out/${tag}_phase_$w.err; cut -c1-500 gpurun_out/${tag}_phase_$w.json
done
# launch lists (shares of a step)
for w in talos_ds_16384 franka_65536 tiago_65536 icub_ds_16384 mixed_icub_tiago_16384; do
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_$w.csv python bench.py --workload $w --steps 2 --warmup 3 --no-cpu --no-secondary > gpurun_out/${tag}_ncu_launch_$w.log 2>&1
done
# walking: skip the seed walk's launches (k_behavior + 2 kernels per tick, ~1100 ticks) and list the 131 072-state steps
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_fast|k_slow|k_behavior' --profile-from-start off -c 60 --csv --log-file gpurun_out/${tag}_launches_talos_walk_131072.csv python tools/walk_step_only.py > gpurun_out/${tag}_ncu_launch_walk.log 2>&1
# full captures of the dominant kernels
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'k_fast|k_slow' --launch-skip 6 -c 2 -f -o gpurun_out/${tag}_full_talos python bench.py --steps 2 --warmup 3 --no-cpu --no-secondary > gpurun_out/${tag}_ncu_full_talos.log 2>&1
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'k_cycle' --launch-skip 3 -c 1 -f -o gpurun_out/${tag}_full_franka python bench.py --workload franka_65536 --steps 2 --warmup 3 --no-cpu --no-secondary > gpurun_out/${tag}_ncu_full_franka.log 2>&1
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'k_cycle' --launch-skip 3 -c 1 -f -o gpurun_out/${tag}_full_tiago python bench.py --workload tiago_65536 --steps 2 --warmup 3 --no-cpu --no-secondary > gpurun_out/${tag}_ncu_full_tiago.log 2>&1
timeout 1500 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"k_fast|k_slow" -c 6 -f -o gpurun_out/${tag}_full_walk python tools/walk_step_only.py > gpurun_out/${tag}_ncu_full_walk.log 2>&1
ls -la gpurun_out | grep ${tag} | wc -l
