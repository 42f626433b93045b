#!/bin/bash
# round 2, visit P: final state of this round's kernels -- full GPU suite, default bench line, ncu evidence for profiles/
tag=r02p
mkdir -p gpurun_out /tmp/ncu
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -5
timeout 900 python bench.py > gpurun_out/${tag}_bench_default.json 2> gpurun_out/${tag}_bench_default.err; tail -c 600 gpurun_out/${tag}_bench_default.json
timeout 300 python tools/phase_profile.py talos_ds_16384 16384 > gpurun_out/${tag}_phase_talos_ds.json 2>/dev/null
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_talos_ds_16384.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-secondary > gpurun_out/${tag}_ncu_launch_talos.log 2>&1
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 60 --csv --log-file gpurun_out/${tag}_launches_talos_walk_131072.csv python tools/walk_step_only.py > gpurun_out/${tag}_ncu_launch_walk.log 2>&1
full() { # name kernel-regex skip count command...
  name=$1; regex=$2; skip=$3; count=$4; shift 4
  timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"$regex" --launch-skip $skip -c $count -f -o /tmp/ncu/${tag}_$name "$@" > gpurun_out/${tag}_ncu_full_$name.log 2>&1
  python tools/ncu_summary.py /tmp/ncu/${tag}_$name.ncu-rep > gpurun_out/${tag}_ncu_$name.txt 2> gpurun_out/${tag}_ncu_$name.err
  head -45 gpurun_out/${tag}_ncu_$name.txt
}
full talos_ds 'k_fast|k_slow' 6 2 python bench.py --steps 2 --warmup 3 --no-cpu --no-secondary
full icub 'k_fast|k_slow' 6 2 python bench.py --workload icub_ds_16384 --steps 2 --warmup 3 --no-cpu --no-secondary
