#!/bin/bash
# round 2, visit G: ncu evidence for profiles/ (launch lists, full captures summarised ON THE BOX: the reports are too large to travel)
tag=r02g
mkdir -p gpurun_out /tmp/ncu
for w in talos_ds_16384 franka_65536 tiago_65536 icub_ds_16384 mixed_icub_tiago_16384; do
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_$w.csv python bench.py --workload $w --steps 2 --warmup 3 --no-cpu --no-secondary > gpurun_out/${tag}_ncu_launch_$w.log 2>&1
done
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 60 --csv --log-file gpurun_out/${tag}_launches_talos_walk_131072.csv python tools/walk_step_only.py > gpurun_out/${tag}_ncu_launch_walk.log 2>&1
full() { # name kernel-regex skip count command...
  name=$1; regex=$2; skip=$3; count=$4; shift 4
  timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"$regex" --launch-skip $skip -c $count -f -o /tmp/ncu/${tag}_$name "$@" > gpurun_out/${tag}_ncu_full_$name.log 2>&1
  python tools/ncu_summary.py /tmp/ncu/${tag}_$name.ncu-rep > gpurun_out/${tag}_ncu_$name.txt 2> gpurun_out/${tag}_ncu_$name.err
  head -25 gpurun_out/${tag}_ncu_$name.txt
}
full talos_ds 'k_fast|k_slow' 6 2 python bench.py --steps 2 --warmup 3 --no-cpu --no-secondary
full franka 'k_cycle' 3 1 python bench.py --workload franka_65536 --steps 2 --warmup 3 --no-cpu --no-secondary
full tiago 'k_cycle' 3 1 python bench.py --workload tiago_65536 --steps 2 --warmup 3 --no-cpu --no-secondary
full icub 'k_slow' 3 1 python bench.py --workload icub_ds_16384 --steps 2 --warmup 3 --no-cpu --no-secondary
timeout 1500 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"k_fast|k_slow" -c 6 -f -o /tmp/ncu/${tag}_walk python tools/walk_step_only.py > gpurun_out/${tag}_ncu_full_walk.log 2>&1
python tools/ncu_summary.py /tmp/ncu/${tag}_walk.ncu-rep > gpurun_out/${tag}_ncu_walk.txt 2> gpurun_out/${tag}_ncu_walk.err
timeout 300 python tools/phase_profile.py franka_65536 65536 > gpurun_out/${tag}_phase_franka.json 2>/dev/null
timeout 300 python tools/phase_profile.py tiago_65536 65536 > gpurun_out/${tag}_phase_tiago.json 2>/dev/null
timeout 300 python tools/phase_profile.py talos_ds_16384 16384 > gpurun_out/${tag}_phase_talos_ds.json 2>/dev/null
timeout 300 python tools/latency_sweep.py talos_ds_16384 > gpurun_out/${tag}_latency_talos_ds.json 2>/dev/null
timeout 300 python tools/latency_sweep.py franka_65536 > gpurun_out/${tag}_latency_franka.json 2>/dev/null
du -sh gpurun_out
