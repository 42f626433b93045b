#!/bin/bash
# Second kind of visit: the active-set-heavy workloads (Talos walking, iCub) -- launch lists, phase profiles, one full capture of k_slow.
tag=${1:-r01b}
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/${tag}_smoke.log
timeout 600 python bench.py --steps 50 --warmup 3 > gpurun_out/${tag}_bench_talos_ds.json 2> gpurun_out/${tag}_bench_talos_ds.err
timeout 600 python bench.py --workload tiago_65536 --steps 10 --warmup 3 --no-cpu > gpurun_out/${tag}_bench_tiago.json 2> gpurun_out/${tag}_bench_tiago.err
for w in talos_walk_131072 icub_ds_16384; do
  timeout 300 python tools/phase_profile.py $w 16384 > gpurun_out/${tag}_phase_$w.json 2> gpurun_out/${tag}_phase_$w.err
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${tag}_launches_$w.csv python bench.py --workload $w --steps 2 --warmup 3 --no-cpu > gpurun_out/${tag}_ncu_launch_$w.log 2>&1
done
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'k_slow' --launch-skip 3 -c 2 -f -o gpurun_out/${tag}_full_walk python bench.py --workload talos_walk_131072 --batch 32768 --steps 1 --warmup 3 --no-cpu > gpurun_out/${tag}_ncu_full_walk.log 2>&1
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'k_cycle' --launch-skip 3 -c 1 -f -o gpurun_out/${tag}_full_franka python bench.py --workload franka_65536 --steps 1 --warmup 3 --no-cpu > gpurun_out/${tag}_ncu_full_franka.log 2>&1
ls -la gpurun_out | tail -20
