#!/bin/bash
# round 2, final visit (r02e): what the driver runs at round end -- smoke(), the default bench line, the reference arm -- plus the final ncu evidence
tag=r02f
mkdir -p gpurun_out /tmp/ncu
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/${tag}_bench_default.json 2> gpurun_out/${tag}_bench_default.err; tail -c 300 gpurun_out/${tag}_bench_default.json; echo
timeout 900 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err; tail -c 700 gpurun_out/${tag}_bench_reference.json; echo
timeout 300 python tools/phase_profile.py talos_ds_16384 16384 > gpurun_out/${tag}_phase_talos_ds.json 2>/dev/null
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_talos_ds_16384.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-secondary > gpurun_out/${tag}_ncu_launch_talos.log 2>&1
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 60 --csv --log-file gpurun_out/${tag}_launches_talos_walk_131072.csv python tools/walk_step_only.py > gpurun_out/${tag}_ncu_launch_walk.log 2>&1
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"k_fast|k_slow" --launch-skip 6 -c 2 -f -o /tmp/ncu/${tag}_talos_ds python bench.py --steps 2 --warmup 3 --no-cpu --no-secondary > gpurun_out/${tag}_ncu_full.log 2>&1
python tools/ncu_summary.py /tmp/ncu/${tag}_talos_ds.ncu-rep > gpurun_out/${tag}_ncu_talos_ds.txt 2> gpurun_out/${tag}_ncu_talos_ds.err
head -24 gpurun_out/${tag}_ncu_talos_ds.txt
