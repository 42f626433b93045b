#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench lines for every workload, the reference arm, ncu launch lists + full captures of the hot kernels.
# usage (through gpurun): bash tools/gpu_round.sh <tag>
tag=${1:-r01}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/${tag}_smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${tag}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/${tag}_smoke.log
timeout 600 python bench.py > gpurun_out/${tag}_bench_talos_ds.json 2> gpurun_out/${tag}_bench_talos_ds.err
tail -c 400 gpurun_out/${tag}_bench_talos_ds.json
for w in talos_fast_ds_16384 franka_65536 talos_walk_131072 talos_walk_standin_131072 icub_ds_16384 tiago_65536; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu > gpurun_out/${tag}_bench_$w.json 2> gpurun_out/${tag}_bench_$w.err
done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err
for w in talos_ds_16384 talos_walk_131072 icub_ds_16384; do
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_$w.csv python bench.py --workload $w --steps 2 --warmup 3 --no-cpu > gpurun_out/${tag}_ncu_launch_$w.log 2>&1
done
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'k_fast|k_slow' --launch-skip 6 -c 2 -f -o gpurun_out/${tag}_full python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/${tag}_ncu_full.log 2>&1
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'k_slow' --launch-skip 6 -c 1 -f -o gpurun_out/${tag}_full_icub python bench.py --workload icub_ds_16384 --steps 1 --warmup 3 --no-cpu > gpurun_out/${tag}_ncu_full_icub.log 2>&1
for w in talos_walk_131072 icub_ds_16384; do timeout 300 python tools/phase_profile.py $w 16384 > gpurun_out/${tag}_phase_$w.json 2> gpurun_out/${tag}_phase_$w.err; done
timeout 300 python tools/phase_profile.py talos_ds_16384 16384 > gpurun_out/${tag}_phase_talos_ds_16384.json 2> gpurun_out/${tag}_phase_talos_ds.err
ls gpurun_out | grep ${tag} | wc -l
