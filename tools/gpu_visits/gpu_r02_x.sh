#!/bin/bash
# round 2, visit X: state after the k_fast trims -- GPU suite, default bench line, ncu evidence
tag=r02x
mkdir -p gpurun_out /tmp/ncu
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/${tag}_pytest.txt
timeout 900 python bench.py > gpurun_out/${tag}_bench_default.json 2> gpurun_out/${tag}_bench_default.err; tail -c 400 gpurun_out/${tag}_bench_default.json
timeout 300 python tools/phase_profile.py talos_ds_16384 16384 > gpurun_out/${tag}_phase_talos_ds.json 2>/dev/null
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_talos_ds_16384.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-secondary > gpurun_out/${tag}_ncu_launch_talos.log 2>&1
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"k_fast|k_slow" --launch-skip 6 -c 2 -f -o /tmp/ncu/${tag}_talos_ds python bench.py --steps 2 --warmup 3 --no-cpu --no-secondary > gpurun_out/${tag}_ncu_full.log 2>&1
python tools/ncu_summary.py /tmp/ncu/${tag}_talos_ds.ncu-rep > gpurun_out/${tag}_ncu_talos_ds.txt 2> gpurun_out/${tag}_ncu_talos_ds.err
head -32 gpurun_out/${tag}_ncu_talos_ds.txt
