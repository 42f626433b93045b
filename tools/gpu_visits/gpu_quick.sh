#!/bin/bash
# quick GPU check during kernel work: parity tests, the three Talos / iCub bench lines, the iCub loop profile
tag=${1:-q}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for w in talos_ds_16384 talos_walk_131072 icub_ds_16384; do
  timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu > gpurun_out/${tag}_bench_$w.json 2> gpurun_out/${tag}_bench_$w.err
  python -c "
import json,sys
d=json.loads(open('gpurun_out/${tag}_bench_$w.json').read().strip().splitlines()[-1]); print('$w', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['config']['mean_active_set_iterations'])"
done
timeout 300 python tools/phase_profile.py icub_ds_16384 16384 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(json.dumps(d['mean_cycles_per_instance'].get('loop'))); print(d['mean_cycles_per_instance']['active_set'])"
