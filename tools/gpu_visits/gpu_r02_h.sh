#!/bin/bash
# round 2, visit H (2 GPUs): the driver's launch of bench.py under torchrun at N = 2 (in-process secondaries on every rank), the
# one-call multi-GPU entry over two real devices, and the last occupancy experiment of the arm kernel
tag=r02h
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${tag}_smi.txt; nvidia-smi topo -m >> gpurun_out/${tag}_smi.txt 2>&1
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 ) > gpurun_out/${tag}_bench_n2.json 2> gpurun_out/${tag}_bench_n2.err
tail -c 2500 gpurun_out/${tag}_bench_n2.json; tail -3 gpurun_out/${tag}_bench_n2.err
( time timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 ) > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err; tail -c 600 gpurun_out/${tag}_bench_reference.json
timeout 600 python tools/multi_gpu_call.py 2 > gpurun_out/${tag}_multi_call.json 2> gpurun_out/${tag}_multi_call.err; cat gpurun_out/${tag}_multi_call.json; tail -3 gpurun_out/${tag}_multi_call.err
for m in 4 5 6; do
  for w in franka_65536 tiago_65536; do
    IWBC_ARM_MINB=$m timeout 300 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu --no-secondary 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w minb$m', round(d['value']), 'e2e', round(d['e2e']['value']))"
  done
done
