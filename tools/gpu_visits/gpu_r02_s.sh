#!/bin/bash
# round 2, visit S (8 GPUs): the driver's launch of bench.py under torchrun at N = 8 (every BASELINE configuration in-process on
# every rank: config 5 = 8 x 131 072 = 1 M walking states), and one host call over the eight devices
tag=r02s
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${tag}_smi.txt; nvidia-smi topo -m >> gpurun_out/${tag}_smi.txt 2>&1
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 3 ) > gpurun_out/${tag}_bench_n8.json 2> gpurun_out/${tag}_bench_n8.err
tail -c 2500 gpurun_out/${tag}_bench_n8.json; tail -3 gpurun_out/${tag}_bench_n8.err
timeout 600 python tools/multi_gpu_call.py 8 > gpurun_out/${tag}_multi_call.json 2> gpurun_out/${tag}_multi_call.err; cat gpurun_out/${tag}_multi_call.json; tail -3 gpurun_out/${tag}_multi_call.err
