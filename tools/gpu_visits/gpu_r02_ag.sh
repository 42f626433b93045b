#!/bin/bash
# round 2, visit AG (8 GPUs): one host call over eight devices with fewer / larger chunks per device
tag=r02ag
mkdir -p gpurun_out
for p in 2,6 4,16 30,30; do
  IWBC_PIPE=$p timeout 300 python tools/multi_gpu_call.py 8 2>/dev/null | sed "s/^/pipe=$p /"
done | tee gpurun_out/${tag}_multi_call_pipe.txt
