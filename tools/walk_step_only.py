#!/usr/bin/env python3
"""ncu helper: the Talos walking workload (config 5, 131 072 states) WITHOUT the one-robot seed walk inside the profiled process'
kernel list: the seed states are generated with profiling paused (cudaProfilerStop / Start), then three walking ticks run.
   ncu --profile-from-start off ... python tools/walk_step_only.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
import inria_wbc_b200 as iw  # noqa: E402

torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
part = bench.Part(iw, torch, "talos_walk_131072", 131072, 0, 0, dev)
part.step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for _ in range(3):
    part.step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("walk steps done, launches", part.bt.launch_count)
