#!/bin/bash
# round 2, visit Q: full GPU suite (log kept) incl. the torque task
tag=r02q
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rA 2>&1 | tail -60 > gpurun_out/${tag}_pytest_gpu.txt; tail -8 gpurun_out/${tag}_pytest_gpu.txt
