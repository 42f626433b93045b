#!/bin/bash
tag=r02w
mkdir -p gpurun_out
timeout 900 python tools/diag_icub_walk.py 2>&1 | tail -12 | tee gpurun_out/${tag}_icub_walk_event.txt
IWBC_B200_LIB=$PWD/build/variants/libiwbc_bf0.so timeout 900 python tools/diag_icub_walk.py 2>&1 | tail -12 | tee gpurun_out/${tag}_icub_walk_event_bf0.txt
