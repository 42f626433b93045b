#!/bin/bash
tag=r02al
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_behaviors.py -m gpu -q -k icub_walk 2>&1 | tail -40 | tee gpurun_out/${tag}_pytest.txt
timeout 600 python tools/diag_icub_walk.py 2>&1 | tail -14 | tee gpurun_out/${tag}_icub_walk_event.txt
