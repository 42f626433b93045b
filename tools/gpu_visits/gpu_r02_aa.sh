#!/bin/bash
tag=r02aa
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_behaviors.py -m gpu -q 2>&1 | tail -12 | tee gpurun_out/${tag}_pytest.txt
