#!/bin/bash
tag=r02ad
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_cpp_shim.py -m gpu -q 2>&1 | tail -15 | tee gpurun_out/${tag}_pytest.txt
