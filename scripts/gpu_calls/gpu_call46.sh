#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_fullsize.py --deselect tests/test_gpu_fullsize_oracle.py --deselect tests/test_gpu_fuzz.py > gpurun_out/r02_c46_full.log 2>&1
echo "tests rc=$?"
grep -n "Error\|error\|Traceback\|assert\|DP_NCCL" gpurun_out/r02_c46_full.log | head -40
