#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_losses.py tests/test_gpu_fullsize.py tests/test_gpu_tiny_step.py -q -x 2>&1 | grep -v "^$" | tail -25 > gpurun_out/r02_c43_losses.log
echo "tests rc=${PIPESTATUS[0]}" >> gpurun_out/r02_c43_losses.log
tail -25 gpurun_out/r02_c43_losses.log
