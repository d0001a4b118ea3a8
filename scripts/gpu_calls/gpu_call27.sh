#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_attention.py tests/test_gpu_fullsize_oracle.py tests/test_gpu_fullsize.py tests/test_gpu_fuzz.py tests/test_gpu_variants.py tests/test_gpu_tiny_step.py -q -x 2>&1 | grep -v "^$" | tail -12 > gpurun_out/r02_c27_tests.log
echo "tests rc=${PIPESTATUS[0]}" >> gpurun_out/r02_c27_tests.log
tail -3 gpurun_out/r02_c27_tests.log
timeout 300 python scripts/ab_bwd.py > gpurun_out/r02_c27_bwd.log 2>&1; echo "bwd rc=$?"
timeout 120 python scripts/time_bwd_small.py > gpurun_out/r02_c27_small.log 2>&1; echo "small rc=$?"
