#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_attention.py tests/test_gpu_fuzz.py tests/test_gpu_fullsize_oracle.py tests/test_gpu_tiny_step.py -q -x 2>&1 | grep -v "^$" | tail -6 > gpurun_out/r02_c42_tests.log
echo "tests rc=${PIPESTATUS[0]}" >> gpurun_out/r02_c42_tests.log
tail -3 gpurun_out/r02_c42_tests.log
timeout 300 python scripts/ab_bwd.py > gpurun_out/r02_c42_bwd.log 2>&1; echo "bwd rc=$?"
grep "active': 1" gpurun_out/r02_c42_bwd.log | cut -c1-170
timeout 300 python scripts/prof_bwd_parts.py > gpurun_out/r02_c42_parts.log 2>&1; tail -8 gpurun_out/r02_c42_parts.log
timeout 300 python scripts/time_bwd_small.py 2>&1 | tail -2
