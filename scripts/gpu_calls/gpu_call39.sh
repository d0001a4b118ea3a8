#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_attention.py tests/test_gpu_fuzz.py tests/test_gpu_fullsize_oracle.py tests/test_gpu_tiny_step.py tests/test_gpu_backbones.py tests/test_gpu_variants.py -q -x 2>&1 | grep -v "^$" | tail -12 > gpurun_out/r02_c39_tests.log
echo "tests rc=${PIPESTATUS[0]}" >> gpurun_out/r02_c39_tests.log
tail -3 gpurun_out/r02_c39_tests.log
timeout 300 python scripts/time_bwd_small.py > gpurun_out/r02_c39_small.log 2>&1; cat gpurun_out/r02_c39_small.log
MIMIC_BWD_NO_DIRECT_DQ=1 timeout 300 python scripts/time_bwd_small.py > gpurun_out/r02_c39_small_nodirect.log 2>&1; cat gpurun_out/r02_c39_small_nodirect.log
