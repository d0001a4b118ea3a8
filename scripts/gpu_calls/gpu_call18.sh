#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_fuzz.py tests/test_gpu_attention.py tests/test_gpu_fullsize.py -q -x 2>&1 | grep -v "^$" | tail -12 > gpurun_out/r02_c18_tests.log
echo "tests rc=${PIPESTATUS[0]}" >> gpurun_out/r02_c18_tests.log
tail -3 gpurun_out/r02_c18_tests.log
timeout 300 python scripts/ab_bwd.py > gpurun_out/r02_c18_bwd.log 2>&1; echo "bwd rc=$?"
MIMIC_FWD_KERNEL=3 timeout 300 python scripts/ab_fwd.py > gpurun_out/r02_c18_ab3.log 2>&1; echo "ab3 rc=$?"
