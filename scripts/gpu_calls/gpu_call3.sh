#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_attention.py tests/test_gpu_fullsize_oracle.py tests/test_gpu_fullsize.py tests/test_gpu_fuzz.py -q -x 2>&1 | tail -30 > gpurun_out/r02_c3_tests.log
echo "tests rc=${PIPESTATUS[0]}" >> gpurun_out/r02_c3_tests.log
tail -3 gpurun_out/r02_c3_tests.log
if grep -q "tests rc=0" gpurun_out/r02_c3_tests.log; then
  MIMIC_FWD_KERNEL=6 timeout 300 python scripts/ab_fwd.py > gpurun_out/r02_c3_ab6.log 2>&1; echo "ab6 rc=$?"
  MIMIC_FWD_KERNEL=3 timeout 300 python scripts/ab_fwd.py > gpurun_out/r02_c3_ab3.log 2>&1; echo "ab3 rc=$?"
fi
