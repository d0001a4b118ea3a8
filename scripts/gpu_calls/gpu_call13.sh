#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_dp_nccl.py 2>&1 | tail -30 > gpurun_out/r02_c13_tests.log
echo "tests rc=${PIPESTATUS[0]}" >> gpurun_out/r02_c13_tests.log
tail -3 gpurun_out/r02_c13_tests.log
MIMIC_FWD_KERNEL=6 timeout 300 python scripts/ab_fwd.py > gpurun_out/r02_c13_ab6.log 2>&1; echo "ab6 rc=$?"
