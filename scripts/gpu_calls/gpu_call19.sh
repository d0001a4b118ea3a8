#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_losses.py tests/test_gpu_fullsize.py tests/test_gpu_tiny_step.py -q -x 2>&1 | grep -v "^$" | tail -12 > gpurun_out/r02_c19_tests.log
echo "tests rc=${PIPESTATUS[0]}" >> gpurun_out/r02_c19_tests.log
tail -3 gpurun_out/r02_c19_tests.log
timeout 300 python scripts/kernel_sweep.py hbm > gpurun_out/r02_c19_hbm.log 2>&1; echo "hbm rc=$?"
MIMIC_CE_TWO_PASS=1 timeout 300 python scripts/kernel_sweep.py hbm > gpurun_out/r02_c19_hbm_twopass.log 2>&1; echo "hbm2 rc=$?"
timeout 120 python scripts/time_bwd_small.py > gpurun_out/r02_c19_small.log 2>&1; echo "small rc=$?"
