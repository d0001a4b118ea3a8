#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_losses.py tests/test_gpu_fullsize.py tests/test_gpu_tiny_step.py -q -x 2>&1 | grep -v "^$" | tail -8 > gpurun_out/r02_c41_losses.log
echo "tests rc=${PIPESTATUS[0]}" >> gpurun_out/r02_c41_losses.log
tail -3 gpurun_out/r02_c41_losses.log
timeout 300 python scripts/kernel_sweep.py hbm > gpurun_out/r02_c41_hbm.log 2>&1; echo "hbm rc=$?"
grep "^mse" gpurun_out/r02_c41_hbm.log
MIMIC_MSE_NO_PIPE=1 timeout 300 python scripts/kernel_sweep.py hbm > gpurun_out/r02_c41_hbm_nopipe.log 2>&1; echo "hbm rc=$?"
grep "^mse" gpurun_out/r02_c41_hbm_nopipe.log
