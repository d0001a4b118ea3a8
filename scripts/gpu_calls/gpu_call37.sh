#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_losses.py tests/test_gpu_fullsize.py -q -x 2>&1 | grep -v "^$" | tail -12 > gpurun_out/r02_c37_losses.log
echo "tests rc=${PIPESTATUS[0]}" >> gpurun_out/r02_c37_losses.log
tail -3 gpurun_out/r02_c37_losses.log
MIMIC_CE_NT=512 timeout 600 python -m pytest tests/test_gpu_losses.py tests/test_gpu_fullsize.py -q -x -k "ce" 2>&1 | grep -v "^$" | tail -4
for rep in 1 2; do
timeout 300 python scripts/kernel_sweep.py hbm > gpurun_out/r02_c37_hbm.log 2>&1; echo "hbm rc=$?"
grep "^ce\|^kl" gpurun_out/r02_c37_hbm.log
MIMIC_CE_NT=512 timeout 300 python scripts/kernel_sweep.py hbm > gpurun_out/r02_c37_hbm512.log 2>&1; echo "hbm512 rc=$?"
grep "^ce" gpurun_out/r02_c37_hbm512.log
done
