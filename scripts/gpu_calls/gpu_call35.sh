#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_losses.py tests/test_gpu_fullsize.py -q -x 2>&1 | grep -v "^$" | tail -12 > gpurun_out/r02_c35_losses.log
echo "tests rc=${PIPESTATUS[0]}" >> gpurun_out/r02_c35_losses.log
tail -3 gpurun_out/r02_c35_losses.log
timeout 300 python scripts/kernel_sweep.py hbm > gpurun_out/r02_c35_hbm.log 2>&1; echo "hbm rc=$?"
grep "^ce\|^kl" gpurun_out/r02_c35_hbm.log
