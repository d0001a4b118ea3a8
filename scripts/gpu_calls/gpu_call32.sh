#!/bin/bash
# KL register kernel: parity + sweep; forward dispatch thresholds (auto); then the whole suite and the N=1 bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_losses.py tests/test_gpu_fullsize.py -q -x 2>&1 | grep -v "^$" | tail -12 > gpurun_out/r02_c32_losses.log
echo "tests rc=${PIPESTATUS[0]}" >> gpurun_out/r02_c32_losses.log
tail -3 gpurun_out/r02_c32_losses.log
timeout 300 python scripts/kernel_sweep.py hbm > gpurun_out/r02_c32_hbm.log 2>&1; echo "hbm rc=$?"
MIMIC_KL_TWO_PASS=1 timeout 300 python scripts/kernel_sweep.py hbm > gpurun_out/r02_c32_hbm_kl2pass.log 2>&1; echo "hbm2 rc=$?"
timeout 300 python scripts/ab_fwd.py > gpurun_out/r02_c32_ab_auto.log 2>&1; echo "ab rc=$?"
timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_gpu_dp_nccl.py 2>&1 | grep -v "^$" | tail -25 > gpurun_out/r02_c32_tests.log
echo "tests rc=${PIPESTATUS[0]}" >> gpurun_out/r02_c32_tests.log
tail -3 gpurun_out/r02_c32_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_c32_bench_n1.json 2> gpurun_out/r02_c32_bench_n1.err; echo "bench rc=$?"
