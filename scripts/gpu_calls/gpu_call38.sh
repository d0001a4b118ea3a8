#!/bin/bash
mkdir -p gpurun_out
SWEEP_FLUSH=read timeout 300 python scripts/kernel_sweep.py hbm > gpurun_out/r02_c38_hbm_readflush.log 2>&1; echo "hbm rc=$?"
cp gpurun_out/kernel_sweep_hbm.json gpurun_out/r02_kernel_sweep_hbm_readflush.json
SWEEP_FLUSH=write timeout 300 python scripts/kernel_sweep.py hbm > gpurun_out/r02_c38_hbm_writeflush.log 2>&1; echo "hbm rc=$?"
cp gpurun_out/kernel_sweep_hbm.json gpurun_out/r02_kernel_sweep_hbm_writeflush.json
grep -h "^mse\|^cos\|^ce\|^kl\|^gather" gpurun_out/r02_c38_hbm_readflush.log gpurun_out/r02_c38_hbm_writeflush.log
timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_gpu_dp_nccl.py 2>&1 | grep -v "^$" | tail -25 > gpurun_out/r02_c38_tests.log
echo "tests rc=${PIPESTATUS[0]}" >> gpurun_out/r02_c38_tests.log
tail -3 gpurun_out/r02_c38_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_c38_bench_n1.json 2> gpurun_out/r02_c38_bench_n1.err; echo "bench rc=$?"
