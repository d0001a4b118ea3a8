#!/bin/bash
# staged CE / KL cluster kernels: parity + sweep; whole-step bf16 three-arm diagnosis; query-pass attention in a graph
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_losses.py tests/test_gpu_fullsize.py tests/test_gpu_tiny_step.py -q -x 2>&1 | grep -v "^$" | tail -12 > gpurun_out/r02_c33_losses.log
echo "tests rc=${PIPESTATUS[0]}" >> gpurun_out/r02_c33_losses.log
tail -3 gpurun_out/r02_c33_losses.log
timeout 300 python scripts/kernel_sweep.py hbm > gpurun_out/r02_c33_hbm.log 2>&1; echo "hbm rc=$?"
grep "^ce\|^kl" gpurun_out/r02_c33_hbm.log
timeout 300 python scripts/time_bwd_small.py > gpurun_out/r02_c33_small.log 2>&1; cat gpurun_out/r02_c33_small.log
timeout 300 python scripts/diag_whole_step_bf16.py llava > gpurun_out/r02_c33_diag_llava.log 2>&1; echo "diag rc=$?"
DIAG_IMPL=simt timeout 300 python scripts/diag_whole_step_bf16.py llava > gpurun_out/r02_c33_diag_llava_simt.log 2>&1; echo "diag rc=$?"
timeout 300 python scripts/diag_whole_step_bf16.py idefics > gpurun_out/r02_c33_diag_idefics.log 2>&1; echo "diag rc=$?"
