#!/bin/bash
# round-2 call 1: GPU test suite, N=1 bench (new evidence chain + sweep), N=2 bench (exit path, NCCL log), NCCL DP test
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02_c1_gpus.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_dp_nccl.py 2>&1 | tail -40 > gpurun_out/r02_c1_tests.log
echo "tests rc=${PIPESTATUS[0]}" >> gpurun_out/r02_c1_tests.log
timeout 600 python -m pytest tests/test_gpu_dp_nccl.py -q 2>&1 | tail -30 > gpurun_out/r02_c1_nccl.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_c1_bench_n1.json 2> gpurun_out/r02_c1_bench_n1.err
echo "bench n1 rc=$?" >> gpurun_out/r02_c1_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_c1_bench_n2.json 2> gpurun_out/r02_c1_bench_n2.err
echo "bench n2 rc=$?" >> gpurun_out/r02_c1_tests.log
tail -5 gpurun_out/r02_c1_tests.log
