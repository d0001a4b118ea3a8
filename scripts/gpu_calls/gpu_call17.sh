#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_gpu_dp_nccl.py 2>&1 | grep -v "^$" | tail -25 > gpurun_out/r02_c17_tests.log
echo "tests rc=${PIPESTATUS[0]}" >> gpurun_out/r02_c17_tests.log
tail -3 gpurun_out/r02_c17_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_c17_bench_n1.json 2> gpurun_out/r02_c17_bench_n1.err; echo "bench rc=$?"
