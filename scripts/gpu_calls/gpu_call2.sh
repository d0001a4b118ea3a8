#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --deselect tests/test_gpu_dp_nccl.py -s 2>&1 | grep -v "^$" | tail -60 > gpurun_out/r02_c2_tests.log
echo "tests rc=${PIPESTATUS[0]}" >> gpurun_out/r02_c2_tests.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_c2_bench_n1.json 2> gpurun_out/r02_c2_bench_n1.err
echo "bench n1 rc=$?" >> gpurun_out/r02_c2_tests.log
tail -5 gpurun_out/r02_c2_tests.log
