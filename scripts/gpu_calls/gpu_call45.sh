#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | grep -v "^$" | tail -15 > gpurun_out/r02_c45_tests_2gpu.log
echo "tests rc=${PIPESTATUS[0]}" >> gpurun_out/r02_c45_tests_2gpu.log
tail -4 gpurun_out/r02_c45_tests_2gpu.log
