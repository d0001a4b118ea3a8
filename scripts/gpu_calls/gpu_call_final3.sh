#!/bin/bash
# final tree re-check, one GPU: whole GPU suite, smoke, bench line, reference arm
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | grep -v "^$" | tail -6 > gpurun_out/r02_final3_tests.log
echo "tests rc=${PIPESTATUS[0]}" >> gpurun_out/r02_final3_tests.log
tail -3 gpurun_out/r02_final3_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02_final3_smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py > gpurun_out/r02_final3_bench_n1.json 2> gpurun_out/r02_final3_bench_n1.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_final3_bench_reference_arm.json 2> gpurun_out/r02_final3_bench_reference_arm.err; echo "ref rc=$?"
