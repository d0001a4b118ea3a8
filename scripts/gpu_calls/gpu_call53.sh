#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_attention.py tests/test_gpu_fuzz.py tests/test_gpu_fullsize_oracle.py -q -x 2>&1 | grep -v "^$" | tail -4
timeout 300 python scripts/ab_fwd.py > gpurun_out/r02_c53_ab.log 2>&1; echo "ab rc=$?"
grep "^default" gpurun_out/r02_c53_ab.log | cut -c1-175
timeout 300 python scripts/trace_fwdp_items.py 1024 1 > gpurun_out/r02_c53_items_1024.log 2>&1; echo "rc=$?"
