#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_attention.py tests/test_gpu_fuzz.py tests/test_gpu_fullsize_oracle.py tests/test_gpu_variants.py -q -x 2>&1 | grep -v "^$" | tail -3
for i in 1 2; do timeout 300 python scripts/ab_fwd.py > gpurun_out/r02_c56_ab$i.log 2>&1; grep "^default" gpurun_out/r02_c56_ab$i.log | grep "active': 1" | cut -c1-175; done
timeout 300 python scripts/trace_fwdp_items.py 4096 1 > gpurun_out/r02_c56_items_4096.log 2>&1; sed -n 2,6p gpurun_out/r02_c56_items_4096.log | cut -c1-330
