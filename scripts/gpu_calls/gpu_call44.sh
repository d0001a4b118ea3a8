#!/bin/bash
mkdir -p gpurun_out
MIMIC_BWD_DIRECT_DKV=1 timeout 900 python -m pytest tests/test_gpu_attention.py tests/test_gpu_fuzz.py tests/test_gpu_fullsize_oracle.py -q -x 2>&1 | grep -v "^$" | tail -4
timeout 300 python scripts/ab_bwd.py > gpurun_out/r02_c44_bwd_tma.log 2>&1; echo "rc=$?"
MIMIC_BWD_DIRECT_DKV=1 timeout 300 python scripts/ab_bwd.py > gpurun_out/r02_c44_bwd_direct.log 2>&1; echo "rc=$?"
grep -h "active': 1" gpurun_out/r02_c44_bwd_tma.log | cut -c1-150
grep -h "active': 1" gpurun_out/r02_c44_bwd_direct.log | cut -c1-150
MIMIC_BWD_DIRECT_DKV=1 timeout 300 python scripts/time_bwd_small.py 2>&1 | tail -2
