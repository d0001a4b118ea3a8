#!/bin/bash
mkdir -p gpurun_out
MIMIC_FWDP_EXCL_POLY=1 timeout 900 python -m pytest tests/test_gpu_attention.py tests/test_gpu_fuzz.py tests/test_gpu_fullsize_oracle.py tests/test_gpu_variants.py -q -x 2>&1 | grep -v "^$" | tail -3
for v in 0 1 0 1; do MIMIC_FWDP_EXCL_POLY=$v timeout 300 python scripts/ab_fwd.py > gpurun_out/r02_c57_ab_poly$v.log 2>&1; echo "poly=$v"; grep "^default" gpurun_out/r02_c57_ab_poly$v.log | grep "active': 1" | cut -c1-170; done
