#!/bin/bash
mkdir -p gpurun_out
MIMIC_FWDP_MUFU_TURNS=1 timeout 600 python -m pytest tests/test_gpu_attention.py tests/test_gpu_fullsize_oracle.py -q -x 2>&1 | grep -v "^$" | tail -2
for v in 0 1 0 1; do MIMIC_FWDP_MUFU_TURNS=$v timeout 300 python scripts/ab_fwd.py > gpurun_out/r02_c55_ab_turns$v.log 2>&1; echo "turns=$v"; grep "^default" gpurun_out/r02_c55_ab_turns$v.log | grep "T': 4096\|T': 8192\|T': 1024\|3200" | cut -c1-170; done
