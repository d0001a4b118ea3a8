#!/bin/bash
mkdir -p gpurun_out
for v in 1 2 3 4 2; do MIMIC_FWDP_EXCL_POLY=$v timeout 300 python scripts/ab_fwd.py > gpurun_out/r02_c58_ab_poly$v.log 2>&1; echo "poly=$v"; grep "^default" gpurun_out/r02_c58_ab_poly$v.log | grep "active': 1" | grep "T': 2048\|T': 4096\|T': 8192" | cut -c1-170; done
