#!/bin/bash
mkdir -p gpurun_out
for v in 0 1 2 0 1; do MIMIC_FWDP_FULL_POLY=$v timeout 300 python scripts/ab_fwd.py > gpurun_out/r02_c59_ab_fpoly$v.log 2>&1; echo "full_poly=$v"; grep "^default" gpurun_out/r02_c59_ab_fpoly$v.log | grep "T': 1024\|T': 4096\|T': 8192" | cut -c1-170; done
