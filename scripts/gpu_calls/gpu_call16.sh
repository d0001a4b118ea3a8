#!/bin/bash
mkdir -p gpurun_out
for P in 0 4 3 2; do
  MIMIC_FWDP_POLY=$P MIMIC_FWD_KERNEL=6 timeout 300 python scripts/ab_fwd.py > gpurun_out/r02_c16_poly$P.log 2>&1; echo "poly$P rc=$?"
done
MIMIC_FWDP_POLY=3 timeout 300 python -m pytest tests/test_gpu_attention.py tests/test_gpu_fullsize_oracle.py -q -x 2>&1 | tail -3
