#!/bin/bash
# the co-residency check of the 64-key forward (family 5 = one query tile per CTA, 64-key tiles): two CTAs per SM vs one
mkdir -p gpurun_out
MIMIC_FWD_KERNEL=5 timeout 300 python scripts/ab_fwd.py > gpurun_out/r02_c40_k5_two_ctas.log 2>&1; echo "rc=$?"
MIMIC_FWD_ONE_CTA=1 MIMIC_FWD_KERNEL=5 timeout 300 python scripts/ab_fwd.py > gpurun_out/r02_c40_k5_one_cta.log 2>&1; echo "rc=$?"
grep -h "T': 4096\|3200" gpurun_out/r02_c40_k5_two_ctas.log gpurun_out/r02_c40_k5_one_cta.log | cut -c1-170
timeout 300 python scripts/kernel_sweep.py hbm > gpurun_out/r02_c40_hbm.log 2>&1; echo "hbm rc=$?"
grep "^mse\|^ce" gpurun_out/r02_c40_hbm.log
