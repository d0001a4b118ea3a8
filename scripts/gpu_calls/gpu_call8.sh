#!/bin/bash
mkdir -p gpurun_out
MIMIC_FWD_KERNEL=6 timeout 300 python scripts/ab_fwd.py > gpurun_out/r02_c8_dyn.log 2>&1; echo "dyn rc=$?"
MIMIC_FWDP_STATIC=1 MIMIC_FWD_KERNEL=6 timeout 300 python scripts/ab_fwd.py > gpurun_out/r02_c8_static.log 2>&1; echo "static rc=$?"
