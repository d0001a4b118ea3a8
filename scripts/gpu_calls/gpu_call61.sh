#!/bin/bash
mkdir -p gpurun_out
for k in 6 3 6 3; do MIMIC_FWD_KERNEL=$k timeout 300 python scripts/ab_fwd.py > gpurun_out/r02_c61_k$k.log 2>&1; echo "k=$k"; grep "active': 0" gpurun_out/r02_c61_k$k.log | cut -c1-175; done
