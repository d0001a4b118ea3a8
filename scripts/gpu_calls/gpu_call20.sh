#!/bin/bash
mkdir -p gpurun_out
python scripts/prof_ce.py > gpurun_out/r02_c20_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:ce_fwd_bwd -s 2 -c 1 -o gpurun_out/r02_ce_reg -f python scripts/prof_ce.py > gpurun_out/r02_c20_ncu.log 2>&1
echo "ncu rc=$?"
