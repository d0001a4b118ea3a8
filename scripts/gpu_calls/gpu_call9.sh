#!/bin/bash
mkdir -p gpurun_out
python scripts/prof_fwdp.py 4096 0 > gpurun_out/r02_c9_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:attn_fwd_tcp -s 2 -c 1 -o gpurun_out/r02_fwdp2_T4096_inactive -f python scripts/prof_fwdp.py 4096 0 > gpurun_out/r02_c9_ncu.log 2>&1
echo "ncu rc=$?"
python scripts/prof_fwdp.py 1024 0 >> gpurun_out/r02_c9_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:attn_fwd_tcp -s 2 -c 1 -o gpurun_out/r02_fwdp2_T1024_inactive -f python scripts/prof_fwdp.py 1024 0 >> gpurun_out/r02_c9_ncu.log 2>&1
echo "ncu rc=$?"
