#!/bin/bash
mkdir -p gpurun_out
python scripts/prof_fwdp.py 4096 0 > gpurun_out/r02_c4_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:attn_fwd_tcp -s 2 -c 1 -o gpurun_out/r02_fwdp_T4096_inactive -f python scripts/prof_fwdp.py 4096 0 > gpurun_out/r02_c4_ncu.log 2>&1
echo "ncu rc=$?"
python scripts/prof_fwdp.py 4096 1 >> gpurun_out/r02_c4_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:attn_fwd_tcp -s 2 -c 1 -o gpurun_out/r02_fwdp_T4096_active -f python scripts/prof_fwdp.py 4096 1 >> gpurun_out/r02_c4_ncu.log 2>&1
echo "ncu rc=$?"
ls -la gpurun_out/*.ncu-rep | tail -3
