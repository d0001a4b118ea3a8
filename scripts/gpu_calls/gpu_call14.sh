#!/bin/bash
mkdir -p gpurun_out
timeout 120 python scripts/trace_fwdp.py 4096 0 > gpurun_out/r02_c14_trace_in.log 2>&1; echo "rc=$?"
timeout 120 python scripts/trace_fwdp.py 4096 1 > gpurun_out/r02_c14_trace_ac.log 2>&1; echo "rc=$?"
