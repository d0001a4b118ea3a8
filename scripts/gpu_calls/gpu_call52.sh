#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/trace_fwdp_items.py 1024 1 > gpurun_out/r02_c52_items_1024.log 2>&1; echo "rc=$?"
timeout 300 python scripts/trace_fwdp_items.py 4096 1 > gpurun_out/r02_c52_items_4096.log 2>&1; echo "rc=$?"
timeout 300 python scripts/trace_fwdp_items.py 4096 0 > gpurun_out/r02_c52_items_4096_inactive.log 2>&1; echo "rc=$?"
