#!/bin/bash
mkdir -p gpurun_out
timeout 120 python scripts/trace_bwd_life.py 512 1 > gpurun_out/r02_c28_life512.log 2>&1; echo rc=$?
timeout 120 python scripts/trace_bwd_life.py 4096 1 > gpurun_out/r02_c28_life4096.log 2>&1; echo rc=$?
