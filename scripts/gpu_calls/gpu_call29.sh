#!/bin/bash
mkdir -p gpurun_out
python scripts/prof_bwd_parts.py > gpurun_out/r02_c29_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02_c29_parts.csv python scripts/prof_bwd_parts.py > gpurun_out/r02_c29_ncu.log 2>&1
echo rc=$?
