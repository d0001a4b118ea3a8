#!/bin/bash
mkdir -p gpurun_out
PROF_REPS=1 python scripts/prof_kernels.py > gpurun_out/r02_prof_plain.log 2>&1 &&
PROF_REPS=1 ncu --set full --clock-control none -k regex:'attn_|mse_|ce_|kl_|cos_|rope_|row_plan|record_gather|convert_dq' -o /tmp/r02_kernels -f python scripts/prof_kernels.py > gpurun_out/r02_prof_ncu.log 2>&1
echo "ncu kernels rc=$?"
ncu -i /tmp/r02_kernels.ncu-rep --page raw --csv > gpurun_out/r02_kernels_raw.csv 2>/dev/null
ls -la /tmp/r02_kernels.ncu-rep
python bench.py --profile-steps 3 --no-cpu-baseline > gpurun_out/r02_launch_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches.csv python bench.py --profile-steps 3 --no-cpu-baseline > gpurun_out/r02_launch_ncu.log 2>&1
echo "ncu launches rc=$?"
gzip -f gpurun_out/r02_launches.csv
timeout 600 python scripts/kernel_sweep.py all > gpurun_out/r02_sweep_all.log 2>&1; echo "sweep rc=$?"
du -sh gpurun_out
