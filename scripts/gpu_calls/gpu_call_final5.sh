#!/bin/bash
# final evidence, one GPU: whole GPU suite, bench line (+ reference arm), ncu table of every kernel, step launch list, sweeps
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02_final5_gpus.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q  2>&1 | grep -v "^$" | tail -25 > gpurun_out/r02_final5_tests.log
echo "tests rc=${PIPESTATUS[0]}" >> gpurun_out/r02_final5_tests.log
tail -3 gpurun_out/r02_final5_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02_final5_smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py > gpurun_out/r02_final5_bench_n1.json 2> gpurun_out/r02_final5_bench_n1.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_final5_bench_reference_arm.json 2> gpurun_out/r02_final5_bench_reference_arm.err; echo "ref rc=$?"
PROF_REPS=1 python scripts/prof_kernels.py > gpurun_out/r02_final5_prof_plain.log 2>&1 &&
PROF_REPS=1 ncu --set full --clock-control none -k regex:'attn_|mse_|ce_|kl_|cos_|rope_|row_plan|record_gather|convert_dq' -o /tmp/r02_kernels -f python scripts/prof_kernels.py > gpurun_out/r02_final5_prof_ncu.log 2>&1
echo "ncu kernels rc=$?"
ncu -i /tmp/r02_kernels.ncu-rep --page raw --csv > gpurun_out/r02_kernels_raw.csv 2>/dev/null
python bench.py --profile-steps 3 --no-cpu-baseline > gpurun_out/r02_final5_launch_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches.csv python bench.py --profile-steps 3 --no-cpu-baseline > gpurun_out/r02_final5_launch_ncu.log 2>&1
echo "ncu launches rc=$?"
gzip -f gpurun_out/r02_launches.csv
timeout 900 python scripts/kernel_sweep.py all > gpurun_out/r02_final5_sweep_all.log 2>&1; echo "sweep rc=$?"
timeout 300 python scripts/ab_fwd.py > gpurun_out/r02_final5_ab_fwd.log 2>&1; echo "abf rc=$?"
timeout 300 python scripts/ab_bwd.py > gpurun_out/r02_final5_ab_bwd.log 2>&1; echo "abb rc=$?"
du -sh gpurun_out
