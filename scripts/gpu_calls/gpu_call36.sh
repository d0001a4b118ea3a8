#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_losses.py tests/test_gpu_fullsize.py tests/test_gpu_backbones.py -q -x 2>&1 | grep -v "^$" | tail -12 > gpurun_out/r02_c36_losses.log
echo "tests rc=${PIPESTATUS[0]}" >> gpurun_out/r02_c36_losses.log
tail -3 gpurun_out/r02_c36_losses.log
timeout 300 python scripts/kernel_sweep.py hbm > gpurun_out/r02_c36_hbm.log 2>&1; echo "hbm rc=$?"
grep "^ce\|^kl" gpurun_out/r02_c36_hbm.log
python scripts/prof_ce.py > gpurun_out/r02_c36_plain.log 2>&1 && python scripts/prof_kl.py >> gpurun_out/r02_c36_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:ce_fwd_bwd_reg -s 2 -c 1 -o /tmp/r02_ce_stage -f python scripts/prof_ce.py > gpurun_out/r02_c36_ncu.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:kl_fwd_bwd_reg -s 2 -c 1 -o /tmp/r02_kl_stage -f python scripts/prof_kl.py >> gpurun_out/r02_c36_ncu.log 2>&1
echo "ncu rc=$?"
for k in ce kl; do
  ncu -i /tmp/r02_${k}_stage.ncu-rep --page raw --csv > gpurun_out/r02_c36_${k}_raw.csv 2>/dev/null
  ncu -i /tmp/r02_${k}_stage.ncu-rep --page source --csv > gpurun_out/r02_c36_${k}_source.csv 2>/dev/null
done
