#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l > gpurun_out/r02_n8_gpus.txt
for WL in idefics2-8b idefics-9b llava-interleave-7b; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 10 --warmup 3 --workload $WL > gpurun_out/r02_bench_n8_$WL.json 2> gpurun_out/r02_bench_n8_$WL.err
  echo "$WL rc=$?"
done
timeout 600 python -m pytest tests/test_gpu_dp_nccl.py -q 2>&1 | tail -3
