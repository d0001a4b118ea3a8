#!/bin/bash
# final evidence, eight GPUs: the N=8 bench line of the headline config (per-rank step and all-reduce times, NCCL log)
mkdir -p gpurun_out
nvidia-smi -L | wc -l > gpurun_out/r02_final7_n8_gpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 8 --warmup 3 > gpurun_out/r02_final7_bench_n8.json 2> gpurun_out/r02_final7_bench_n8.err
echo "N=8 rc=$?"
grep -c "nranks 8" gpurun_out/r02_final7_bench_n8.err
