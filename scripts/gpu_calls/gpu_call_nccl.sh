#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dp_nccl.py -q 2>&1 | tail -60 > gpurun_out/r02_nccl_test.log
for N in 2 4; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_bench_n${N}_idefics2-8b.json 2> gpurun_out/r02_bench_n${N}_idefics2-8b.err
  echo "N=$N rc=$?"
done
