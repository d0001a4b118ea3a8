#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3 4 5 6; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29600 + i)) tests/_dp_nccl_worker.py > gpurun_out/r02_c47_run$i.log 2>&1
  echo "run $i rc=$?"
done
timeout 1500 python -m pytest tests/test_gpu_attention.py tests/test_gpu_backbones.py tests/test_gpu_dp_nccl.py -m gpu -q > gpurun_out/r02_c47_three.log 2>&1; echo "three rc=$?"
grep -l "Error\|error" gpurun_out/r02_c47_*.log
