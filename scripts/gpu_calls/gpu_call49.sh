#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29800 + i)) tests/_dp_nccl_worker.py > gpurun_out/r02_c49_run$i.log 2>&1
  echo "run $i rc=$? $(grep -c DP_NCCL_OK gpurun_out/r02_c49_run$i.log)"
  grep "AssertionError" gpurun_out/r02_c49_run$i.log | head -2
done
nvidia-smi topo -m 2>&1 | head -8
