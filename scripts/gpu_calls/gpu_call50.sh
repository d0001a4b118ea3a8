#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3 4; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29900 + i)) tests/_dp_nccl_worker.py > gpurun_out/r02_c50_run$i.log 2>&1
  echo "run $i rc=$? $(grep -c DP_NCCL_OK gpurun_out/r02_c50_run$i.log)"
  grep "AssertionError" gpurun_out/r02_c50_run$i.log | head -2
done
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | grep -v "^$" | tail -6 > gpurun_out/r02_c50_tests.log; echo "tests rc=${PIPESTATUS[0]}"; tail -3 gpurun_out/r02_c50_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r02_c50_bench_n2.json 2> gpurun_out/r02_c50_bench_n2.err; echo "bench rc=$?"
