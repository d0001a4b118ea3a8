#!/bin/bash
# final evidence, two GPUs: the NCCL data-parallel test and the N=2 bench line (per-rank step and all-reduce times, NCCL log)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dp_nccl.py -q 2>&1 | tail -5 > gpurun_out/r02_final_nccl_test.log; cat gpurun_out/r02_final_nccl_test.log | tail -2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/r02_final_bench_n2.json 2> gpurun_out/r02_final_bench_n2.err
echo "N=2 rc=$?"
grep -c "nranks 2" gpurun_out/r02_final_bench_n2.err
