#!/bin/bash
# final tree, two GPUs: whole GPU suite incl. the NCCL data-parallel test, N=2 bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | grep -v "^$" | tail -4 > gpurun_out/r02_final6_tests.log; echo "tests rc=${PIPESTATUS[0]}"; tail -2 gpurun_out/r02_final6_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/r02_final6_bench_n2.json 2> gpurun_out/r02_final6_bench_n2.err; echo "bench rc=$?"
grep -c "nranks 2" gpurun_out/r02_final6_bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r02_final6_ref_n2.json 2> gpurun_out/r02_final6_ref_n2.err; echo "ref rc=$?"; head -c 300 gpurun_out/r02_final6_ref_n2.json
