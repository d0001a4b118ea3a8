"""GPU-side data-parallel correctness: two NCCL ranks with the CUDA AdamW end bit-identical and equal to the
mean-gradient step.  Needs two GPUs (skipped on a one-GPU box; `gpurun --gpus 2 -- python -m pytest tests/test_gpu_dp_nccl.py`)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_nccl_ranks_identical_and_equal_mean_gradient_step():
    import socket
    with socket.socket() as sk:                          # a port that is free right now (fixed ports collide with other jobs on a shared box)
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "_dp_nccl_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + "\n" + r.stderr[-3000:]
    assert "DP_NCCL_OK world=2" in r.stdout, r.stdout[-3000:]
