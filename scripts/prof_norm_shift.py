"""A few launches of the norm-preserving shift kernels (AttnFFNShift hook body) for `ncu --set full` (one GPU)."""
import sys
import torch
sys.path.insert(0, ".")
from mimic_b200 import ops

for N in (65536, 6400):
    h = torch.randn(N, 4096, device="cuda").bfloat16().requires_grad_(True)
    dy = torch.randn(N, 4096, device="cuda").bfloat16()
    shift = (torch.randn(4096, device="cuda") * 0.1).requires_grad_(True)
    scale = torch.tensor(0.5, device="cuda", requires_grad=True)
    for _ in range(3):
        y = ops.norm_shift(h, shift, scale)
        torch.autograd.grad(y, [h, shift, scale], dy)
torch.cuda.synchronize()
print("ok")
