"""ncu target: a few launches of the tcgen05 backward at one sweep shape (scripts/prof_bwd.py T active B H)."""
import sys
import torch
sys.path.insert(0, ".")
from mimic_b200 import ops

T = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
active = (sys.argv[2] == "1") if len(sys.argv) > 2 else True
B = int(sys.argv[3]) if len(sys.argv) > 3 else 2
H = int(sys.argv[4]) if len(sys.argv) > 4 else 16
d = 128
q, k, v = (torch.randn(B, H, T, d, device="cuda").bfloat16() for _ in range(3))
do = torch.randn(B, T, H, d, device="cuda").bfloat16()
sh, w, b = (torch.randn(H, d, device="cuda"), torch.randn(H, d, device="cuda") * 0.1, torch.randn(H, device="cuda")) if active else (None,) * 3
out, lse_c, lse_f, mu, saved = ops.shifted_attention_forward(q, k, v, None, sh, w, b, causal=True, active=active, scale=d ** -0.5)
for _ in range(4):
    ops.shifted_attention_backward(saved, out, do, lse_c, lse_f, mu, causal=True, active=active, scale=d ** -0.5)
torch.cuda.synchronize()
print("ok")
