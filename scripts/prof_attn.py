"""Short driver for ncu: a few launches of the tcgen05 attention forward / backward at sweep shape."""
import sys
import torch
sys.path.insert(0, ".")
from mimic_b200 import ops

which = sys.argv[1] if len(sys.argv) > 1 else "fwd"
T = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
active = (sys.argv[3] == "1") if len(sys.argv) > 3 else True
B, H, d = 2, 32, 128
q, k, v = (torch.randn(B, H, T, d, device="cuda").bfloat16() for _ in range(3))
do = torch.randn(B, T, H, d, device="cuda").bfloat16()
sh, w, b = (torch.randn(H, d, device="cuda") * 0.5, torch.randn(H, d, device="cuda") * 0.2, torch.randn(H, device="cuda")) if active else (None,) * 3
for _ in range(3):
    out, lse_c, lse_f, mu, saved = ops.shifted_attention_forward(q, k, v, None, sh, w, b, causal=True, active=active, scale=d ** -0.5, impl="tcgen05")
    if which == "bwd":
        ops.shifted_attention_backward(saved, out, do, lse_c, lse_f, mu, causal=True, active=active, scale=d ** -0.5, impl="tcgen05")
torch.cuda.synchronize()
print("ok")
