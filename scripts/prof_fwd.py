"""ncu target: a few launches of the tcgen05 forward at one shape (scripts/prof_fwd.py T B active Hkv)."""
import sys
import torch
sys.path.insert(0, ".")
from mimic_b200 import ops

T = int(sys.argv[1]) if len(sys.argv) > 1 else 3200
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
active = (sys.argv[3] == "1") if len(sys.argv) > 3 else False
Hkv = int(sys.argv[4]) if len(sys.argv) > 4 else 8
H, d = 32, 128
q = torch.randn(B, H, T, d, device="cuda").bfloat16()
k, v = (torch.randn(B, Hkv, T, d, device="cuda").bfloat16() for _ in range(2))
sh, w, b = (torch.randn(H, d, device="cuda"), torch.randn(H, d, device="cuda") * 0.1, torch.randn(H, device="cuda")) if active else (None,) * 3
for _ in range(4):
    ops.shifted_attention_forward(q, k, v, None, sh, w, b, causal=True, active=active, scale=d ** -0.5)
torch.cuda.synchronize()
print("ok")
