"""One shape of the forward a few times (for ncu): python scripts/prof_fwdp.py T active [B H Hkv]"""
import sys
import torch
sys.path.insert(0, ".")
from mimic_b200 import ops
T, active = int(sys.argv[1]), bool(int(sys.argv[2]))
B = int(sys.argv[3]) if len(sys.argv) > 3 else 16384 // T
H = int(sys.argv[4]) if len(sys.argv) > 4 else 32
Hkv = int(sys.argv[5]) if len(sys.argv) > 5 else H
d, dev = 128, "cuda"
g = torch.Generator(device=dev).manual_seed(3407 + T)
q = torch.randn(B, H, T, d, device=dev, generator=g).bfloat16()
k, v = (torch.randn(B, Hkv, T, d, device=dev, generator=g).bfloat16() for _ in range(2))
sh, w, b = torch.randn(H, d, device=dev, generator=g), torch.randn(H, d, device=dev, generator=g) * 0.1, torch.randn(H, device=dev, generator=g)
args = (sh, w, b) if active else (None, None, None)
for _ in range(4):
    ops.shifted_attention_forward(q, k, v, None, *args, causal=True, active=active, scale=d ** -0.5, impl="tcgen05")
torch.cuda.synchronize()
print("ok", ops.launch_count)
