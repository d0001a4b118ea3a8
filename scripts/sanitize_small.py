"""Small, fast pass over every hand-written kernel for compute-sanitizer (memcheck / racecheck / synccheck)."""
import sys
import torch
sys.path.insert(0, ".")
from mimic_b200 import ops

d, H, Hkv = 128, 4, 2
for (B, T, S, causal, active, pad) in [(1, 200, 200, True, True, True), (2, 64, 300, True, True, False), (1, 257, 257, True, False, False)]:
    q = torch.randn(B, H, T, d, device="cuda").bfloat16()
    k, v = (torch.randn(B, Hkv, S, d, device="cuda").bfloat16() for _ in range(2))
    do = torch.randn(B, T, H, d, device="cuda").bfloat16()
    km = None
    if pad:
        km = torch.ones(B, S, dtype=torch.bool, device="cuda")
        km[0, :5] = False
    sh, w, b = (torch.randn(H, d, device="cuda"), torch.randn(H, d, device="cuda") * 0.1, torch.randn(H, device="cuda")) if active else (None,) * 3
    for impl in ("tcgen05", "simt"):
        out, lse_c, lse_f, mu, saved = ops.shifted_attention_forward(q, k, v, km, sh, w, b, causal=causal, active=active, scale=d ** -0.5, impl=impl)
        ops.shifted_attention_backward(saved, out, do, lse_c, lse_f, mu, causal=causal, active=active, scale=d ** -0.5, impl=impl)
L, rows, D = 3, 37, 512
hs = [torch.randn(rows, D, device="cuda").bfloat16().requires_grad_(True) for _ in range(L)]
hp = torch.randn(L, rows, D, device="cuda").bfloat16()
r2c = torch.arange(rows, dtype=torch.int32, device="cuda")
r2c[3] = -1
rw = torch.full((rows,), 1.0 / (L * rows * D), device="cuda")
ops.align_mse(hs, hp, r2c, rw)
ops.align_cos(hs, hp, r2c, rw)
logits = torch.randn(9, 32003, device="cuda").bfloat16().requires_grad_(True)
labels = torch.randint(0, 32003, (9,), device="cuda")
labels[2] = -100
ops.ce_loss(logits, labels)
rows_kl = torch.tensor([3, -1, 0, 7, 8, 1, -1, 2, 5], dtype=torch.int32, device="cuda")
ops.kl_loss(logits, torch.randn(11, 32003, device="cuda").bfloat16(), rows_kl)
src = torch.randn(50, D, device="cuda").bfloat16()
dst = torch.empty(6, D, device="cuda", dtype=torch.bfloat16)
ops.record_gather(src, torch.tensor([0, 7, 49, 33, -1, 5], dtype=torch.int32, device="cuda"), dst)
qkv = torch.randn(2, 9, (8 + 4) * 128, device="cuda").bfloat16()
ops.rope_qk(qkv[..., : 8 * 128].view(2, 9, 8, 128), qkv[..., 8 * 128: 10 * 128].view(2, 9, 2, 128),
            torch.randn(2, 9, 128, device="cuda").bfloat16(), torch.randn(2, 9, 128, device="cuda").bfloat16())
torch.cuda.synchronize()
print("ok")
