"""A few launches of each hand-written kernel at its bench shape, for `ncu --set full` (one GPU)."""
import sys
import torch
sys.path.insert(0, ".")
from mimic_b200 import ops

d, H = 128, 32
# context-pass forward (Idefics2 step shape): B=2, S=3200, GQA 32:8, inactive
q = torch.randn(2, H, 3200, d, device="cuda").bfloat16()
k, v = (torch.randn(2, 8, 3200, d, device="cuda").bfloat16() for _ in range(2))
for _ in range(3):
    ops.shifted_attention_forward(q, k, v, None, None, None, None, causal=True, active=False, scale=d ** -0.5)
# sweep shape, active forward + backward: B=2, T=4096, MHA
q, k, v = (torch.randn(2, H, 4096, d, device="cuda").bfloat16() for _ in range(3))
do = torch.randn(2, 4096, H, d, device="cuda").bfloat16()
sh, w, b = torch.randn(H, d, device="cuda"), torch.randn(H, d, device="cuda") * 0.1, torch.randn(H, device="cuda")
for _ in range(3):
    out, lse_c, lse_f, mu, saved = ops.shifted_attention_forward(q, k, v, None, sh, w, b, causal=True, active=True, scale=d ** -0.5)
    ops.shifted_attention_backward(saved, out, do, lse_c, lse_f, mu, causal=True, active=True, scale=d ** -0.5)
# alignment loss, bandwidth regime: L=32, 4096 rows, D=4096 (3.2 GB algorithmic) and the step shape (201 MB)
for rows in (4096, 256):
    hs = [torch.randn(rows, 4096, device="cuda").bfloat16().requires_grad_(True) for _ in range(32)]
    hp = torch.randn(32, rows, 4096, device="cuda").bfloat16()
    r2c = torch.arange(rows, dtype=torch.int32, device="cuda")
    rw = torch.full((rows,), 1.0 / (32 * rows * 4096), device="cuda")
    for _ in range(3):
        ops.align_mse(hs, hp, r2c, rw)
    del hs, hp
logits = torch.randn(1540, 152000, device="cuda").bfloat16().requires_grad_(True)
labels = torch.randint(0, 152000, (1540,), device="cuda")
for _ in range(3):
    ops.ce_loss(logits, labels)
# Strategy.LAYER_WISE_COS_SIM at the bandwidth shape, Strategy.LOGITS_KL_DIV at the LLaVA logits shape
hs = [torch.randn(4096, 4096, device="cuda").bfloat16().requires_grad_(True) for _ in range(32)]
hp = torch.randn(32, 4096, 4096, device="cuda").bfloat16()
r2c = torch.arange(4096, dtype=torch.int32, device="cuda")
rw = torch.full((4096,), 1.0 / (32 * 4096), device="cuda")
for _ in range(3):
    ops.align_cos(hs, hp, r2c, rw)
del hs, hp
prefix = torch.randn(1540, 152000, device="cuda").bfloat16()
rows = torch.arange(1540, dtype=torch.int32, device="cuda")
for _ in range(3):
    ops.kl_loss(logits, prefix, rows)
# RoPE of q and k at the context-pass shape (B=2, T=3200, H=32, Hkv=8, d=128), strided projection-output views
qkv = torch.randn(2, 3200, (32 + 16) * 128, device="cuda").bfloat16()
qv, kv_ = qkv[..., : 32 * 128].view(2, 3200, 32, 128), qkv[..., 32 * 128: 40 * 128].view(2, 3200, 8, 128)
cs = torch.randn(2, 3200, 128, device="cuda").bfloat16()
for _ in range(3):
    ops.rope_qk(qv, kv_, cs, cs)
torch.cuda.synchronize()
print("ok")
