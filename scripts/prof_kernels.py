"""A few launches of each hand-written kernel at its bench shape, for `ncu --set full` (one GPU)."""
import os
import sys
import torch
sys.path.insert(0, ".")
from mimic_b200 import ops

d, H = 128, 32
REPS = int(os.environ.get("PROF_REPS", "3"))      # 1 under `ncu --set full` (every launch is replayed ~40 times)
# context-pass forward (Idefics2 step shape): B=2, S=3200, GQA 32:8, inactive, WITH the left-padding mask of the bench batch
# (persistent 64-key kernel) and without it (persistent ping-pong kernel)
q = torch.randn(2, H, 3200, d, device="cuda").bfloat16()
k, v = (torch.randn(2, 8, 3200, d, device="cuda").bfloat16() for _ in range(2))
km = torch.ones(2, 3200, dtype=torch.bool, device="cuda")
km[1, :13] = False
for _ in range(REPS):
    ops.shifted_attention_forward(q, k, v, km, None, None, None, causal=True, active=False, scale=d ** -0.5)
for _ in range(REPS):
    ops.shifted_attention_forward(q, k, v, None, None, None, None, causal=True, active=False, scale=d ** -0.5)
# BASELINE config 5, T = 4096 (B = 4, MHA), active forward (persistent ping-pong kernel) + backward
q, k, v = (torch.randn(4, H, 4096, d, device="cuda").bfloat16() for _ in range(3))
do = torch.randn(4, 4096, H, d, device="cuda").bfloat16()
sh, w, b = torch.randn(H, d, device="cuda"), torch.randn(H, d, device="cuda") * 0.1, torch.randn(H, device="cuda")
for _ in range(REPS):
    out, lse_c, lse_f, mu, saved = ops.shifted_attention_forward(q, k, v, None, sh, w, b, causal=True, active=True, scale=d ** -0.5)
    ops.shifted_attention_backward(saved, out, do, lse_c, lse_f, mu, causal=True, active=True, scale=d ** -0.5)
# the query pass of the step: B = 2, GQA 32:8, T = S = 128, active (single-tile forward; backward with the direct dq path)
q2 = torch.randn(2, H, 128, d, device="cuda").bfloat16()
k2, v2 = (torch.randn(2, 8, 128, d, device="cuda").bfloat16() for _ in range(2))
do2 = torch.randn(2, 128, H, d, device="cuda").bfloat16()
for _ in range(REPS):
    out2, lc2, lf2, mu2, saved2 = ops.shifted_attention_forward(q2, k2, v2, None, sh, w, b, causal=True, active=True, scale=d ** -0.5)
    ops.shifted_attention_backward(saved2, out2, do2, lc2, lf2, mu2, causal=True, active=True, scale=d ** -0.5)
# alignment loss, bandwidth regime: L=32, 4096 rows, D=4096 (3.2 GB algorithmic) and the step shape (201 MB)
for rows in (4096, 256):
    hs = [torch.randn(rows, 4096, device="cuda").bfloat16().requires_grad_(True) for _ in range(32)]
    hp = torch.randn(32, rows, 4096, device="cuda").bfloat16()
    r2c = torch.arange(rows, dtype=torch.int32, device="cuda")
    rw = torch.full((rows,), 1.0 / (32 * rows * 4096), device="cuda")
    for _ in range(REPS):
        ops.align_mse(hs, hp, r2c, rw)
    del hs, hp
logits = torch.randn(1540, 152000, device="cuda").bfloat16().requires_grad_(True)
labels = torch.randint(0, 152000, (1540,), device="cuda")
for _ in range(REPS):
    ops.ce_loss(logits, labels)
# Strategy.LAYER_WISE_COS_SIM at the bandwidth shape, Strategy.LOGITS_KL_DIV at the LLaVA logits shape
hs = [torch.randn(4096, 4096, device="cuda").bfloat16().requires_grad_(True) for _ in range(32)]
hp = torch.randn(32, 4096, 4096, device="cuda").bfloat16()
r2c = torch.arange(4096, dtype=torch.int32, device="cuda")
rw = torch.full((4096,), 1.0 / (32 * 4096), device="cuda")
for _ in range(REPS):
    ops.align_cos(hs, hp, r2c, rw)
del hs, hp
prefix = torch.randn(1540, 152000, device="cuda").bfloat16()
rows = torch.arange(1540, dtype=torch.int32, device="cuda")
for _ in range(REPS):
    ops.kl_loss(logits, prefix, rows)
# RoPE of q and k at the context-pass shape (B=2, T=3200, H=32, Hkv=8, d=128), strided projection-output views
qkv = torch.randn(2, 3200, (32 + 16) * 128, device="cuda").bfloat16()
qv, kv_ = qkv[..., : 32 * 128].view(2, 3200, 32, 128), qkv[..., 32 * 128: 40 * 128].view(2, 3200, 8, 128)
cs = torch.randn(2, 3200, 128, device="cuda").bfloat16()
for _ in range(REPS):
    ops.rope_qk(qv, kv_, cs, cs)
# kernel P (row plan) and kernel R (record gather) at the step shapes
ids = torch.randint(3, 32000, (2, 3200), device="cuda")
ids[:, 1000] = 0
ids[:, 3100] = 0
for _ in range(REPS):
    plan = ops.row_plan(ids, pad_id=0, bos_id=1, left_padding=True, num_separator=1, select="label", cap=256)
src = torch.randn(2, 3200, 4096, device="cuda").bfloat16()
dst = torch.empty(256, 4096, device="cuda", dtype=torch.bfloat16)
for _ in range(REPS):
    ops.record_gather(src, plan.rows, dst)
torch.cuda.synchronize()
print("ok")
