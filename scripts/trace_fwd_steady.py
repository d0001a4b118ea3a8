"""Steady-state timeline of the tcgen05 forward (one-tile-per-CTA kernel, MIMIC_NO_PERSISTENT): steps 8..15 of the
heaviest query tile of the last head, stamped by one softmax thread and by the MMA issuer.
    MIMIC_NO_PERSISTENT=1 python scripts/trace_fwd_steady.py [S] [B] [active]"""
import os
import sys
import torch
os.environ.setdefault("MIMIC_NO_PERSISTENT", "1")
sys.path.insert(0, ".")
from mimic_b200 import ops, _lib

S = int(sys.argv[1]) if len(sys.argv) > 1 else 3200
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
active = (sys.argv[3] == "1") if len(sys.argv) > 3 else False
H, Hkv, d = 32, 8, 128
q = torch.randn(B, H, S, d, device="cuda").bfloat16()
k, v = (torch.randn(B, Hkv, S, d, device="cuda").bfloat16() for _ in range(2))
sh, w, b = (torch.randn(H, d, device="cuda"), torch.randn(H, d, device="cuda") * 0.1, torch.randn(H, device="cuda")) if active else (None,) * 3
for _ in range(2):
    ops.shifted_attention_forward(q, k, v, None, sh, w, b, causal=True, active=active, scale=d ** -0.5)
buf = torch.zeros(64, dtype=torch.int64, device="cuda")
_lib.lib.mimic_debug_set_trace(buf.data_ptr(), buf.numel())
ops.shifted_attention_forward(q, k, v, None, sh, w, b, causal=True, active=active, scale=d ** -0.5)
torch.cuda.synchronize()
_lib.lib.mimic_debug_set_trace(None, 0)
t = buf.cpu().tolist()
t0 = t[24]
print("softmax thread, cycles since its wait began: S seen | scores in regs | max+decision | P stored | SOFT_DONE arrived || next wait begins")
for i in range(4):
    e = [x - t[24 + 6 * i] for x in t[24 + 6 * i: 30 + 6 * i]]
    nxt = (t[24 + 6 * (i + 1)] - t[24 + 6 * i]) if i < 3 else 0
    print(f"  j={8 + i:2d} | {e[1]:5d} | {e[2]:5d} (+{e[2] - e[1]:4d}) | {e[3]:5d} (+{e[3] - e[2]:4d}) | {e[4]:5d} (+{e[4] - e[3]:4d}) | {e[5]:5d} (+{e[5] - e[4]:4d}) || {nxt:5d}")
print("issuer, cycles since the softmax thread's wait for S(8) began: saw SOFT_DONE(j) | V(j) landed | PV(j) issued | K(j+2) landed")
for i in range(4):
    e = [x - t0 for x in t[48 + 4 * i: 52 + 4 * i]]
    print(f"  j={8 + i:2d} | {e[0]:6d} | {e[1]:6d} (+{e[1] - e[0]:4d}) | {e[2]:6d} (+{e[2] - e[1]:4d}) | {e[3]:6d} (+{e[3] - e[2]:4d})")
