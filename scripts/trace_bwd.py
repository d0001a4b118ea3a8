"""Per-role timeline of CTA (0,0,0) of the tcgen05 backward kernel (debug facility mimic_debug_set_trace)."""
import sys
import torch
sys.path.insert(0, ".")
from mimic_b200 import ops, _lib

T = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
active = (sys.argv[2] == "1") if len(sys.argv) > 2 else False
B, H, d = 2, 32, 128
q, k, v = (torch.randn(B, H, T, d, device="cuda").bfloat16() for _ in range(3))
do = torch.randn(B, T, H, d, device="cuda").bfloat16()
sh, w, b = (torch.randn(H, d, device="cuda"), torch.randn(H, d, device="cuda") * 0.1, torch.randn(H, device="cuda")) if active else (None,) * 3
out, lse_c, lse_f, mu, saved = ops.shifted_attention_forward(q, k, v, None, sh, w, b, causal=True, active=active, scale=d ** -0.5)
for _ in range(2):
    ops.shifted_attention_backward(saved, out, do, lse_c, lse_f, mu, causal=True, active=active, scale=d ** -0.5)
buf = torch.zeros(24 * 64, dtype=torch.int64, device="cuda")
_lib.lib.mimic_debug_set_trace(buf.data_ptr(), buf.numel())
ops.shifted_attention_backward(saved, out, do, lse_c, lse_f, mu, causal=True, active=active, scale=d ** -0.5)
torch.cuda.synchronize()
_lib.lib.mimic_debug_set_trace(None, 0)
t = buf.cpu().view(24, 8, 8)
t0 = int(t[t > 0].min())
roles = {0: ("producer", ["empty", "issued", "landed"]), 1: ("mma-A", ["-", "-", "-", "soft_done", "dV,dQ,dK issued"]),
         2: ("softmax-g0", ["top", "bar", "s_full", "ldtm", "phase1", "arrived"]), 3: ("softmax-g1", ["top", "bar", "s_full", "ldtm", "phase1", "arrived"]),
         4: ("drain", ["top", "dq_full", "arrived", "staged"]),
         5: ("mma-B", ["qdo_full", "dv_done", "S issued/drain_done", "dP issued"])}
for i in range(4, 12):
    print(f"--- iteration {i}")
    for r, (name, tags) in roles.items():
        vals = [int(t[i, r, k_]) - t0 if int(t[i, r, k_]) else None for k_ in range(len(tags))]
        print(f"  {name:11s} " + "  ".join(f"{tg}={v_}" for tg, v_ in zip(tags, vals)))
