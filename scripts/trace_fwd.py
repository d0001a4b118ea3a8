"""Per-role timeline of CTA 0 of the persistent tcgen05 forward kernel, one record per work item
(debug facility mimic_debug_set_trace).  python scripts/trace_fwd.py T B active"""
import sys
import torch
sys.path.insert(0, ".")
from mimic_b200 import ops, _lib

T = int(sys.argv[1]) if len(sys.argv) > 1 else 128
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
active = (sys.argv[3] == "1") if len(sys.argv) > 3 else False
H, d = 32, 128
q, k, v = (torch.randn(B, H, T, d, device="cuda").bfloat16() for _ in range(3))
sh, w, b = (torch.randn(H, d, device="cuda"), torch.randn(H, d, device="cuda") * 0.1, torch.randn(H, device="cuda")) if active else (None,) * 3
for _ in range(2):
    ops.shifted_attention_forward(q, k, v, None, sh, w, b, causal=True, active=active, scale=d ** -0.5)
buf = torch.zeros(24 * 64, dtype=torch.int64, device="cuda")
_lib.lib.mimic_debug_set_trace(buf.data_ptr(), buf.numel())
ops.shifted_attention_forward(q, k, v, None, sh, w, b, causal=True, active=active, scale=d ** -0.5)
torch.cuda.synchronize()
_lib.lib.mimic_debug_set_trace(None, 0)
t = buf.cpu().view(24, 8, 8)
t0 = int(t[t > 0].min())
roles = {0: ("producer", ["q_empty", "q_issued", "kv_issued"]), 1: ("mma", ["q_full", "qk01", "soft0", "done"]),
         2: ("softmax", ["top", "q_full", "tables", "s_full0", "soft0", "steps", "o_full", "epi"])}
for i in range(0, 10):
    print(f"--- item {i}")
    for r, (name, tags) in roles.items():
        vals = [int(t[i, r, k_]) - t0 if int(t[i, r, k_]) else None for k_ in range(len(tags))]
        print(f"  {name:9s} " + "  ".join(f"{tg}={v_}" for tg, v_ in zip(tags, vals)))
