"""Per-item timeline of CTA 0 of the persistent ping-pong forward (debug twin, %globaltimer): where the time between two
items goes.  Roles: MMA issuer, epilogue warpgroup, softmax warpgroups of query tile 0 / 1.
    python scripts/trace_fwdp_items.py [T] [active]"""
import sys
import torch
sys.path.insert(0, ".")
from mimic_b200 import ops, _lib

T = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
active = (sys.argv[2] == "1") if len(sys.argv) > 2 else True
B, H, d = 16384 // T, 32, 128
q, k, v = (torch.randn(B, H, T, d, device="cuda").bfloat16() for _ in range(3))
sh, w, b = (torch.randn(H, d, device="cuda"), torch.randn(H, d, device="cuda") * 0.1, torch.randn(H, device="cuda")) if active else (None,) * 3
kw = dict(causal=True, active=active, scale=d ** -0.5, impl="tcgen05")
_lib.lib.mimic_debug_force_fwd_kernel(6)
for _ in range(3):
    ops.shifted_attention_forward(q, k, v, None, sh, w, b, **kw)
buf = torch.zeros(4 * 16 * 16, dtype=torch.int64, device="cuda")
_lib.lib.mimic_debug_set_trace(buf.data_ptr(), buf.numel())
ops.shifted_attention_forward(q, k, v, None, sh, w, b, **kw)
torch.cuda.synchronize()
_lib.lib.mimic_debug_set_trace(None, 0)
_lib.lib.mimic_debug_force_fwd_kernel(0)
t = buf.cpu().view(4, 16, 16).double()
names = {0: ["top", "first QK out", "O_EMPTY0 wait", "O_EMPTY0 got", "O_EMPTY1 wait", "O_EMPTY1 got", "P0(0) q0 got", "P1(0) q0 got",
             "item end", "next QK0 wait", "next QK0 out", "next QK1 wait", "next QK1 out", "P0(jd) q3 got", "P1(jd) q3 got"],
         1: ["top", "ST_FULL0 got", "O_FULL0 got", "O_EMPTY0 sent", "tile0 stored", "ST_FULL1 got", "O_FULL1 got", "O_EMPTY1 sent", "tile1 stored"],
         2: ["top", "Q_FULL got", "z1 done", "S(0) got", "S(1) got", "steps done", "ST_EMPTY got", "stats out", "", "S(last) got"],
         3: ["top", "Q_FULL got", "z1 done", "S(0) got", "S(1) got", "steps done", "ST_EMPTY got", "stats out", "", "S(last) got"]}
role_names = ["MMA issuer", "epilogue", "softmax tile 0", "softmax tile 1"]
print(f"T={T} active={active}: CTA 0, microseconds relative to the MMA issuer's top of each item")
for kk in range(2, 9):
    base = t[0, kk, 0]
    if base == 0:
        break
    print(f"item {kk}: issuer item length {(t[0, kk, 8] - base) / 1e3:6.2f} us; previous item ended {(t[0, kk - 1, 8] - base) / 1e3:+6.2f}")
    for role in range(4):
        ev = [(names[role][e], (t[role, kk, e] - base) / 1e3) for e in range(len(names[role])) if names[role][e] and t[role, kk, e] > 0]
        ev.sort(key=lambda p: p[1])
        print(f"   {role_names[role]:15s} " + "  ".join(f"{n} {v:+.2f}" for n, v in ev))
