"""Steady-state timeline of the persistent ping-pong forward: CTA 0, its second item, key steps 8..15.
    python scripts/trace_fwdp.py [T] [active]"""
import sys
import torch
sys.path.insert(0, ".")
from mimic_b200 import ops, _lib

T = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
active = (sys.argv[2] == "1") if len(sys.argv) > 2 else False
B, H, d = 16384 // T, 32, 128
q, k, v = (torch.randn(B, H, T, d, device="cuda").bfloat16() for _ in range(3))
sh, w, b = (torch.randn(H, d, device="cuda"), torch.randn(H, d, device="cuda") * 0.1, torch.randn(H, device="cuda")) if active else (None,) * 3
run = lambda: ops.shifted_attention_forward(q, k, v, None, sh, w, b, causal=True, active=active, scale=d ** -0.5, impl="tcgen05")
for _ in range(3):
    run()
buf = torch.zeros(128, dtype=torch.int64, device="cuda")
_lib.lib.mimic_debug_set_trace(buf.data_ptr(), buf.numel())
run()
torch.cuda.synchronize()
_lib.lib.mimic_debug_set_trace(None, 0)
t = buf.cpu().tolist()
t0 = min(x for x in t if x > 0)
rel = lambda x: (x - t0) if x > 0 else -1
print(f"T={T} active={active}: cycles relative to the first stamp")
print("softmax: j | tile0: wait-begin S-seen max-done handed-over | tile1: ...")
for i in range(8):
    a = [rel(x) for x in t[(0 * 8 + i) * 4:(0 * 8 + i) * 4 + 4]]
    c = [rel(x) for x in t[(1 * 8 + i) * 4:(1 * 8 + i) * 4 + 4]]
    print(f"  j={8 + i:2d} | {a[0]:6d} {a[1]:6d} {a[2]:6d} {a[3]:6d}  (wait {a[1]-a[0]:5d} ld+max {a[2]-a[1]:5d} exp {a[3]-a[2]:5d}) | "
          f"{c[0]:6d} {c[1]:6d} {c[2]:6d} {c[3]:6d}  (wait {c[1]-c[0]:5d} ld+max {c[2]-c[1]:5d} exp {c[3]-c[2]:5d})")
print("issuer: j | tile0: P-q0-seen last-PV-issued QK-issued | tile1: ...")
for i in range(8):
    a = [rel(x) for x in t[64 + (0 * 8 + i) * 4:64 + (0 * 8 + i) * 4 + 3]]
    c = [rel(x) for x in t[64 + (1 * 8 + i) * 4:64 + (1 * 8 + i) * 4 + 3]]
    print(f"  j={8 + i:2d} | {a[0]:6d} {a[1]:6d} {a[2]:6d} | {c[0]:6d} {c[1]:6d} {c[2]:6d}")
