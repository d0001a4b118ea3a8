"""Timeline of one CTA (query tile 0 of head 0, batch 0) of the tcgen05 forward kernel (debug facility
mimic_debug_set_trace): where do the ~7 us of per-tile fixed cost go?   python scripts/trace_fwd.py T B active"""
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
buf = torch.zeros(64, dtype=torch.int64, device="cuda")
_lib.lib.mimic_debug_set_trace(buf.data_ptr(), buf.numel())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
ops.shifted_attention_forward(q, k, v, None, sh, w, b, causal=True, active=active, scale=d ** -0.5)
e1.record()
torch.cuda.synchronize()
_lib.lib.mimic_debug_set_trace(None, 0)
t = buf.cpu().view(8, 8)
t0 = int(t[0, 0])
names = {(0, 0): "cta entry", (0, 1): "prologue done (barriers, TMEM alloc, key bits, sync)", (2, 0): "Q landed", (2, 1): "first S tile ready",
         (2, 2): "first softmax done", (2, 3): "all steps done", (2, 4): "O complete", (2, 5): "output tile handed to TMA", (0, 2): "cta exit"}
print(f"kernel {e0.elapsed_time(e1) * 1e3:.1f} us for {B * H * ((T + 127) // 128)} CTAs; query tile 0 (1-2 key tiles):")
for key, name in sorted(names.items(), key=lambda kv: int(t[kv[0][0], kv[0][1]])):
    print(f"  {int(t[key[0], key[1]]) - t0:7d} cycles  {name}")
