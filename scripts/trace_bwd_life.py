"""Life cycle of every CTA of the tcgen05 backward (debug twin, %globaltimer): per SM, how long a CTA lives, how much of that
is prologue / K,V round trip / iterations / epilogue, and how long the SM sits empty between two CTAs.
    python scripts/trace_bwd_life.py [T] [active]"""
import sys
import torch
sys.path.insert(0, ".")
from mimic_b200 import ops, _lib

T = int(sys.argv[1]) if len(sys.argv) > 1 else 512
active = (sys.argv[2] == "1") if len(sys.argv) > 2 else True
B, H, d = 16384 // T, 32, 128
q, k, v = (torch.randn(B, H, T, d, device="cuda").bfloat16() for _ in range(3))
do = torch.randn(B, T, H, d, device="cuda").bfloat16()
sh, w, b = (torch.randn(H, d, device="cuda"), torch.randn(H, d, device="cuda") * 0.1, torch.randn(H, device="cuda")) if active else (None,) * 3
kw = dict(causal=True, active=active, scale=d ** -0.5, impl="tcgen05")
out, lse_c, lse_f, mu, saved = ops.shifted_attention_forward(q, k, v, None, sh, w, b, **kw)
for _ in range(2):
    ops.shifted_attention_backward(saved, out, do, lse_c, lse_f, mu, **kw)
n_cta = ((T + 127) // 128) * H * B
buf = torch.zeros(2048 + 8 * n_cta, dtype=torch.int64, device="cuda")
_lib.lib.mimic_debug_set_trace(buf.data_ptr(), buf.numel())
ops.shifted_attention_backward(saved, out, do, lse_c, lse_f, mu, **kw)
torch.cuda.synchronize()
_lib.lib.mimic_debug_set_trace(None, 0)
t = buf.cpu()[2048:].view(n_cta, 8).double()
t0 = t[:, 1].min()
life = t[:, 7] - t[:, 1]
print(f"T={T} active={active}: {n_cta} CTAs, kernel span {(t[:, 7].max() - t0) / 1e3:.1f} us (instrumented twin)")
names = ["entry->prologue done", "prologue->K,V landed", "K,V landed->last MMA done", "last MMA->stores issued", "stores issued->smem read", "smem read->exit"]
for i, nm in enumerate(names):
    dlt = t[:, 2 + i] - t[:, 1 + i]
    print(f"  {nm:28s} median {dlt.median() / 1e3:6.2f} us   mean {dlt.mean() / 1e3:6.2f}")
print(f"  CTA life                     median {life.median() / 1e3:6.2f} us   mean {life.mean() / 1e3:6.2f}")
gaps = []
for sm in t[:, 0].unique():
    rows = t[t[:, 0] == sm]
    rows = rows[rows[:, 1].argsort()]
    gaps += list(((rows[1:, 1] - rows[:-1, 7]) / 1e3).tolist())
g = torch.tensor(gaps)
print(f"  gap between CTAs on an SM    median {g.median():6.2f} us   mean {g.mean():6.2f}   ({len(gaps)} gaps)")
