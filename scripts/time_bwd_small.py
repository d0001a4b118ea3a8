"""Steady-state cost of the query-pass backward (T = S = 128, GQA 32:8, B = 2): N calls streamed back to back, one pair of
events around all of them (what a CUDA-graph replay pays), plus the same for the forward."""
import sys
import torch
sys.path.insert(0, ".")
from mimic_b200 import ops
dev, d = "cuda", 128
B, H, Hkv, T = 2, 32, 8, 128
g = torch.Generator(device=dev).manual_seed(1)
q = torch.randn(B, H, T, d, device=dev, generator=g).bfloat16()
k, v = (torch.randn(B, Hkv, T, d, device=dev, generator=g).bfloat16() for _ in range(2))
do = torch.randn(B, T, H, d, device=dev, generator=g).bfloat16()
sh, w, b = torch.randn(H, d, device=dev, generator=g), torch.randn(H, d, device=dev, generator=g) * 0.1, torch.randn(H, device=dev, generator=g)
kw = dict(causal=True, active=True, scale=d ** -0.5, impl="tcgen05")
out, lse_c, lse_f, mu, saved = ops.shifted_attention_forward(q, k, v, None, sh, w, b, **kw)
N = 200
for name, fn in (("bwd", lambda: ops.shifted_attention_backward(saved, out, do, lse_c, lse_f, mu, **kw)),
                 ("fwd", lambda: ops.shifted_attention_forward(q, k, v, None, sh, w, b, **kw))):
    gr = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        with torch.cuda.graph(gr, stream=s):
            for _ in range(N):
                fn()
    torch.cuda.synchronize()
    for _ in range(2):
        gr.replay()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    gr.replay()
    e.record()
    torch.cuda.synchronize()
    print(f"{name}: {a.elapsed_time(e) / N * 1e3:.1f} us per call in a graph of {N}")
