import sys, torch
sys.path.insert(0, ".")
from mimic_b200 import ops
d, H = 128, 32
for T in (512, 4096):
    B = 16384 // T
    q, k, v = (torch.randn(B, H, T, d, device="cuda").bfloat16() for _ in range(3))
    do = torch.randn(B, T, H, d, device="cuda").bfloat16()
    sh, w, b = torch.randn(H, d, device="cuda"), torch.randn(H, d, device="cuda") * 0.1, torch.randn(H, device="cuda")
    kw = dict(causal=True, active=True, scale=d ** -0.5, impl="tcgen05")
    out, lse_c, lse_f, mu, saved = ops.shifted_attention_forward(q, k, v, None, sh, w, b, **kw)
    for _ in range(3):
        ops.shifted_attention_backward(saved, out, do, lse_c, lse_f, mu, **kw)
torch.cuda.synchronize()
print("ok")
