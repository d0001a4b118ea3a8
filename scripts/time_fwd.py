"""Forward-only timing of the shapes that matter: context pass (B=2, S=3200, GQA 32:8, inactive) and the sweep shapes
(B*T = 16384, MHA, active / inactive).  20 reps after warm-up, CUDA events, L2 flushed between reps."""
import sys
import torch
sys.path.insert(0, ".")
from mimic_b200 import ops

flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def run(B, H, Hkv, T, active, reps=20):
    d = 128
    q = torch.randn(B, H, T, d, device="cuda").bfloat16()
    k, v = (torch.randn(B, Hkv, T, d, device="cuda").bfloat16() for _ in range(2))
    sh, w, b = (torch.randn(H, d, device="cuda"), torch.randn(H, d, device="cuda") * 0.1, torch.randn(H, device="cuda")) if active else (None,) * 3
    fn = lambda: ops.shifted_attention_forward(q, k, v, None, sh, w, b, causal=True, active=active, scale=d ** -0.5)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        flush.zero_()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        tot += a.elapsed_time(e)
    ms = tot / reps
    flops = (3 if active else 2) * B * H * T * T * d
    print(f"B{B} H{H}/{Hkv} T{T} active{int(active)}: {ms * 1e3:8.1f} us  {flops / ms / 1e9:7.1f} TFLOP/s", flush=True)


run(2, 32, 8, 3200, False)
for T in (2048, 4096, 8192):
    run(16384 // T, 32, 32, T, False)
    run(16384 // T, 32, 32, T, True)
