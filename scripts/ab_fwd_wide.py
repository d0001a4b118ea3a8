"""A/B in one process: default forward kernels vs the experimental wide-softmax kernel (MIMIC_FWD_KERNEL=5) vs the
one-tile-per-CTA kernel held to one CTA per SM (MIMIC_FWD_ONE_CTA, diagnosis of how well two co-resident CTAs overlap)."""
import os
import sys
import torch
sys.path.insert(0, ".")
from mimic_b200 import ops


def run(B, H, Hkv, T, active, reps=10):
    d = 128
    q = torch.randn(B, H, T, d, device="cuda").bfloat16()
    k, v = (torch.randn(B, Hkv, T, d, device="cuda").bfloat16() for _ in range(2))
    sh, w, b = (torch.randn(H, d, device="cuda"), torch.randn(H, d, device="cuda") * 0.1, torch.randn(H, device="cuda")) if active else (None,) * 3
    fn = lambda: ops.shifted_attention_forward(q, k, v, None, sh, w, b, causal=True, active=active, scale=d ** -0.5)
    out = {}
    for name, env in (("default", None), ("wide", "5"), ("one_cta_per_sm", "one")):
        os.environ.pop("MIMIC_FWD_KERNEL", None)
        os.environ.pop("MIMIC_FWD_ONE_CTA", None)
        os.environ.pop("MIMIC_NO_PERSISTENT", None)
        if env == "one":                      # the one-tile-per-CTA kernel with occupancy forced to 1 (diagnosis)
            os.environ["MIMIC_FWD_ONE_CTA"] = "1"
            os.environ["MIMIC_NO_PERSISTENT"] = "1"
        elif env is not None:
            os.environ["MIMIC_FWD_KERNEL"] = env
        for _ in range(2):
            fn()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(reps):
            fn()
        e.record()
        torch.cuda.synchronize()
        out[name] = a.elapsed_time(e) / reps
    flops = (3 if active else 2) * B * H * T * T * d
    print(f"B{B} H{H}/{Hkv} T{T} active{int(active)}: " + "  ".join(f"{n} {ms * 1e3:7.1f} us {flops / ms / 1e9:6.1f} TF/s" for n, ms in out.items()), flush=True)


run(2, 32, 8, 3200, False)
run(4, 32, 32, 4096, False)
run(4, 32, 32, 4096, True)
