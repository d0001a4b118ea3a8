"""Backward timing: config-5 sweep shapes + the Idefics2 query-pass shape (T = S = 128, GQA 32:8, B = 2), CUDA events."""
import json
import os
import sys

import torch

sys.path.insert(0, ".")
from mimic_b200 import ops  # noqa: E402

dev = "cuda"
PEAK = json.load(open("MEASURED_PEAKS.json"))["bf16_tflops"] if os.path.exists("MEASURED_PEAKS.json") else 1590.0


def timeit(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    return sum(ms) / len(ms), min(ms)


def run(B, H, Hkv, T, active, d=128):
    g = torch.Generator(device=dev).manual_seed(3407 + T)
    q = torch.randn(B, H, T, d, device=dev, generator=g).bfloat16()
    k, v = (torch.randn(B, Hkv, T, d, device=dev, generator=g).bfloat16() for _ in range(2))
    do = torch.randn(B, T, H, d, device=dev, generator=g).bfloat16()
    sh, w, b = torch.randn(H, d, device=dev, generator=g), torch.randn(H, d, device=dev, generator=g) * 0.1, torch.randn(H, device=dev, generator=g)
    args = (sh, w, b) if active else (None, None, None)
    kw = dict(causal=True, active=active, scale=d ** -0.5, impl="tcgen05")
    out, lse_c, lse_f, mu, saved = ops.shifted_attention_forward(q, k, v, None, *args, **kw)
    fn = lambda: ops.shifted_attention_backward(saved, out, do, lse_c, lse_f, mu, **kw)
    ms, mn = timeit(fn)
    fl = (8 if active else 5) * B * H * T * T * d
    return {"B": B, "H": H, "Hkv": Hkv, "T": T, "active": int(active), "ms": round(ms, 4), "ms_min": round(mn, 4),
            "tflops": round(fl / ms / 1e9, 1), "frac_of_burst": round(fl / ms / 1e9 / PEAK, 3)}


if __name__ == "__main__":
    rows = [run(2, 32, 8, 128, True)]
    print("query-pass", rows[-1], flush=True)
    for active in (True, False):
        for T in (512, 1024, 2048, 4096, 8192):
            rows.append(run(16384 // T, 32, 32, T, active))
            print(rows[-1], flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(rows, open("gpurun_out/ab_bwd.json", "w"), indent=1)
