"""Forward-kernel A/B: one process per kernel family (MIMIC_FWD_KERNEL is read once per process).
    MIMIC_FWD_KERNEL=3 python scripts/ab_fwd.py ; MIMIC_FWD_KERNEL=6 python scripts/ab_fwd.py
Prints TFLOP/s (algorithmic FLOPs) for the config-5 sweep shapes and the Idefics2 context pass."""
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


def run(B, H, Hkv, T, active, pad=False, d=128):
    g = torch.Generator(device=dev).manual_seed(3407 + T)
    q = torch.randn(B, H, T, d, device=dev, generator=g).bfloat16()
    k, v = (torch.randn(B, Hkv, T, d, device=dev, generator=g).bfloat16() for _ in range(2))
    sh, w, b = torch.randn(H, d, device=dev, generator=g), torch.randn(H, d, device=dev, generator=g) * 0.1, torch.randn(H, device=dev, generator=g)
    km = None
    if pad:
        km = torch.ones(B, T, dtype=torch.bool, device=dev)
        km[-1, :13] = False
    args = (sh, w, b) if active else (None, None, None)
    fn = lambda: ops.shifted_attention_forward(q, k, v, km, *args, causal=True, active=active, scale=d ** -0.5, impl="tcgen05")
    ms, mn = timeit(fn)
    fl = (3 if active else 2) * B * H * T * T * d
    return {"B": B, "H": H, "Hkv": Hkv, "T": T, "active": int(active), "pad": int(pad), "ms": round(ms, 4), "ms_min": round(mn, 4),
            "tflops": round(fl / ms / 1e9, 1), "frac_of_burst": round(fl / ms / 1e9 / PEAK, 3)}


if __name__ == "__main__":
    kern = os.environ.get("MIMIC_FWD_KERNEL", "default")
    rows = []
    for active in (True, False):
        for T in (512, 1024, 2048, 4096, 8192):
            rows.append(run(16384 // T, 32, 32, T, active))
            print(kern, rows[-1], flush=True)
    for pad in (False, True):
        rows.append(run(2, 32, 8, 3200, False, pad))
        print(kern, "context-pass", rows[-1], flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(rows, open(f"gpurun_out/ab_fwd_k{kern}.json", "w"), indent=1)
