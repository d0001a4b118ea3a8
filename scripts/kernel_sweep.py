"""Kernel-level sweep (BASELINE config 5 + the HBM-bound kernels): CUDA-event timing of every kernel family
through the C ABI, >= 20 reps after warm-up, L2 flushed between reps for the HBM-bound kernels (flush_l2), the reference's
own PyTorch path (SDPA + do_shift) on the same GPU for comparison.  Writes gpurun_out/kernel_sweep.json."""
import json
import math
import os
import sys

import torch

sys.path.insert(0, ".")
from mimic_b200 import ops  # noqa: E402
from oracle import shifted_attention as OA  # noqa: E402

PEAK = json.load(open("MEASURED_PEAKS.json")) if os.path.exists("MEASURED_PEAKS.json") else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}
dev = "cuda"
flush_buf = torch.zeros(256 << 20, dtype=torch.uint8, device=dev)
FLUSH = os.environ.get("SWEEP_FLUSH", "read")


def flush_l2():
    """Evict the kernel's inputs from the 126 MB L2 before a rep.  "read" (default): stream 256 MB through L2 with a
    reduction — the cache is left full of CLEAN lines.  "write" (round 1's protocol): memset 256 MB — the cache is left
    full of DIRTY lines whose write-back lands inside the timed kernel.  Measured on the same box
    (profiles/r02_kernel_sweep_hbm_{read,write}flush.json): MSE 0.56 vs 0.52 of the HBM peak at 201 MB, 0.86 vs 0.82 at
    3.2 GB; CE 0.86 vs 0.84; gather 0.95 vs 0.81 at 134 MB."""
    if FLUSH == "write":
        flush_buf.zero_()
    else:
        flush_buf.view(torch.float32).sum()


def time_kernel(fn, key, reps=20, warm=3):
    """Device time of ONE kernel family: events recorded right around the C-ABI call (ops.timer), L2 flushed
    (256 MB written) before every rep."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ops.timer.enabled = True
    ops.timer.reset()
    for _ in range(reps):
        flush_l2()
        fn()
    torch.cuda.synchronize()
    ops.timer.enabled = False
    return ops.timer.summary()[key][1]


def time_in_graph(fns, rounds=5):
    """Per-call device time the way the step runs a kernel: inside a CUDA graph.  `fns` = the same op on DIFFERENT buffer
    sets (together larger than L2, so no call finds its inputs cached); rounds x len(fns) calls captured back to back,
    one event pair around a replay.  No launch latency, no host gap between the event and the kernel — which is what the
    event bracket of a single small launch mostly measures (MSE at 201 MB: 55 us bracketed, 42 us under ncu)."""
    gr, st = torch.cuda.CUDAGraph(), torch.cuda.Stream()
    with torch.cuda.stream(st):
        for fn in fns:
            fn()
        torch.cuda.synchronize()
        with torch.cuda.graph(gr, stream=st):
            for _ in range(rounds):
                for fn in fns:
                    fn()
    torch.cuda.synchronize()
    gr.replay()
    best = None
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        gr.replay()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / (rounds * len(fns))
        best = ms if best is None else min(best, ms)
    return best


def timeit(fn, reps=20, warm=3, flush=False):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        if flush:
            flush_l2()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / reps


def attn(T, active, impl="tcgen05", bwd=False, H=32, d=128, total=16384):
    B = max(1, total // T)
    g = torch.Generator(device=dev).manual_seed(3407 + T)
    q, k, v = (torch.randn(B, H, T, d, device=dev, generator=g).bfloat16() for _ in range(3))
    do = torch.randn(B, T, H, d, device=dev, generator=g).bfloat16()
    sh = torch.randn(H, d, device=dev, generator=g)        # trained-like: mu not ~0
    w = torch.randn(H, d, device=dev, generator=g) * 0.1
    b = torch.randn(H, device=dev, generator=g)
    args = (sh, w, b) if active else (None, None, None)
    out, lse_c, lse_f, mu, saved = ops.shifted_attention_forward(q, k, v, None, *args, causal=True, active=active, scale=d ** -0.5, impl=impl)
    if bwd:
        fn = lambda: ops.shifted_attention_backward(saved, out, do, lse_c, lse_f, mu, causal=True, active=active, scale=d ** -0.5, impl=impl)
        fl = (8 if active else 5) * B * H * T * T * d
    else:
        fn = lambda: ops.shifted_attention_forward(q, k, v, None, *args, causal=True, active=active, scale=d ** -0.5, impl=impl)
        fl = (3 if active else 2) * B * H * T * T * d
    ms = timeit(fn, reps=20 if T <= 4096 else 10)
    return {"B": B, "T": T, "ms": ms, "tflops": fl / ms / 1e9, "frac_of_sustained": fl / ms / 1e9 / PEAK["bf16_tflops_sustained"],
            "frac_of_burst": fl / ms / 1e9 / PEAK["bf16_tflops"]}


def attn_reference(T, active, H=32, d=128, total=16384, bwd=False):
    """The reference's PyTorch path on the GPU: SDPA then do_shift (materialises [B,H,T,T] scores)."""
    B = max(1, total // T)
    q, k, v = (torch.randn(B, H, T, d, device=dev).bfloat16().requires_grad_(bwd) for _ in range(3))
    sh, w, b = torch.randn(H, d, device=dev), torch.randn(H, d, device=dev) * 0.1, torch.randn(H, device=dev)
    do = torch.randn(B, T, H, d, device=dev)

    def fn():
        with torch.autocast("cuda", dtype=torch.bfloat16):     # the reference trains under Lightning mixed precision
            a = torch.nn.functional.scaled_dot_product_attention(q, k, v, is_causal=True).transpose(1, 2)
            o = OA.do_shift(q, k, a, sh, w, b, active=active)
        if bwd:
            torch.autograd.grad(o, (q, k, v), do.to(o.dtype))
    fl = ((11 if bwd else 3) if active else (7 if bwd else 2)) * B * H * T * T * d
    try:
        ms = timeit(fn, reps=5, warm=2)
    except torch.OutOfMemoryError:
        return None
    return {"B": B, "T": T, "ms": ms, "tflops": fl / ms / 1e9}


def mse(L, rows, D, dtype=torch.bfloat16):
    hs = [torch.randn(rows, D, device=dev).to(dtype).requires_grad_(True) for _ in range(L)]
    hp = torch.randn(L, rows, D, device=dev).to(dtype)
    r2c = torch.arange(rows, dtype=torch.int32, device=dev)
    rw = torch.full((rows,), 1.0 / (L * rows * D), device=dev)
    fn = lambda: ops.align_mse(hs, hp, r2c, rw)
    ms = time_kernel(fn, "mse_fwd_bwd")
    nbytes = 3 * L * rows * D * hp.element_size()
    r = {"L": L, "rows": rows, "D": D, "ms": ms, "algorithmic_MB": nbytes / 1e6, "gbs": nbytes / ms / 1e6, "frac_of_hbm": nbytes / ms / 1e6 / PEAK["hbm_gbs"]}
    if nbytes < 1e9:                       # the step's shape: also as the step runs it (4 buffer sets = 0.8 GB)
        sets = [([torch.randn(rows, D, device=dev).to(dtype).requires_grad_(True) for _ in range(L)],
                 torch.randn(L, rows, D, device=dev).to(dtype)) for _ in range(4)]
        g_ms = time_in_graph([(lambda a=a, b=b: ops.align_mse(a, b, r2c, rw)) for a, b in sets])
        r.update(ms_in_graph=g_ms, frac_of_hbm_in_graph=nbytes / g_ms / 1e6 / PEAK["hbm_gbs"])
    return r


def cos(L, rows, D, dtype=torch.bfloat16):
    hs = [torch.randn(rows, D, device=dev).to(dtype).requires_grad_(True) for _ in range(L)]
    hp = torch.randn(L, rows, D, device=dev).to(dtype)
    r2c = torch.arange(rows, dtype=torch.int32, device=dev)
    rw = torch.full((rows,), 1.0 / (L * rows), device=dev)
    ms = time_kernel(lambda: ops.align_cos(hs, hp, r2c, rw), "cos_fwd_bwd")
    nbytes = 3 * L * rows * D * hp.element_size()
    return {"L": L, "rows": rows, "D": D, "ms": ms, "algorithmic_MB": nbytes / 1e6, "gbs": nbytes / ms / 1e6, "frac_of_hbm": nbytes / ms / 1e6 / PEAK["hbm_gbs"]}


def kl(R, V):
    z = (torch.randn(R, V, device=dev) * 2).bfloat16().requires_grad_(True)
    y = (torch.randn(R, V, device=dev) * 2).bfloat16()
    rows = torch.arange(R, dtype=torch.int32, device=dev)
    ms = time_kernel(lambda: ops.kl_loss(z, y, rows), "kl_fwd_bwd")
    nbytes = 3 * R * V * 2
    return {"R": R, "V": V, "ms": ms, "algorithmic_MB": nbytes / 1e6, "gbs": nbytes / ms / 1e6, "frac_of_hbm": nbytes / ms / 1e6 / PEAK["hbm_gbs"]}


def ce(R, V):
    logits = (torch.randn(R, V, device=dev) * 2).bfloat16().requires_grad_(True)
    labels = torch.randint(0, V, (R,), device=dev)
    fn = lambda: ops.ce_loss(logits, labels)
    ms = time_kernel(fn, "ce_fwd_bwd")
    nbytes = 2 * R * V * 2
    r = {"R": R, "V": V, "ms": ms, "algorithmic_MB": nbytes / 1e6, "gbs": nbytes / ms / 1e6, "frac_of_hbm": nbytes / ms / 1e6 / PEAK["hbm_gbs"]}
    if nbytes < 1e8:                       # the Idefics2 step's shape: also as the step runs it (12 buffer sets = 0.4 GB)
        sets = [(torch.randn(R, V, device=dev) * 2).bfloat16().requires_grad_(True) for _ in range(12)]
        g_ms = time_in_graph([(lambda a=a: ops.ce_loss(a, labels)) for a in sets], rounds=2)
        r.update(ms_in_graph=g_ms, frac_of_hbm_in_graph=nbytes / g_ms / 1e6 / PEAK["hbm_gbs"])
    return r


def norm_shift(N, D):
    """AttnFFNShift's hook body (src/shift_encoder.py:233-240): fwd 2*N*D*2 B, bwd 3*N*D*2 B; the reference's own
    eager arithmetic on the same GPU beside it."""
    h = torch.randn(N, D, device=dev).bfloat16().requires_grad_(True)
    dy = torch.randn(N, D, device=dev).bfloat16()
    shift = (torch.randn(D, device=dev) * 0.1).requires_grad_(True)
    scale = torch.tensor(0.5, device=dev, requires_grad=True)

    def fn():
        y = ops.norm_shift(h, shift, scale)
        torch.autograd.grad(y, [h, shift, scale], dy)
    ms_f = time_kernel(fn, "norm_shift_fwd")
    ms_b = time_kernel(fn, "norm_shift_bwd")

    def ref():
        shifted = h + scale * shift
        y = shifted / shifted.norm(dim=-1, keepdim=True) * h.norm(dim=-1, keepdim=True)
        torch.autograd.grad(y, [h, shift, scale], dy.float())
    ms_ref = timeit(ref, flush=True)
    nb = N * D * 2
    return {"N": N, "D": D, "fwd_ms": ms_f, "bwd_ms": ms_b, "fwd_gbs": 2 * nb / ms_f / 1e6, "bwd_gbs": 3 * nb / ms_b / 1e6,
            "fwd_frac_of_hbm": 2 * nb / ms_f / 1e6 / PEAK["hbm_gbs"], "bwd_frac_of_hbm": 3 * nb / ms_b / 1e6 / PEAK["hbm_gbs"],
            "reference_eager_fwd_bwd_ms": ms_ref, "speedup_vs_reference_eager": ms_ref / (ms_f + ms_b)}


def gather(rows_src, N, D):
    src = torch.randn(rows_src, D, device=dev).bfloat16()
    idx = torch.randperm(rows_src, device=dev)[:N].sort().values.to(torch.int32)
    dst = torch.empty(N, D, dtype=torch.bfloat16, device=dev)
    ms = time_kernel(lambda: ops.record_gather(src, idx, dst), "record_gather")
    nbytes = 2 * N * D * 2
    return {"N": N, "D": D, "ms": ms, "gbs": nbytes / ms / 1e6, "frac_of_hbm": nbytes / ms / 1e6 / PEAK["hbm_gbs"]}


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    res = {"peaks": PEAK}
    if which == "norm_shift":
        res["norm_shift"] = [norm_shift(256, 4096), norm_shift(6400, 4096), norm_shift(65536, 4096)]
        for r in res["norm_shift"]:
            print("norm_shift", {a: (round(b, 4) if isinstance(b, float) else b) for a, b in r.items()}, flush=True)
    if which in ("all", "hbm"):
        res["mse"] = [mse(32, 256, 4096), mse(32, 4096, 4096)]
        res["ce"] = [ce(256, 32003), ce(1540, 152000)]
        res["gather"] = [gather(6400, 256, 4096), gather(65536, 16384, 4096)]
        res["cos"] = [cos(32, 256, 4096), cos(32, 4096, 4096)]
        res["kl"] = [kl(256, 32003), kl(1540, 152000)]
        res["norm_shift"] = [norm_shift(256, 4096), norm_shift(6400, 4096), norm_shift(65536, 4096)]
        for k in ("mse", "cos", "ce", "kl", "gather", "norm_shift"):
            for r in res[k]:
                print(k, {a: (round(b, 4) if isinstance(b, float) else b) for a, b in r.items()}, flush=True)
    if which in ("all", "ref"):
        res["attn_reference_pytorch_gpu"] = {}
        for name, active, bwd in (("fwd_active", True, False), ("fwdbwd_active", True, True)):
            res["attn_reference_pytorch_gpu"][name] = [attn_reference(T, active, bwd=bwd) for T in (512, 1024, 2048, 4096)]
            for r in res["attn_reference_pytorch_gpu"][name]:
                print("reference(torch, GPU)", name, r, flush=True)
    if which in ("all", "attn"):
        res["attn"] = {}
        for name, active, bwd in (("fwd_active", True, False), ("bwd_active", True, True), ("fwd_inactive", False, False), ("bwd_inactive", False, True)):
            res["attn"][name] = [attn(T, active, bwd=bwd) for T in (512, 1024, 2048, 4096, 8192)]
            for r in res["attn"][name]:
                print(name, {a: (round(b, 3) if isinstance(b, float) else b) for a, b in r.items()}, flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open(f"gpurun_out/kernel_sweep_{which}.json", "w"), indent=1)
