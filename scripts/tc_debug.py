"""Bring-up harness for the tcgen05 attention kernels: each case compares impl=tcgen05 with
impl=simt (same inputs) and with the CPU oracle, forward statistics included.  Run under gpurun."""
import sys
import time

import torch

sys.path.insert(0, ".")
from mimic_b200 import ops  # noqa: E402
from oracle import shifted_attention as OA  # noqa: E402


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def case(name, B, H, Hkv, T, S, causal, active, pad=False, seed=0, oracle=True):
    g = torch.Generator().manual_seed(seed)
    d = 128
    q = torch.randn(B, H, T, d, generator=g).bfloat16()
    k = torch.randn(B, Hkv, S, d, generator=g).bfloat16()
    v = torch.randn(B, Hkv, S, d, generator=g).bfloat16()
    shift, w, b = torch.randn(H, d, generator=g) * 0.5, torch.randn(H, d, generator=g) * 0.2, torch.randn(H, generator=g)
    km = None
    if pad:
        km = torch.ones(B, S, dtype=torch.bool)
        km[0, :5] = False
        km[B - 1, S // 2] = False
    args = [t.cuda() for t in (q, k, v)]
    kmc = None if km is None else km.cuda()
    sh, ww, bb = (shift.cuda(), w.cuda(), b.cuda()) if active else (None, None, None)
    res = {}
    for impl in ("simt", "tcgen05"):
        out, lse_c, lse_f, mu, _ = ops.shifted_attention_forward(*args, kmc, sh, ww, bb, causal=causal, active=active,
                                                                 scale=d ** -0.5, impl=impl)
        torch.cuda.synchronize()
        res[impl] = (out.float(), lse_c, lse_f, mu)
    a, t_ = res["simt"], res["tcgen05"]
    fin = torch.isfinite(a[1])
    msg = f"{name:28s} tc-vs-simt out={rel(t_[0], a[0]):.2e} lse_c={rel(t_[1][fin], a[1][fin]):.2e}"
    if active:
        msg += f" lse_full={rel(t_[2], a[2]):.2e} mu={rel(t_[3], a[3]):.2e}"
    if oracle:
        want = OA.shifted_attention_fwd(q.float(), k.float(), v.float(), shift, w, b, causal=causal, key_mask=km, active=active)
        msg += f" | tc-vs-oracle out={rel(t_[0], want['out']):.2e}"
    print(msg, flush=True)
    return rel(t_[0], a[0])


def bcase(name, B, H, Hkv, T, S, causal, active, pad=False, seed=0):
    """backward: tcgen05 vs simt on the same forward statistics, and vs the oracle."""
    g = torch.Generator().manual_seed(seed)
    d = 128
    q = torch.randn(B, H, T, d, generator=g).bfloat16()
    k = torch.randn(B, Hkv, S, d, generator=g).bfloat16()
    v = torch.randn(B, Hkv, S, d, generator=g).bfloat16()
    do = torch.randn(B, T, H, d, generator=g).bfloat16()
    shift, w, b = torch.randn(H, d, generator=g) * 0.5, torch.randn(H, d, generator=g) * 0.2, torch.randn(H, generator=g)
    km = None
    if pad:
        km = torch.ones(B, S, dtype=torch.bool)
        km[0, :5] = False
        km[B - 1, S // 2] = False
        vis = OA.visibility(T, S, B, causal, km).any(-1)[:, 0]
        do = do * vis[:, :, None, None]
    args = [t.cuda() for t in (q, k, v)]
    kmc = None if km is None else km.cuda()
    sh, ww, bb = (shift.cuda(), w.cuda(), b.cuda()) if active else (None, None, None)
    out, lse_c, lse_f, mu, saved = ops.shifted_attention_forward(*args, kmc, sh, ww, bb, causal=causal, active=active,
                                                                 scale=d ** -0.5, impl="tcgen05")
    res = {}
    for impl in ("simt", "tcgen05"):
        r = ops.shifted_attention_backward(saved, out, do.cuda(), lse_c, lse_f, mu, causal=causal, active=active,
                                           scale=d ** -0.5, impl=impl)
        torch.cuda.synchronize()
        res[impl] = r
    names = ["dq", "dk", "dv"] + (["dshift", "dw", "db"] if active else [])
    msg = f"{name:28s} tc-vs-simt " + " ".join(f"{n}={rel(res['tcgen05'][i], res['simt'][i]):.1e}" for i, n in enumerate(names))
    want = OA.shifted_attention_bwd(q.float(), k.float(), v.float(), shift, w, b, do.float(), causal=causal, key_mask=km, active=active)
    msg += " | tc-vs-oracle " + " ".join(f"{n}={rel(res['tcgen05'][i], want[n]):.1e}" for i, n in enumerate(names[:3]))
    print(msg, flush=True)


def bperf(B, H, Hkv, T, active, reps=5):
    d = 128
    q = torch.randn(B, H, T, d, device="cuda").bfloat16()
    k = torch.randn(B, Hkv, T, d, device="cuda").bfloat16()
    v = torch.randn(B, Hkv, T, d, device="cuda").bfloat16()
    do = torch.randn(B, T, H, d, device="cuda").bfloat16()
    sh, ww, bb = (torch.randn(H, d, device="cuda") * 0.5, torch.randn(H, d, device="cuda") * 0.2, torch.randn(H, device="cuda")) if active else (None,) * 3
    out, lse_c, lse_f, mu, saved = ops.shifted_attention_forward(q, k, v, None, sh, ww, bb, causal=True, active=active, scale=d ** -0.5, impl="tcgen05")
    for impl in ("tcgen05",):
        for _ in range(2):
            ops.shifted_attention_backward(saved, out, do, lse_c, lse_f, mu, causal=True, active=active, scale=d ** -0.5, impl=impl)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            ops.shifted_attention_backward(saved, out, do, lse_c, lse_f, mu, causal=True, active=active, scale=d ** -0.5, impl=impl)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        fl = (8 if active else 5) * B * H * T * T * d
        print(f"bwd perf {impl} B={B} H={H} T={T} active={int(active)}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s (prep+main+convert)", flush=True)


def perf(B, H, Hkv, T, active, reps=10, mask=False):
    d = 128
    km = None
    if mask:
        km = torch.ones(B, T, dtype=torch.bool, device="cuda")
        km[B - 1, :13] = False
    q = torch.randn(B, H, T, d, device="cuda").bfloat16()
    k = torch.randn(B, Hkv, T, d, device="cuda").bfloat16()
    v = torch.randn(B, Hkv, T, d, device="cuda").bfloat16()
    sh, ww, bb = (torch.randn(H, d, device="cuda") * 0.5, torch.randn(H, d, device="cuda") * 0.2, torch.randn(H, device="cuda")) if active else (None,) * 3
    for impl in ("tcgen05",):
        for _ in range(3):
            ops.shifted_attention_forward(q, k, v, km, sh, ww, bb, causal=True, active=active, scale=d ** -0.5, impl=impl)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            ops.shifted_attention_forward(q, k, v, km, sh, ww, bb, causal=True, active=active, scale=d ** -0.5, impl=impl)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        fl = (3 if active else 2) * B * H * T * T * d
        print(f"perf {impl} B={B} H={H} T={T} active={int(active)} mask={int(mask)}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "basic"):
        case("1tile noncausal inactive", 1, 1, 1, 128, 128, False, False)
        case("1tile causal inactive", 1, 1, 1, 128, 128, True, False)
        case("2tile noncausal inactive", 1, 1, 1, 128, 256, False, False)
        case("1tile causal active", 1, 1, 1, 128, 128, True, True)
        case("4tile causal active", 1, 2, 2, 512, 512, True, True)
        case("ragged gqa pad active", 2, 8, 2, 200, 200, True, True, pad=True)
        case("ragged gqa pad inactive", 2, 8, 2, 333, 333, True, False, pad=True)
        case("T<S cache active", 2, 4, 4, 64, 448, True, True, pad=True)
        case("T=1 decode", 2, 4, 2, 1, 300, False, True)
        case("ctx 3200 inactive", 1, 4, 1, 3200, 3200, True, False, oracle=False)
    if which in ("all", "bwd"):
        bcase("b 1tile noncausal inactive", 1, 1, 1, 64, 128, False, False)
        bcase("b 1tile causal inactive", 1, 1, 1, 128, 128, True, False)
        bcase("b 1tile causal active", 1, 1, 1, 128, 128, True, True)
        bcase("b 4tile causal active", 1, 2, 2, 512, 512, True, True)
        bcase("b ragged gqa pad active", 2, 8, 2, 200, 200, True, True, pad=True)
        bcase("b ragged gqa pad inactive", 2, 8, 2, 333, 333, True, False, pad=True)
        bcase("b T<S cache active", 2, 4, 4, 64, 448, True, True, pad=True)
        bcase("b T=1 decode", 2, 4, 2, 1, 300, False, True)
    if which in ("all", "bperf"):
        bperf(2, 32, 8, 128, True)
        bperf(4, 32, 32, 4096, False)
        bperf(4, 32, 32, 4096, True)
        bperf(2, 32, 32, 8192, True)
    if which == "bperf1":
        bperf(4, 32, 32, 4096, False)
        bperf(4, 32, 32, 4096, True)
    if which in ("all", "perf"):
        perf(2, 32, 8, 3200, False)
        perf(2, 32, 8, 3200, False, mask=True)
        perf(2, 32, 8, 3200, True)
        perf(4, 32, 32, 4096, False)
        perf(4, 32, 32, 4096, True)
        perf(2, 32, 32, 8192, True)
    if which == "cta":
        perf(64, 32, 32, 128, False)      # 2048 CTAs x 1 tile: per-CTA fixed cost
        perf(32, 32, 32, 256, False)      # 2048 CTAs x avg 1.5 tiles (v1) / 1024 CTAs x 3 (v2)
        perf(16, 32, 32, 512, False)
