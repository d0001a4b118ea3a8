"""Randomised shape sweep (scripts/fuzz_attention.py): tcgen05 forward / backward against the SIMT fp32-math kernels —
odd and even tile counts, ragged tails, T != S, GQA groups, padding holes, rows without a visible key."""
import importlib.util
import os
import random

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _fuzz():
    spec = importlib.util.spec_from_file_location("fuzz_attention", os.path.join(ROOT, "scripts", "fuzz_attention.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("seed", [1, 2])
def test_random_shapes_tcgen05_vs_simt(seed):
    fz = _fuzz()
    rng = random.Random(seed)
    for i in range(60):
        tag, errs, bad = fz.one(rng, 10000 * seed + i)
        assert not bad, (tag, bad)


@pytest.fixture
def force_kernel():
    """Pin the forward kernel family through the library's debug hook (the environment variable is read once per process)."""
    from mimic_b200 import _lib
    yield _lib.lib.mimic_debug_force_fwd_kernel
    _lib.lib.mimic_debug_force_fwd_kernel(-1)


def test_random_shapes_persistent_forward(force_kernel):
    """The same sweep with the persistent 64-key forward forced (family 4; it is otherwise only picked for large
    context-pass shapes): several work items per CTA, batch rows and heads changing between items."""
    force_kernel(4)
    fz = _fuzz()
    rng = random.Random(3)
    for i in range(60):
        tag, errs, bad = fz.one(rng, 30000 + i)
        assert not bad, (tag, bad)


@pytest.mark.parametrize("B,H,Hkv,T,pad", [(2, 16, 4, 1536, True), (3, 12, 12, 1100, False), (5, 8, 2, 2048, True),
                                          (100, 2, 1, 1100, True)])   # last: key-bit tables too big for all rows -> rebuilt per item
def test_persistent_forward_several_items_per_cta(force_kernel, B, H, Hkv, T, pad):
    """More work items than the 2 x #SMs persistent CTAs: every CTA walks several (query tile, head, batch) items, so the
    Q hand-over (Q_EMPTY), the running K/V ring phases, the per-batch-row key bits and the O reuse across items are all on
    the path.  Persistent kernel vs the one-tile-per-CTA kernel (bit-exact: same arithmetic per tile) and vs SIMT."""
    import torch
    from mimic_b200 import ops
    d = 128
    g = torch.Generator(device="cuda").manual_seed(B * 1000 + T)
    q = torch.randn(B, H, T, d, device="cuda", generator=g).bfloat16()
    k = torch.randn(B, Hkv, T, d, device="cuda", generator=g).bfloat16()
    v = torch.randn(B, Hkv, T, d, device="cuda", generator=g).bfloat16()
    km = None
    if pad:
        km = torch.rand(B, T, device="cuda", generator=g) > 0.1
        km[0, :70] = False
        km[B - 1, :3] = False
    outs = {}
    for name, fam in (("persistent", 4), ("per_tile", 5)):
        force_kernel(fam)
        out, lse_c, _, _, _ = ops.shifted_attention_forward(q, k, v, km, None, None, None, causal=True, active=False,
                                                            scale=d ** -0.5, impl="tcgen05")
        outs[name] = (out.clone(), lse_c.clone())
    assert torch.equal(outs["persistent"][0], outs["per_tile"][0])
    fin = torch.isfinite(outs["per_tile"][1])
    assert torch.equal(outs["persistent"][1][fin], outs["per_tile"][1][fin])
    ref, lse_ref, _, _, _ = ops.shifted_attention_forward(q, k, v, km, None, None, None, causal=True, active=False,
                                                          scale=d ** -0.5, impl="simt")
    num = (outs["persistent"][0].double() - ref.double()).norm()
    assert float(num / ref.double().norm()) < 6e-3


def test_random_shapes_ping_pong_forward(force_kernel):
    """The randomised sweep with the persistent ping-pong forward (attn_tc_fwdp.cu, family 6) forced on every shape with more
    than one query tile: ragged tiles, T != S, GQA, padding holes, several items per CTA, single-tile last pairs."""
    force_kernel(6)
    fz = _fuzz()
    rng = random.Random(5)
    for i in range(80):
        tag, errs, bad = fz.one(rng, 50000 + i)
        assert not bad, (tag, bad)


@pytest.mark.parametrize("B,H,Hkv,T,pad", [(2, 16, 4, 1536, True), (5, 8, 2, 2048, True), (3, 12, 12, 1100, False)])
def test_ping_pong_forward_several_items_per_cta(force_kernel, B, H, Hkv, T, pad):
    """Ping-pong kernel with more items than CTAs (dynamic item ring, Q / O hand-over across items, K / V ring phases running
    on) against the one-tile-per-CTA 64-key kernel and the SIMT fp32-math kernel."""
    import torch
    from mimic_b200 import ops
    d = 128
    g = torch.Generator(device="cuda").manual_seed(B * 1000 + T + 1)
    q = torch.randn(B, H, T, d, device="cuda", generator=g).bfloat16()
    k = torch.randn(B, Hkv, T, d, device="cuda", generator=g).bfloat16()
    v = torch.randn(B, Hkv, T, d, device="cuda", generator=g).bfloat16()
    km = None
    if pad:
        km = torch.rand(B, T, device="cuda", generator=g) > 0.1
        km[0, :70] = False
        km[B - 1, :3] = False
    sh, w, b = torch.randn(H, d, device="cuda", generator=g) * 0.5, torch.randn(H, d, device="cuda", generator=g) * 0.1, torch.randn(H, device="cuda", generator=g)
    for active in (False, True):
        args = (sh, w, b) if active else (None, None, None)
        res = {}
        for name, fam, impl in (("pp", 6, "tcgen05"), ("tile", 5, "tcgen05"), ("simt", -1, "simt")):
            force_kernel(fam)
            out, lse_c, lse_f, mu, _ = ops.shifted_attention_forward(q, k, v, km, *args, causal=True, active=active,
                                                                     scale=d ** -0.5, impl=impl)
            res[name] = (out.float(), lse_c, lse_f, mu)
        fin = torch.isfinite(res["simt"][1])
        for other in ("tile", "simt"):
            num = (res["pp"][0].double() - res[other][0].double()).norm() / res[other][0].double().norm()
            assert float(num) < 6e-3, (active, other, float(num))
            assert float((res["pp"][1][fin] - res[other][1][fin]).abs().max()) < 2e-3, (active, other, "lse_c")
            if active:
                assert float((res["pp"][2] - res[other][2]).abs().max()) < 2e-3, (active, other, "lse_full")
                assert float((res["pp"][3] - res[other][3]).abs().max()) < 2e-3, (active, other, "mu")
