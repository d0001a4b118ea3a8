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


def test_random_shapes_persistent_forward(monkeypatch):
    """The same sweep with the persistent-CTA forward forced (MIMIC_FWD_KERNEL=4; it is otherwise only picked for large
    context-pass shapes): several work items per CTA, batch rows and heads changing between items."""
    monkeypatch.setenv("MIMIC_FWD_KERNEL", "4")
    fz = _fuzz()
    rng = random.Random(3)
    for i in range(60):
        tag, errs, bad = fz.one(rng, 30000 + i)
        assert not bad, (tag, bad)


@pytest.mark.parametrize("B,H,Hkv,T,pad", [(2, 16, 4, 1536, True), (3, 12, 12, 1100, False), (5, 8, 2, 2048, True),
                                          (100, 2, 1, 1100, True)])   # last: key-bit tables too big for all rows -> rebuilt per item
def test_persistent_forward_several_items_per_cta(monkeypatch, B, H, Hkv, T, pad):
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
    for name, env in (("persistent", "4"), ("per_tile", "3")):
        monkeypatch.setenv("MIMIC_FWD_KERNEL", env)
        if name == "per_tile":
            monkeypatch.setenv("MIMIC_NO_PERSISTENT", "1")
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


@pytest.mark.skipif(not os.environ.get("MIMIC_TEST_EXPERIMENTAL"),
                    reason="attn_tc_fwd64w.cu (eight softmax warps per CTA, MIMIC_FWD_KERNEL=5) is an experiment that is not "
                           "on any default path (correct, but slower); set MIMIC_TEST_EXPERIMENTAL=1 to include it")
def test_random_shapes_wide_softmax_forward(monkeypatch):
    """The randomised sweep with the experimental wide-softmax forward forced (passed 60 cases when it was written)."""
    monkeypatch.setenv("MIMIC_FWD_KERNEL", "5")
    fz = _fuzz()
    rng = random.Random(5)
    for i in range(80):
        tag, errs, bad = fz.one(rng, 50000 + i)
        assert not bad, (tag, bad)
