"""the encoder variants other than mimic.yaml's, on the CUDA path, against what the REFERENCE produced
(tests/golden/variants.npz, made by oracle/gen_golden.py::gen_variants) and against the oracle.

  * AttnApproximator with VECTOR_SHIFT | LEARNABLE_SHIFT_SCALE (head-averaged mu, src/shift_encoder.py:689-692, 705-707):
    SA and the unmasked log-sum-exp come from the fused kernel, the gradient of mu re-enters through
    mimic_shifted_attn_bwd_lse;
  * AttnApproximator with VECTOR_SHIFT [| MULTI_HEAD] alone (717-729);
  * the norm-preserving shift hook of AttnFFNShift (222-247, LIVE / licv.yaml) and of AttnApproximator's FFN hook
    (653-672): mimic_norm_shift_fwd / _bwd.
"""
import os

import numpy as np
import pytest
import torch

from oracle import norm_shift as NS
from oracle import shifted_attention as OA
from oracle.ref_loader import make_fake_lmm

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = {torch.float32: 1e-5, torch.bfloat16: 1e-2}
_V, _L, _M = "ShiftStrategy.VECTOR_SHIFT", "ShiftStrategy.LEARNABLE_SHIFT_SCALE", "ShiftStrategy.MULTI_HEAD"
VARIANTS = [("single_head_mu", f"{_V} | {_L}"), ("sh_pad", f"{_V} | {_L}"), ("sh_cache", f"{_V} | {_L}"),
            ("vec_multi", f"{_V} | {_M}"), ("vec_flat", _V)]


def rel(a, b):
    a, b = a.detach().double().cpu(), torch.as_tensor(b).detach().double().cpu()
    return float((a.reshape(b.shape) - b).norm() / (b.norm() + 1e-30))


def _variant(prefix):
    if prefix == "single_head_mu":                   # the round-1 fixture of the head-averaged variant
        z = np.load(os.path.join(GOLDEN, "attn_single_head_mu.npz"))
        return {k: torch.from_numpy(z[k]) for k in z.files}
    z = np.load(os.path.join(GOLDEN, "variants.npz"))
    return {k[len(prefix) + 1:]: torch.from_numpy(z[k]) for k in z.files if k.startswith(prefix + ".")}


def _encoder(g, strategy, layer=1):
    from mimic_b200 import AttnApproximator
    B, H, Hkv, T, S, d = [int(x) for x in g["meta"][:6]]
    enc = AttnApproximator(make_fake_lmm(H * d, 2, H), strategy, None).cuda()
    with torch.no_grad():
        enc.attn_shift[layer].copy_(g["shift"].view(enc.attn_shift[layer].shape))
        if "w" in g:
            enc.log_Z1_lin[layer].weight.copy_(g["w"])
            enc.log_Z1_lin[layer].bias.copy_(g["b"])
    enc.attn_shift_handles[layer].active = True
    return enc


def _oracle(g, q, k, v, d_out, strategy):
    """fp32 oracle on the (possibly bf16-rounded) inputs: forward + autograd of the restatement, itself pinned to the
    reference by tests/test_oracle_golden.py::test_do_shift_variants_oracle_matches_reference."""
    B, H, Hkv, T, S, d, causal = [int(x) for x in g["meta"][:7]]
    multi, learn = _M in strategy, _L in strategy
    q, k, v = (t.clone().requires_grad_(True) for t in (q, k, v))
    shift = g["shift"].clone().requires_grad_(True)
    w = g["w"].clone().requires_grad_(True) if learn else None
    b = g["b"].clone().requires_grad_(True) if learn else None
    kk, vv = OA.repeat_kv(k, H // Hkv), OA.repeat_kv(v, H // Hkv)
    sa, _ = OA.masked_attention(q, kk, vv, OA.visibility(T, S, B, bool(causal), g["key_mask"].bool()), d ** -0.5)
    out = OA.do_shift(q, kk, sa.transpose(1, 2), shift.view(H, d) if multi else shift, w, b, multi_head=multi,
                      learnable_scale=learn).reshape(B, T, H, d)
    leaves = [q, k, v, shift] + ([w, b] if learn else [])
    names = ["dq", "dk", "dv", "dshift"] + (["dw", "db"] if learn else [])
    return out.detach(), dict(zip(names, torch.autograd.grad(out, leaves, d_out)))


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
@pytest.mark.parametrize("key,strategy", VARIANTS, ids=[v[0] for v in VARIANTS])
def test_attn_approximator_variants_match_reference_golden(key, strategy, dtype):
    g = _variant(key)
    B, H, Hkv, T, S, d, causal = [int(x) for x in g["meta"][:7]]
    tol = TOL[dtype]
    rd = (lambda t: t.to(dtype).float())
    q, k, v, d_out = rd(g["q"]), rd(g["k"]), rd(g["v"]), rd(g["d_out"])
    enc = _encoder(g, strategy)
    ql, kl, vl = (t.cuda().to(dtype).requires_grad_(True) for t in (q, k, v))
    km = g["key_mask"].bool()
    out = enc.fused_attention(1, ql, kl, vl, None if bool(km.all()) else km.cuda(), bool(causal))
    assert out.shape == (B, T, H, d) and out.dtype == dtype
    leaves = [ql, kl, vl, enc.attn_shift] + (list(enc.log_Z1_lin[1].parameters()) if "w" in g else [])
    grads = list(torch.autograd.grad(out, leaves, d_out.cuda().to(dtype)))
    grads[3] = grads[3][1]
    names = ["dq", "dk", "dv", "dshift"] + (["dw", "db"] if "w" in g else [])
    want, wg = _oracle(g, q, k, v, d_out, strategy)
    rv = g["row_valid"].bool()[:, :, None, None].expand_as(g["out"])
    assert rel(out.cpu()[rv], want[rv]) < tol
    for name, got in zip(names, grads):
        assert rel(got, wg[name]) < tol * (2.0 if dtype == torch.bfloat16 else 1.0), name
    if dtype == torch.float32:                       # and directly against what the reference itself produced
        assert rel(out.cpu()[rv], g["out"][rv]) < tol
        for name, got in zip(names, grads):
            assert rel(got, g[name]) < tol, "golden " + name


@pytest.mark.parametrize("key,strategy", VARIANTS, ids=[v[0] for v in VARIANTS])
def test_do_shift_variants_api_entry(key, strategy):
    """do_shift(layer, q, k_repeated, attn_output) itself (src/shift_encoder.py:682-732), caller-made attention output."""
    g = _variant(key)
    B, H, Hkv, T, S, d, causal = [int(x) for x in g["meta"][:7]]
    enc = _encoder(g, strategy)
    kk, vv = OA.repeat_kv(g["k"], H // Hkv), OA.repeat_kv(g["v"], H // Hkv)
    sa, _ = OA.masked_attention(g["q"], kk, vv, OA.visibility(T, S, B, bool(causal), g["key_mask"].bool()), d ** -0.5)
    out = enc.do_shift(1, g["q"].cuda(), kk.cuda(), sa.transpose(1, 2).contiguous().cuda())
    assert out.shape == ((B, T, H, d) if _M in strategy else (B, T, H * d))       # the reference's return shapes
    rv = g["row_valid"].bool()[:, :, None, None].expand_as(g["out"])
    assert rel(out.reshape(B, T, H, d).cpu()[rv], g["out"][rv]) < 1e-5
    enc.attn_shift_handles[1].active = False
    x = torch.randn(B, T, H, d, device="cuda")
    assert enc.do_shift(1, g["q"].cuda(), kk.cuda(), x) is x                      # inactive handle: identity (685, 732)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
def test_full_lse_gradient_on_the_tensor_core_path(dtype):
    """d_lse_full through mimic_shifted_attn_bwd_lse at a shape the tcgen05 kernels take (bf16, d = 128, ragged tiles,
    GQA, padding): loss = <out, g_o> + <lse_full, g_l> against autograd of the oracle."""
    from mimic_b200 import ops
    B, H, Hkv, T, S, d = 2, 4, 2, 200, 328, 128
    gen = torch.Generator().manual_seed(77)
    rd = (lambda t: t.to(dtype).float())
    q, k, v = rd(torch.randn(B, H, T, d, generator=gen)), rd(torch.randn(B, Hkv, S, d, generator=gen)), rd(torch.randn(B, Hkv, S, d, generator=gen))
    g_o, g_l = rd(torch.randn(B, T, H, d, generator=gen)), torch.randn(B, H, T, generator=gen)
    km = torch.ones(B, S, dtype=torch.bool)
    km[0, :9] = False
    km[1, 100] = False
    ql, kl, vl = (t.cuda().to(dtype).requires_grad_(True) for t in (q, k, v))
    out, lse = ops.attention_with_full_lse(ql, kl, vl, km.cuda(), causal=True)
    if dtype == torch.bfloat16:
        from mimic_b200 import _lib
        assert _lib.last_attn_impl() == "tcgen05"
    got = torch.autograd.grad([out, lse], [ql, kl, vl], [g_o.cuda().to(dtype), g_l.cuda()])
    qo, ko, vo = (t.clone().requires_grad_(True) for t in (q, k, v))
    kk, vv = OA.repeat_kv(ko, H // Hkv), OA.repeat_kv(vo, H // Hkv)
    sa, _ = OA.masked_attention(qo, kk, vv, OA.visibility(T, S, B, True, km), d ** -0.5)
    lse_w = torch.logsumexp(torch.matmul(qo, kk.transpose(-2, -1)) * d ** -0.5, dim=-1)
    want = torch.autograd.grad([sa.transpose(1, 2), lse_w], [qo, ko, vo], [g_o, g_l])
    tol = TOL[dtype]
    assert rel(out, sa.transpose(1, 2)) < tol
    assert rel(lse, lse_w) < 1e-5 if dtype == torch.float32 else rel(lse, lse_w) < 2e-3
    for name, a, b in zip(("dq", "dk", "dv"), got, want):
        assert rel(a, b) < tol * (2.0 if dtype == torch.bfloat16 else 1.0), name


# ---------------------------------------------------------------- norm-preserving shift
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
@pytest.mark.parametrize("key", ["ns_attn", "ns_ffn", "ns_approx"])
def test_norm_shift_matches_reference_golden(key, dtype):
    from mimic_b200 import ops
    g = _variant(key)
    tol = TOL[dtype]
    rd = (lambda t: t.to(dtype).float())
    h0, dy0 = rd(g["h"]), rd(g["dy"])
    h = h0.cuda().to(dtype).requires_grad_(True)
    shift = g["shift"].cuda().requires_grad_(True)
    scale = g["scale"].cuda().requires_grad_(True) if "scale" in g else None
    n0 = ops.launch_count
    y = ops.norm_shift(h, shift, scale)
    leaves = [h, shift] + ([scale] if scale is not None else [])
    grads = torch.autograd.grad(y, leaves, dy0.cuda().to(dtype))
    assert ops.launch_count - n0 == 2                    # one launch each way
    sc = g["scale"] if "scale" in g else torch.tensor(1.0)
    want = NS.norm_shift_fwd(h0, g["shift"], sc)
    wdh, wds, wdc = NS.norm_shift_bwd(h0, g["shift"], sc, dy0)
    assert y.dtype == dtype and rel(y, want) < tol
    assert rel(grads[0], wdh) < tol * (2.0 if dtype == torch.bfloat16 else 1.0)
    assert rel(grads[1], wds) < tol
    if scale is not None:
        assert rel(grads[2], wdc) < tol
    if dtype == torch.float32:
        assert rel(y, g["y"]) < tol and rel(grads[0], g["dh"]) < tol and rel(grads[1], g["dshift"]) < tol
        if scale is not None and "dscale" in g:
            assert rel(grads[2], g["dscale"]) < tol


@pytest.mark.parametrize("N,D,stride", [(1, 8, 8), (3, 1000, 1000), (257, 2048, 2056), (64, 4096, 3 * 4096),
                                        (19, 8192, 8192), (6400, 4096, 4096)])
def test_norm_shift_shapes_and_properties(N, D, stride):
    """Every vector-count specialisation, ragged widths, strided rows (a slice of a wider projection output), and the
    full step shape (Idefics2 query + context rows).  Size-independent properties: the row norm is preserved; a zero
    scale returns h bit-for-bit; the result equals the oracle on a sample of rows."""
    from mimic_b200 import ops
    gen = torch.Generator().manual_seed(N + D)
    base = torch.randn(N, stride, generator=gen).bfloat16().cuda()
    h = base[:, :D] if stride != D else base
    shift = (torch.randn(D, generator=gen) * 0.7).cuda()
    scale = torch.tensor(0.6, device="cuda")
    y = ops.norm_shift(h, shift, scale)
    assert y.shape == (N, D) and y.is_contiguous()
    hn, yn = h.float().norm(dim=-1), y.float().norm(dim=-1)
    assert float(((yn - hn).abs() / hn).max()) < 4e-3                       # bf16 rounding of the outputs only
    assert torch.equal(ops.norm_shift(h, shift, torch.zeros((), device="cuda")), h)
    rows = torch.randperm(N, generator=gen)[:64]
    want = NS.norm_shift_fwd(h[rows.cuda()].float().cpu(), shift.cpu(), 0.6)
    assert rel(y[rows.cuda()], want) < 1e-2
    # backward: dh on the sample, the column sums over ALL rows (fp64 oracle)
    dy = torch.randn(N, D, generator=gen).bfloat16().cuda()
    hl = h.detach().clone().requires_grad_(True) if stride == D else h.detach().requires_grad_(True)
    sl, cl = shift.clone().requires_grad_(True), scale.clone().requires_grad_(True)
    dh, ds, dc = torch.autograd.grad(ops.norm_shift(hl, sl, cl), [hl, sl, cl], dy)
    wdh, wds, wdc = NS.norm_shift_bwd(h.double().cpu(), shift.double().cpu(), 0.6, dy.double().cpu())
    assert rel(dh[rows.cuda()], wdh[rows]) < 1e-2
    assert rel(ds, wds) < 1e-3 and rel(dc, wdc) < 1e-3


def test_norm_shift_unaligned_parameter_view():
    """The shift row may live at any 4-byte offset of the trainer's flat parameter buffer (e.g. behind a 30-element
    scale vector); the result must not depend on it."""
    from mimic_b200 import ops
    gen = torch.Generator().manual_seed(5)
    flat = torch.randn(3 + 64, generator=gen).cuda()
    h = torch.randn(7, 64, generator=gen).cuda()
    a = ops.norm_shift(h, flat[3:], None)
    b = ops.norm_shift(h, flat[3:].clone(), None)
    assert flat[3:].data_ptr() % 16 != 0 and torch.equal(a, b)
    hl, sl = h.clone().requires_grad_(True), flat.clone().requires_grad_(True)
    g1 = torch.autograd.grad(ops.norm_shift(hl, sl[3:], None), [hl, sl], torch.ones_like(h))
    want = NS.norm_shift_bwd(h.cpu(), flat[3:].cpu(), 1.0, torch.ones(7, 64))
    assert rel(g1[0], want[0]) < 1e-5 and rel(g1[1][3:], want[1]) < 1e-5 and float(g1[1][:3].abs().max()) == 0.0


def test_norm_shift_errors_and_empty():
    from mimic_b200 import ops
    shift = torch.zeros(16, device="cuda")
    assert ops.norm_shift(torch.empty(0, 16, device="cuda"), shift).shape == (0, 16)
    with pytest.raises(ops.MimicError):
        ops.norm_shift(torch.zeros(2, 16), shift.cpu())                      # no CPU fallback
    with pytest.raises(ops.MimicError):
        ops.norm_shift(torch.zeros(2, 12, device="cuda"), torch.zeros(12, device="cuda"))     # D % 8
    with pytest.raises(ops.MimicError):
        ops.norm_shift(torch.zeros(2, 8200, device="cuda"), torch.zeros(8200, device="cuda"))  # D > 8192
    with pytest.raises(ops.MimicError):
        ops.norm_shift(torch.zeros(2, 16, device="cuda", dtype=torch.float16), shift)
