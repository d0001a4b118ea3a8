"""BASELINE.json's full sizes through size-independent properties (the oracle cannot finish these in seconds):
Idefics2 context-pass attention B=2 H=32/8 S=3200, the sweep shape T=4096, the LLaVA logits 1540 x 152000, the
alignment-loss step shape L=32 D=4096.  Checks: two independent kernel families agree (tcgen05 vs SIMT fp32 math),
softmax rows sum to one across every tile, gradient checksums that hold at any size, losses with closed forms."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.detach().double(), b.detach().double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def _qkv(B, H, Hkv, T, d=128, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    q = torch.randn(B, H, T, d, device="cuda", generator=g).bfloat16()
    k = torch.randn(B, Hkv, T, d, device="cuda", generator=g).bfloat16()
    v = torch.randn(B, Hkv, T, d, device="cuda", generator=g).bfloat16()
    return q, k, v


def test_context_pass_shape_tcgen05_agrees_with_simt_and_rows_sum_to_one():
    from mimic_b200 import ops
    B, H, Hkv, S, d = 2, 32, 8, 3200, 128
    q, k, v = _qkv(B, H, Hkv, S)
    km = torch.ones(B, S, dtype=torch.bool, device="cuda")
    km[1, :13] = False                                          # left padding like the synthetic batch
    outs = {}
    for impl in ("tcgen05", "simt"):
        out, lse_c, _, _, _ = ops.shifted_attention_forward(q, k, v, km, None, None, None, causal=True, active=False,
                                                            scale=d ** -0.5, impl=impl)
        outs[impl] = (out.float(), lse_c)
    assert rel(outs["tcgen05"][0], outs["simt"][0]) < 6e-3
    fin = torch.isfinite(outs["simt"][1])
    assert rel(outs["tcgen05"][1][fin], outs["simt"][1][fin]) < 1e-5
    # V = 1  =>  every row with a visible key returns exactly the softmax row sum = 1, whatever the number of tiles
    ones = torch.ones_like(v)
    out1, _, _, _, _ = ops.shifted_attention_forward(q, k, ones, km, None, None, None, causal=True, active=False,
                                                     scale=d ** -0.5, impl="tcgen05")
    valid = torch.ones(B, S, dtype=torch.bool, device="cuda")
    valid[1, :13] = False                                       # padded query rows see no key: defined as 0
    assert float((out1.float()[valid] - 1.0).abs().max()) < 8e-3
    assert float(out1.float()[~valid].abs().max()) == 0.0


@pytest.mark.parametrize("active", [False, True], ids=["inactive", "active"])
def test_sweep_shape_backward_checksums(active):
    """T = 4096, dO = 1: each softmax row sums to one, so sum_s dV[b,kvh,s,c] = (#heads in the group) * T exactly;
    with V constant along keys dP is constant along keys and dS = P (dP - Delta) vanishes: dq = dk = 0 (inactive)."""
    from mimic_b200 import ops
    B, H, Hkv, T, d = 1, 8, 2, 4096, 128
    q, k, v = _qkv(B, H, Hkv, T, seed=1)
    sh, w, b = ((torch.randn(H, d, device="cuda") * 0.5, torch.randn(H, d, device="cuda") * 0.1, torch.randn(H, device="cuda"))
                if active else (None, None, None))
    do = torch.ones(B, T, H, d, device="cuda", dtype=torch.bfloat16)
    res = {}
    for impl in ("tcgen05", "simt"):
        out, lse_c, lse_f, mu, saved = ops.shifted_attention_forward(q, k, v, None, sh, w, b, causal=True, active=active,
                                                                     scale=d ** -0.5, impl=impl)
        res[impl] = ops.shifted_attention_backward(saved, out, do, lse_c, lse_f, mu, causal=True, active=active,
                                                   scale=d ** -0.5, impl=impl)
    dq, dk, dv = res["tcgen05"][:3]
    assert rel(dv.float().sum(dim=2), torch.full((B, Hkv, d), float(H // Hkv * T), device="cuda")) < 2e-3
    for i, name in enumerate(("dq", "dk", "dv")):
        assert rel(res["tcgen05"][i].float(), res["simt"][i].float()) < 1.5e-2, name
    if active:
        for i, name in ((3, "dshift"), (4, "dw"), (5, "db")):
            assert rel(res["tcgen05"][i], res["simt"][i]) < 1e-4, name   # same prep kernel, fp32 accumulation
    else:
        vc = v[:, :, :1].expand_as(v).contiguous()              # V constant along keys
        out, lse_c, lse_f, mu, saved = ops.shifted_attention_forward(q, k, vc, None, None, None, None, causal=True,
                                                                     active=False, scale=d ** -0.5, impl="tcgen05")
        g = torch.randn(B, T, H, d, device="cuda").bfloat16()
        dq0, dk0, _ = ops.shifted_attention_backward(saved, out, g, lse_c, lse_f, mu, causal=True, active=False,
                                                     scale=d ** -0.5, impl="tcgen05")[:3]
        scale_ref = float(res["tcgen05"][0].float().abs().mean()) + 1e-6
        assert float(dq0.float().abs().max()) < 0.05 * max(1.0, scale_ref * 50)   # zero up to bf16 rounding of P and dP
        assert float(dk0.float().abs().max()) < 0.05 * max(1.0, scale_ref * 50)


def test_alignment_loss_step_shape_closed_forms():
    """L = 32, 256 rows, D = 4096 (the Idefics2 step): loss(H, H) = 0, loss(H, H + c) = c^2, dH = 2 w c; cos(H, aH) = 1."""
    from mimic_b200 import ops
    L, rows, D, c = 32, 256, 4096, 0.25
    hp = torch.randn(L, rows, D, device="cuda").bfloat16()
    r2c = torch.arange(rows, dtype=torch.int32, device="cuda")
    w_mse = torch.full((rows,), 1.0 / (L * rows * D), device="cuda")
    same = [hp[l].clone().requires_grad_(True) for l in range(L)]
    assert float(ops.align_mse(same, hp, r2c, w_mse)) == 0.0
    off = [(hp[l].float() + c).bfloat16().requires_grad_(True) for l in range(L)]
    want = float(torch.stack([(o.float() - hp[l].float()).pow(2).mean() for l, o in enumerate(off)]).mean())
    loss = ops.align_mse(off, hp, r2c, w_mse)
    assert abs(float(loss) - want) < 1e-5 * want and abs(want - c * c) < 0.02 * c * c
    g = torch.autograd.grad(loss, off)
    assert rel(torch.stack(g).float(), torch.stack([2.0 / (L * rows * D) * (o.float() - hp[l].float()) for l, o in enumerate(off)])) < 5e-3
    w_cos = torch.full((rows,), 1.0 / (L * rows), device="cuda")
    scaled = [(hp[l].float() * 2.0).bfloat16().requires_grad_(True) for l in range(L)]      # exact in bf16
    assert abs(float(ops.align_cos(scaled, hp, r2c, w_cos))) < 1e-5
    anti = [(-hp[l].float()).bfloat16().requires_grad_(True) for l in range(L)]
    assert abs(float(ops.align_cos(anti, hp, r2c, w_cos)) - 2.0) < 1e-5


def test_llava_logits_shape_closed_forms():
    """1540 x 152000 bf16: uniform logits give CE = log V; the CE gradient of every kept row sums to zero and equals
    (1/V - onehot)/n; KL(p || p) = 0 with zero gradient; rows marked -100 / -1 get exactly zero."""
    from mimic_b200 import ops
    R, V = 1540, 152000
    labels = torch.randint(0, V, (R,), device="cuda")
    labels[::7] = -100
    n = int((labels >= 0).sum())
    flat = torch.zeros(R, V, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    loss = ops.ce_loss(flat, labels)
    assert abs(float(loss) - math.log(V)) < 1e-5 * math.log(V)
    g, = torch.autograd.grad(loss, flat)
    kept = labels >= 0
    assert float(g[~kept].abs().max()) == 0.0
    at_label = g[kept].float().gather(1, labels[kept][:, None])[:, 0]
    assert rel(at_label, torch.full_like(at_label, (1.0 / V - 1.0) / n)) < 5e-3
    z = (torch.randn(R, V, device="cuda") * 2).bfloat16().requires_grad_(True)
    rows = torch.arange(R, dtype=torch.int32, device="cuda")
    rows[::5] = -1
    sel = rows >= 0
    kl = ops.kl_loss(z, z.detach(), rows)
    assert abs(float(kl)) < 1e-5
    gz, = torch.autograd.grad(kl, z)
    assert float(gz.float().abs().max()) < 1e-6
    y = (torch.randn(R, V, device="cuda") * 2).bfloat16()
    z2 = z.detach().clone().requires_grad_(True)
    kl2 = ops.kl_loss(z2, y, rows)
    g2, = torch.autograd.grad(kl2, z2)
    assert float(kl2) > 0 and float(g2[~sel].abs().max()) == 0.0
    assert float(g2[sel].float().sum(dim=1).abs().max()) < 2e-3 / int(sel.sum()) * 50     # softmax(z) - softmax(y) sums to 0 per row
