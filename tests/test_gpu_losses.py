"""Parity of the HBM-bound kernels (record gather, fused MSE, fused CE, AdamW) with the oracle."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import losses as OL

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.mark.parametrize("dtype,dst", [(torch.float32, torch.float32), (torch.bfloat16, torch.bfloat16),
                                       (torch.float32, torch.bfloat16)])
def test_record_gather(dtype, dst):
    from mimic_b200 import ops
    torch.manual_seed(0)
    src = torch.randn(3, 50, 256, device="cuda").to(dtype)
    rows = torch.tensor([0, 7, 149, 33, -1, 50], dtype=torch.int32, device="cuda")
    out = torch.full((6, 256), 7.0, dtype=dst, device="cuda")
    ops.record_gather(src, rows, out)
    want = src.reshape(-1, 256)[rows.clamp_min(0).long()].to(dst)
    want[4] = 0
    assert torch.equal(out, want)


def _mse_case(dtype, L, B, T, D, counts, seed=0):
    g = torch.Generator().manual_seed(seed)
    hs = [torch.randn(B, T, D, generator=g).to(dtype) for _ in range(L)]
    mask = torch.zeros(B, T, dtype=torch.bool)
    for b, c in enumerate(counts):
        mask[b, T - c:] = True
    N = int(mask.sum())
    hp = torch.randn(L, N, D, generator=g).to(dtype)
    return hs, mask, hp


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
@pytest.mark.parametrize("shape", [(3, 2, 8, 64, (5, 6)), (4, 2, 40, 1024, (33, 17)), (2, 1, 16, 4096, (16,))])
def test_align_mse_matches_oracle(dtype, shape):
    from mimic_b200 import ops
    L, B, T, D, counts = shape
    hs, mask, hp = _mse_case(dtype, L, B, T, D, counts)
    # oracle on the same (rounded) values, fp32
    hs32 = [h.float().requires_grad_(True) for h in hs]
    want = OL.layer_wise_mse(OL.get_hidden_states(hs32, mask),
                             [t for t in hp.float().split([int(c) for c in counts], dim=1)])
    wg = torch.autograd.grad(want, hs32)
    flat = mask.reshape(-1)
    rank = flat.long().cumsum(0) - 1
    r2c = torch.where(flat, rank, torch.full_like(rank, -1)).to(torch.int32).cuda()
    w_rows = torch.cat([torch.full((c,), 1.0 / (B * L * c * D)) for c in counts]).cuda()
    hs_c = [h.cuda().requires_grad_(True) for h in hs]
    loss = ops.align_mse(hs_c, hp.cuda(), r2c, w_rows)
    grads = torch.autograd.grad(loss, hs_c)
    assert abs(float(loss) - float(want)) <= 2e-6 * abs(float(want))
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    for a, b in zip(grads, wg):
        assert rel(a, b) < tol
        if not bool(mask.all()):
            assert float(a[~mask.cuda()].abs().max()) == 0.0      # grad lands only on selected rows
    # folded upstream scale
    hs_c2 = [h.cuda().requires_grad_(True) for h in hs]
    loss2 = ops.align_mse(hs_c2, hp.cuda(), r2c, w_rows, grad_scale=0.25, assume_folded=True)
    g2 = torch.autograd.grad(loss2 * 0.25, hs_c2)
    for a, b in zip(g2, wg):
        assert rel(a, 0.25 * b) < tol


def test_align_mse_reference_golden():
    """The exact ragged case the reference produced (tests/golden/loss_pieces.npz)."""
    from mimic_b200 import ops
    z = np.load(os.path.join(GOLDEN, "loss_pieces.npz"))
    qry = torch.from_numpy(z["qry"]).cuda()                    # [L,B,T,D]
    mask_q = torch.from_numpy(z["mask_q"]).bool()
    hp = torch.cat([torch.from_numpy(z["hp0"]), torch.from_numpy(z["hp1"])], dim=1).cuda()
    L, B, T, D = qry.shape
    counts = mask_q.sum(1).tolist()
    flat = mask_q.reshape(-1)
    rank = flat.long().cumsum(0) - 1
    r2c = torch.where(flat, rank, torch.full_like(rank, -1)).to(torch.int32).cuda()
    w_rows = torch.cat([torch.full((c,), 1.0 / (B * L * c * D)) for c in counts]).cuda()
    hs = [qry[l].clone().requires_grad_(True) for l in range(L)]
    loss = ops.align_mse(hs, hp, r2c, w_rows)
    grads = torch.stack(torch.autograd.grad(loss, hs))
    assert abs(float(loss) - float(z["ffn_mse_loss"])) < 2e-6 * float(z["ffn_mse_loss"])
    assert rel(grads, torch.from_numpy(z["dqry"])) < 1e-5


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
@pytest.mark.parametrize("shape", [(3, 2, 8, 64, (5, 6)), (4, 2, 40, 1024, (33, 17)), (2, 1, 16, 4096, (16,)),
                                   (2, 1, 6, 8192, (4,))])
def test_align_cos_matches_oracle(dtype, shape):
    """Strategy.LAYER_WISE_COS_SIM kernel vs the oracle restatement of src/shift_model.py:152-160."""
    from mimic_b200 import ops
    L, B, T, D, counts = shape
    hs, mask, hp = _mse_case(dtype, L, B, T, D, counts, seed=7)
    hs32 = [h.float().requires_grad_(True) for h in hs]
    want = OL.layer_wise_cos_sim(OL.get_hidden_states(hs32, mask),
                                 [t for t in hp.float().split([int(c) for c in counts], dim=1)])
    wg = torch.autograd.grad(want, hs32)
    flat = mask.reshape(-1)
    rank = flat.long().cumsum(0) - 1
    r2c = torch.where(flat, rank, torch.full_like(rank, -1)).to(torch.int32).cuda()
    w_rows = torch.cat([torch.full((c,), 1.0 / (B * L * c)) for c in counts]).cuda()
    hs_c = [h.cuda().requires_grad_(True) for h in hs]
    loss = ops.align_cos(hs_c, hp.cuda(), r2c, w_rows)
    grads = torch.autograd.grad(loss, hs_c)
    assert abs(float(loss) - float(want)) <= 1e-5 * abs(float(want))
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    for a, b in zip(grads, wg):
        assert rel(a, b) < tol
        if not bool(mask.all()):
            assert float(a[~mask.cuda()].abs().max()) == 0.0
    with pytest.raises(Exception):                             # one row must fit the CTA's registers
        big = [torch.zeros(1, 2, 16384, device="cuda")]
        ops.align_cos(big, torch.zeros(1, 2, 16384, device="cuda"), torch.arange(2, dtype=torch.int32, device="cuda"),
                      torch.ones(2, device="cuda"))


def test_align_cos_and_kl_reference_golden():
    """The values and gradients the reference itself produced (tests/golden/loss_pieces.npz)."""
    from mimic_b200 import ops
    z = np.load(os.path.join(GOLDEN, "loss_pieces.npz"))
    qry = torch.from_numpy(z["qry"]).cuda()
    mask_q = torch.from_numpy(z["mask_q"]).bool()
    hp = torch.cat([torch.from_numpy(z["hp0"]), torch.from_numpy(z["hp1"])], dim=1).cuda()
    L, B, T, D = qry.shape
    counts = mask_q.sum(1).tolist()
    flat = mask_q.reshape(-1)
    rank = flat.long().cumsum(0) - 1
    r2c = torch.where(flat, rank, torch.full_like(rank, -1)).to(torch.int32).cuda()
    w_rows = torch.cat([torch.full((c,), 1.0 / (B * L * c)) for c in counts]).cuda()
    hs = [qry[l].clone().requires_grad_(True) for l in range(L)]
    loss = ops.align_cos(hs, hp, r2c, w_rows)
    assert abs(float(loss) - float(z["ffn_cos_sim"])) < 2e-6
    assert rel(torch.stack(torch.autograd.grad(loss, hs)), torch.from_numpy(z["dqry_cos"])) < 1e-5
    # KL, all rows selected
    a = torch.from_numpy(z["kl_a"]).cuda().requires_grad_(True)
    rows = torch.arange(a.shape[0], dtype=torch.int32, device="cuda")
    kl = ops.kl_loss(a, torch.from_numpy(z["kl_b"]).cuda(), rows)
    assert abs(float(kl) - float(z["kl"])) < 2e-6
    assert rel(torch.autograd.grad(kl, a)[0], torch.from_numpy(z["dkl_a"])) < 1e-5
    # KL, ragged pairing built the way ShiftModel does it (no masked copies)
    from mimic_b200.shift_model import _paired_rows
    a2 = torch.from_numpy(z["kl2_a"]).cuda().requires_grad_(True)
    pr = _paired_rows(torch.from_numpy(z["kl2_ma"]).cuda().bool()[None], torch.from_numpy(z["kl2_mb"]).cuda().bool()[None])
    kl2 = ops.kl_loss(a2, torch.from_numpy(z["kl2_b"]).cuda(), pr)
    g2, = torch.autograd.grad(kl2, a2)
    assert abs(float(kl2) - float(z["kl2"])) < 2e-6
    assert rel(g2, torch.from_numpy(z["dkl2_a"])) < 1e-5
    assert float(g2[~torch.from_numpy(z["kl2_ma"]).cuda().bool()].abs().max()) == 0.0


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
@pytest.mark.parametrize("R,V", [(3, 5), (7, 50), (37, 32003), (3, 60001), (3, 151655), (300, 513)])
@pytest.mark.parametrize("inplace", [False, True])
def test_kl_matches_oracle(dtype, R, V, inplace):
    """Strategy.LOGITS_KL_DIV kernel vs F.kl_div (src/shift_model.py:181-191), odd vocabularies, unselected rows."""
    from mimic_b200 import ops
    g = torch.Generator().manual_seed(R * 7 + V)
    zl = (torch.randn(R, V, generator=g) * 3).to(dtype)
    yl = (torch.randn(R + 5, V, generator=g) * 3).to(dtype)
    sel = torch.ones(R, dtype=torch.bool)
    sel[::4] = False
    y_row = torch.where(sel, torch.randperm(R + 5, generator=g)[:R], torch.full((R,), -1)).to(torch.int32)
    z32 = zl.float().requires_grad_(True)
    want = OL.logits_kl(z32[sel], yl.float()[y_row[sel].long()])
    wg, = torch.autograd.grad(want, z32)
    zc = zl.cuda().requires_grad_(True)
    x = zc * 1.0 if inplace else zc
    loss = ops.kl_loss(x, yl.cuda(), y_row.cuda(), inplace=inplace)
    got, = torch.autograd.grad(loss, zc)
    assert abs(float(loss) - float(want)) < 1e-4 * abs(float(want)) + 1e-6
    assert rel(got, wg) < (1e-5 if dtype == torch.float32 else 1e-2)
    assert float(got[~sel.cuda()].abs().max()) == 0.0


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
@pytest.mark.parametrize("R,V", [(3, 5), (7, 50), (37, 32003), (3, 60001), (5, 151655), (300, 513)])
@pytest.mark.parametrize("inplace", [False, True])
def test_ce_matches_oracle(dtype, R, V, inplace):
    from mimic_b200 import ops
    g = torch.Generator().manual_seed(R + V)
    logits = (torch.randn(R, V, generator=g) * 3).to(dtype)
    labels = torch.randint(0, V, (R,), generator=g)
    labels[::3] = -100
    l32 = logits.float().requires_grad_(True)
    want = F.cross_entropy(l32, labels, ignore_index=-100)
    wg, = torch.autograd.grad(want, l32)
    lc = logits.cuda().requires_grad_(True)
    x = lc * 1.0 if inplace else lc          # in-place needs a non-leaf
    loss = ops.ce_loss(x, labels.cuda(), inplace=inplace)
    got, = torch.autograd.grad(loss, lc)
    assert abs(float(loss) - float(want)) < 1e-5 * abs(float(want)) + 1e-6
    assert rel(got, wg) < (1e-5 if dtype == torch.float32 else 1e-2)
    assert float(got[labels.cuda() < 0].abs().max()) == 0.0


def test_ce_all_ignored_and_lm_rule():
    from mimic_b200 import ops
    logits = torch.randn(4, 33, device="cuda", requires_grad=True)
    loss = ops.ce_loss(logits, torch.full((4,), -100, device="cuda"))
    assert float(loss) == 0.0
    # the LM rule of oracle.losses.lm_ce_loss, built the way ShiftModel.forward_inputs builds labels
    torch.manual_seed(3)
    B, T, V = 2, 9, 17
    lg = torch.randn(B, T, V)
    ids = torch.randint(1, V, (B, T))
    am = torch.ones(B, T, dtype=torch.bool)
    am[0, :3] = False
    am[1, 5] = False
    want = OL.lm_ce_loss(lg, ids, am)
    labels = torch.full_like(ids, -100)
    labels[:, :-1] = torch.where(am[:, 1:], ids[:, 1:], torch.full_like(ids[:, 1:], -100))
    got = ops.ce_loss(lg.cuda(), labels.cuda())
    assert abs(float(got) - float(want)) < 1e-5


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
@pytest.mark.parametrize("B,T,H,Hkv,d,bcast", [(2, 9, 8, 2, 128, False), (1, 33, 4, 4, 64, True), (3, 5, 32, 8, 128, False)])
def test_rope_qk_matches_hf_formula(dtype, B, T, H, Hkv, d, bcast):
    """One-launch RoPE vs HF apply_rotary_pos_emb (x*cos + rotate_half(x)*sin), strided projection-output views,
    forward and backward (the transposed rotation)."""
    from mimic_b200 import ops
    g = torch.Generator().manual_seed(B * 100 + T)
    qkv = torch.randn(B, T, (H + 2 * Hkv) * d, generator=g).to(dtype)      # a fused projection: q and k are strided views
    q = qkv[..., : H * d].view(B, T, H, d)
    k = qkv[..., H * d: (H + Hkv) * d].view(B, T, Hkv, d)
    pos = torch.arange(T)[None].expand(1 if bcast else B, T) + (0 if bcast else torch.arange(B)[:, None] * 3)
    inv = 1.0 / (10000.0 ** (torch.arange(0, d, 2).float() / d))
    ang = pos[..., None].float() * inv
    emb = torch.cat((ang, ang), -1)
    cos, sin = emb.cos().to(dtype), emb.sin().to(dtype)

    def rot(x):
        return torch.cat((-x[..., d // 2:], x[..., : d // 2]), -1)

    q32, k32 = q.float().requires_grad_(True), k.float().requires_grad_(True)
    c32, s32 = cos.float()[:, :, None], sin.float()[:, :, None]
    wq, wk = q32 * c32 + rot(q32) * s32, k32 * c32 + rot(k32) * s32
    gq, gk = torch.randn(B, T, H, d, generator=g), torch.randn(B, T, Hkv, d, generator=g)
    wdq, wdk = torch.autograd.grad([wq, wk], [q32, k32], [gq, gk])
    qc, kc = q.cuda().requires_grad_(True), k.cuda().requires_grad_(True)
    oq, ok = ops.rope_qk(qc, kc, cos.cuda(), sin.cuda())
    assert oq.shape == (B, T, H, d) and ok.shape == (B, T, Hkv, d) and oq.is_contiguous()
    tol = 1e-6 if dtype == torch.float32 else 6e-3
    assert rel(oq, wq) < tol and rel(ok, wk) < tol
    dq, dk = torch.autograd.grad([oq, ok], [qc, kc], [gq.cuda().to(dtype), gk.cuda().to(dtype)])
    assert rel(dq, wdq) < tol and rel(dk, wdk) < tol


def test_adamw_step_matches_torch():
    from mimic_b200 import ops
    torch.manual_seed(0)
    n = 10007
    p0 = torch.randn(n)
    wd = torch.where(torch.arange(n) % 5 == 0, torch.zeros(n), torch.full((n,), 1e-3))
    pa = torch.nn.Parameter(p0[wd == 0].clone())
    pb = torch.nn.Parameter(p0[wd != 0].clone())
    opt = torch.optim.AdamW([{"params": [pa], "weight_decay": 0.0}, {"params": [pb], "weight_decay": 1e-3}], lr=5e-3)
    p = p0.clone().cuda()
    m, v = torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    for step in range(1, 4):
        g = torch.randn(n) * 3
        pa.grad, pb.grad = g[wd == 0].clone(), g[wd != 0].clone()
        torch.nn.utils.clip_grad_norm_([pa, pb], 1.0)
        opt.step()
        ops.adamw_step(p, g.cuda(), m, v, wd.cuda(), ops.adamw_hyper(5e-3, step).cuda(), max_norm=1.0)
        want = torch.empty(n)
        want[wd == 0], want[wd != 0] = pa.detach(), pb.detach()
        assert rel(p, want) < 1e-6


def test_adamw_scale_lr_group_matches_torch():
    """cfg.peft.scale_lr (src/shift_model.py:341-355): parameters named *log_Z1* / *scale* train at a second learning
    rate; four torch param groups (bias / non-bias x regular / scale) against the flat buffer with per-element vectors."""
    from mimic_b200 import ops
    torch.manual_seed(1)
    n, lr, scale_lr = 4099, 5e-3, 1e-4
    p0 = torch.randn(n)
    idx = torch.arange(n)
    no_decay, scaled = idx % 5 == 0, idx % 3 == 0
    groups, masks = [], []
    for nd in (True, False):
        for sc in (True, False):
            mk = (no_decay == nd) & (scaled == sc)
            masks.append(mk)
            groups.append({"params": [torch.nn.Parameter(p0[mk].clone())], "weight_decay": 0.0 if nd else 1e-3,
                           "lr": scale_lr if sc else lr})
    opt = torch.optim.AdamW(groups, lr=lr)
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda s: 0.5 ** s)       # one lambda for every group, like the reference
    wd = torch.where(no_decay, torch.zeros(n), torch.full((n,), 1e-3)).cuda()
    lr_mul = torch.where(scaled, torch.full((n,), scale_lr / lr), torch.ones(n)).cuda()
    p = p0.clone().cuda()
    m, v = torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    for step in range(1, 4):
        g = torch.randn(n) * 3
        params = [gr["params"][0] for gr in groups]
        for prm, mk in zip(params, masks):
            prm.grad = g[mk].clone()
        torch.nn.utils.clip_grad_norm_(params, 1.0)
        opt.step()
        ops.adamw_step(p, g.cuda(), m, v, wd, ops.adamw_hyper(lr * 0.5 ** (step - 1), step).cuda(), max_norm=1.0,
                       lr_mul=lr_mul)
        sched.step()
        want = torch.empty(n)
        for prm, mk in zip(params, masks):
            want[mk] = prm.detach()
        assert rel(p, want) < 1e-6


@pytest.mark.parametrize("V", [1000, 32003, 60001], ids=["V1000", "V32003", "V60001"])
def test_ce_kl_strided_offset_rows_many_rows_per_cluster(V):
    """The register / cluster kernels stage every row with a bulk copy whose source must be 16-byte aligned: rows that
    are views into a wider buffer (row stride > V, odd element offset of the first row), more rows than resident clusters
    (every cluster walks several rows: stage re-arming, barrier phases), ignored rows in a row (nothing staged for them)."""
    from mimic_b200 import ops
    g = torch.Generator().manual_seed(V)
    R, pad = (700 if V <= 1000 else 150), 13
    wide = (torch.randn(R + 1, V + pad, generator=g) * 3).bfloat16()
    wide_y = (torch.randn(R + 1, V + 2 * pad + 1, generator=g) * 3).bfloat16()
    labels = torch.randint(0, V, (R,), generator=g)
    labels[5:40] = -100                                    # a run of ignored rows, longer than one cluster's stride can skip
    labels[::9] = -100
    wide_c, wide_yc = wide.cuda(), wide_y.cuda()
    z = wide_c[1:, 3:3 + V]                                # row stride V + 13, first element at an odd offset
    y = wide_yc[:-1, 8:8 + V]
    assert z.stride(0) == V + pad and z.stride(1) == 1
    # CE
    l32 = wide[1:, 3:3 + V].float().requires_grad_(True)
    want = F.cross_entropy(l32, labels, ignore_index=-100)
    wg, = torch.autograd.grad(want, l32)
    zc = z.detach().requires_grad_(True)
    loss = ops.ce_loss(zc, labels.cuda())
    got, = torch.autograd.grad(loss, zc)
    assert abs(float(loss) - float(want)) < 1e-5 * abs(float(want)) + 1e-6
    assert rel(got, wg) < 1e-2 and float(got[labels.cuda() < 0].abs().max()) == 0.0
    # KL against rows of the other buffer
    sel = labels >= 0
    y_row = torch.where(sel, torch.randperm(R, generator=g), torch.full((R,), -1)).to(torch.int32)
    z32 = wide[1:, 3:3 + V].float().requires_grad_(True)
    want = OL.logits_kl(z32[sel], wide_y[:-1, 8:8 + V].float()[y_row[sel].long()])
    wg, = torch.autograd.grad(want, z32)
    zc = z.detach().requires_grad_(True)
    loss = ops.kl_loss(zc, y, y_row.cuda())
    got, = torch.autograd.grad(loss, zc)
    assert abs(float(loss) - float(want)) < 1e-4 * abs(float(want)) + 1e-6
    assert rel(got, wg) < 1e-2 and float(got[~sel.cuda()].abs().max()) == 0.0


def test_clip_norm_is_summed_in_a_fixed_order():
    """The clipped AdamW step is bit-reproducible: the squared gradient norm is reduced per CTA and then in CTA order (a
    float atomicAdd per CTA made the clip factor depend on arrival order — data-parallel ranks with identical all-reduced
    gradients drifted apart by an ulp per step)."""
    from mimic_b200 import ops
    g = torch.Generator().manual_seed(9)
    n = 263168 + 4                                           # the Idefics2 flat buffer (+ scalars): several hundred CTAs
    grad = (torch.randn(n, generator=g) * torch.logspace(-6, 2, n)).cuda()     # wide dynamic range: order matters in fp32
    p0 = torch.randn(n, generator=g).cuda()
    wd = torch.full((n,), 1e-3).cuda()
    hyper = ops.adamw_hyper(5e-3, 1).cuda()
    outs, norms = [], []
    for _ in range(8):
        p, m, v = p0.clone(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
        ws = ops.adamw_workspace(p.device)
        for _ in range(3):                                   # the workspace is reusable (self-resetting counter)
            ops.adamw_step(p, grad.clone(), m, v, wd, hyper, max_norm=1.0, sumsq=ws)
        outs.append(p)
        norms.append(float(ws[0]))
    assert all(torch.equal(outs[0], o) for o in outs) and len(set(norms)) == 1
    want = float((grad.double() ** 2).sum())
    assert abs(norms[0] - want) < 1e-5 * want
