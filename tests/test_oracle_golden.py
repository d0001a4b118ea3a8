"""Pin the oracle: every fixture under tests/golden/ was produced by RUNNING THE REFERENCE
(oracle/gen_golden.py); the oracle's restatement must reproduce it.  CPU only."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import losses as OL
from oracle import shifted_attention as OA

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ATTN_FILES = sorted(glob.glob(os.path.join(GOLDEN, "attn_*.npz")))


def rel(a, b):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def load(path):
    z = np.load(path)
    return {k: torch.from_numpy(z[k]) for k in z.files}


def test_fixtures_present():
    assert len(ATTN_FILES) >= 10
    assert os.path.isfile(os.path.join(GOLDEN, "loss_pieces.npz"))
    assert os.path.isfile(os.path.join(GOLDEN, "tiny_step.npz"))


@pytest.mark.parametrize("path", ATTN_FILES, ids=[os.path.basename(p)[5:-4] for p in ATTN_FILES])
def test_attention_oracle_matches_reference(path):
    g = load(path)
    B, H, Hkv, T, S, d, causal, active, multi_head = [int(x) for x in g["meta"]]
    km = g["key_mask"].bool()
    rv = g["row_valid"].bool()
    if not multi_head:
        # single-head mu variant (src/shift_encoder.py:689-692, 705-707): check do_shift restatement
        kk = OA.repeat_kv(g["k"], H // Hkv)
        vv = OA.repeat_kv(g["v"], H // Hkv)
        sa, _ = OA.masked_attention(g["q"], kk, vv, OA.visibility(T, S, B, bool(causal), km), d ** -0.5)
        out = OA.do_shift(g["q"], kk, sa.transpose(1, 2), g["shift"], g["w"], g["b"], multi_head=False)
        assert rel(out.reshape(B, T, H, d), g["out"]) < 2e-6
        return
    shift = g["shift"].view(H, d)
    kw = dict(causal=bool(causal), key_mask=km, active=bool(active))
    fwd = OA.shifted_attention_fwd(g["q"], g["k"], g["v"], shift, g["w"], g["b"], **kw)
    m = rv[:, :, None, None].expand_as(g["out"])
    assert rel(fwd["out"][m], g["out"][m]) < 2e-6
    # rows with no visible key are defined as SA = 0 by the oracle
    if (~rv).any():
        assert float(fwd["sa"][~rv].abs().max()) == 0.0
    bwd = OA.shifted_attention_bwd(g["q"], g["k"], g["v"], shift, g["w"], g["b"], g["d_out"], **kw)
    for name in ("dq", "dk", "dv"):
        assert rel(bwd[name], g[name]) < 5e-6, name
    if active:
        assert rel(bwd["dshift"].reshape(-1), g["dshift"]) < 5e-6
        assert rel(bwd["dw"], g["dw"]) < 5e-6
        assert rel(bwd["db"], g["db"]) < 5e-6
    # explicit backward == autograd of the forward restatement (fp64)
    dd = {k: v.double() for k, v in g.items() if v.dtype == torch.float32}
    ag, _ = OA.shifted_attention_autograd(dd["q"], dd["k"], dd["v"], dd["shift"].view(H, d), dd["w"], dd["b"],
                                          dd["d_out"], **kw)
    ex = OA.shifted_attention_bwd(dd["q"], dd["k"], dd["v"], dd["shift"].view(H, d), dd["w"], dd["b"],
                                  dd["d_out"], **kw)
    names = ("dq", "dk", "dv") + (("dshift", "dw", "db") if active else ())
    for name in names:
        assert rel(ex[name], ag[name]) < 1e-12, name


def test_virtual_key_identity():
    """mu == softmax([S_full, z1])[-1]  (SURVEY.md §0, fact 2)."""
    torch.manual_seed(3407)
    B, H, T, d = 2, 4, 32, 16
    q, k, v = (torch.randn(B, H, T, d, dtype=torch.float64) for _ in range(3))
    w, b, sh = torch.randn(H, d, dtype=torch.float64), torch.randn(H, dtype=torch.float64), torch.randn(H, d, dtype=torch.float64)
    f = OA.shifted_attention_fwd(q, k, v, sh, w, b)
    s = q @ k.transpose(-2, -1) / d ** 0.5
    z1 = torch.einsum("bhtd,hd->bht", q, w) + b[None, :, None]
    p = torch.softmax(torch.cat([s, z1[..., None]], -1), -1)[..., -1]
    assert float((p - f["mu"]).abs().max()) < 1e-12
    assert float(f["mu"].min()) > 0 and float(f["mu"].max()) < 1


def test_future_and_padded_keys_change_mu_not_sa():
    """The unmasked-Z2 rule (src/shift_encoder.py:697-703)."""
    torch.manual_seed(0)
    B, H, T, d = 1, 2, 16, 8
    q, k, v = (torch.randn(B, H, T, d) for _ in range(3))
    w, b, sh = torch.randn(H, d), torch.randn(H), torch.randn(H, d)
    km = torch.ones(B, T, dtype=torch.bool)
    km[0, 3] = False
    f0 = OA.shifted_attention_fwd(q, k, v, sh, w, b, key_mask=km)
    k2 = k.clone()
    k2[:, :, -1] += 3.0          # a FUTURE key for every row but the last
    k2[:, :, 3] -= 2.0           # a PADDED key
    f1 = OA.shifted_attention_fwd(q, k2, v, sh, w, b, key_mask=km)
    assert torch.equal(f0["sa"][:, : T - 1], f1["sa"][:, : T - 1])
    assert float((f0["mu"][..., : T - 1] - f1["mu"][..., : T - 1]).abs().max()) > 1e-4


def test_loss_pieces_match_reference():
    g = load(os.path.join(GOLDEN, "loss_pieces.npz"))
    ids = g["ids"]
    for n in (1, 2):
        assert torch.equal(OL.generate_label_mask(ids, 0, n), g[f"label_mask_sep{n}"].bool())
    assert torch.equal(OL.generate_label_mask(ids, 0, 1, keep_bos=True), g["label_mask_sep1_keep_bos"].bool())
    # SURVEY.md §8c hand example
    assert OL.generate_label_mask(ids, 0, 1)[0].nonzero().flatten().tolist() == [6, 7, 9, 10, 11]
    assert OL.generate_label_mask(ids, 0, 2)[0].nonzero().flatten().tolist() == [9, 10, 11]
    with pytest.raises(ValueError):
        OL.generate_label_mask(ids, 0, 3)
    hp = OL.get_hidden_states(list(g["ctx"]), g["mask_ctx"].bool())
    assert torch.equal(hp[0], g["hp0"]) and torch.equal(hp[1], g["hp1"])
    qry = [t.clone().requires_grad_(True) for t in g["qry"]]
    hq = OL.get_hidden_states(qry, g["mask_q"].bool())
    loss = OL.layer_wise_mse(hq, hp)
    assert abs(float(loss.detach()) - float(g["ffn_mse_loss"])) < 1e-6 * abs(float(g["ffn_mse_loss"]))
    grads = torch.autograd.grad(loss, qry, retain_graph=True)
    assert rel(torch.stack(grads), g["dqry"]) < 1e-6
    cos = OL.layer_wise_cos_sim(hq, hp)
    assert abs(float(cos) - float(g["ffn_cos_sim"])) < 1e-6
    assert rel(torch.stack(torch.autograd.grad(cos, qry)), g["dqry_cos"]) < 1e-6
    ka = g["kl_a"].clone().requires_grad_(True)
    kl = OL.logits_kl(ka, g["kl_b"])
    assert abs(float(kl) - float(g["kl"])) < 1e-6
    assert rel(torch.autograd.grad(kl, ka)[0], g["dkl_a"]) < 1e-6
    # ragged pairing: k-th selected query row <-> k-th selected context row (src/shift_model.py:185-186)
    k2 = g["kl2_a"].clone().requires_grad_(True)
    kl2 = OL.logits_kl(k2[g["kl2_ma"].bool()], g["kl2_b"][g["kl2_mb"].bool()])
    assert abs(float(kl2) - float(g["kl2"])) < 1e-6
    assert rel(torch.autograd.grad(kl2, k2)[0], g["dkl2_a"]) < 1e-6
    # manual formula: mean_i mean_{l,t,d}
    manual = sum(float(((a.detach() - b) ** 2).mean()) for a, b in zip(hq, hp)) / len(hp)
    assert abs(manual - float(g["ffn_mse_loss"])) < 1e-6


def test_lm_ce_rule():
    torch.manual_seed(1)
    B, T, V = 2, 9, 17
    logits = torch.randn(B, T, V)
    ids = torch.randint(1, V, (B, T))
    am = torch.ones(B, T, dtype=torch.bool)
    am[0, :3] = False
    am[1, 5] = False
    got = OL.lm_ce_loss(logits, ids, am)
    rows, tgt = [], []
    for b in range(B):
        for t in range(T - 1):
            if am[b, t + 1]:
                rows.append(logits[b, t])
                tgt.append(ids[b, t + 1])
    want = torch.nn.functional.cross_entropy(torch.stack(rows), torch.stack(tgt))
    assert abs(float(got) - float(want)) < 1e-6


# ---------------------------------------------------------------- other encoder variants
VARIANTS = [("sh_pad", False, True), ("sh_cache", False, True), ("vec_multi", True, False), ("vec_flat", False, False)]


def _variant(prefix):
    z = np.load(os.path.join(GOLDEN, "variants.npz"))
    return {k[len(prefix) + 1:]: torch.from_numpy(z[k]) for k in z.files if k.startswith(prefix + ".")}


@pytest.mark.parametrize("key,multi_head,learnable", VARIANTS, ids=[v[0] for v in VARIANTS])
def test_do_shift_variants_oracle_matches_reference(key, multi_head, learnable):
    """Head-averaged mu (src/shift_encoder.py:689-692, 705-707) and the bare vector shift (717-729): forward against
    what the reference produced, autograd of the restatement against the reference's gradients."""
    g = _variant(key)
    B, H, Hkv, T, S, d, causal, active, mh = [int(x) for x in g["meta"]]
    assert bool(mh) == multi_head and ("w" in g) == learnable
    km, rv = g["key_mask"].bool(), g["row_valid"].bool()
    q, k, v = (g[n].clone().requires_grad_(True) for n in ("q", "k", "v"))
    shift = g["shift"].clone().requires_grad_(True)
    w = g["w"].clone().requires_grad_(True) if learnable else None
    b = g["b"].clone().requires_grad_(True) if learnable else None
    kk, vv = OA.repeat_kv(k, H // Hkv), OA.repeat_kv(v, H // Hkv)
    sa, _ = OA.masked_attention(q, kk, vv, OA.visibility(T, S, B, bool(causal), km), d ** -0.5)
    sh = shift.view(H, d) if multi_head else shift
    out = OA.do_shift(q, kk, sa.transpose(1, 2), sh, w, b, multi_head=multi_head, learnable_scale=learnable)
    out = out.reshape(B, T, H, d)
    m = rv[:, :, None, None].expand_as(g["out"])
    assert rel(out[m], g["out"][m]) < 2e-6
    leaves = [q, k, v, shift] + ([w, b] if learnable else [])
    names = ["dq", "dk", "dv", "dshift"] + (["dw", "db"] if learnable else [])
    for name, got in zip(names, torch.autograd.grad(out, leaves, g["d_out"])):
        assert rel(got.reshape(g[name].shape), g[name]) < 5e-6, name


@pytest.mark.parametrize("key", ["ns_attn", "ns_ffn", "ns_approx"])
def test_norm_shift_oracle_matches_reference(key):
    """AttnFFNShift._shift_hook (src/shift_encoder.py:233-240) / AttnApproximator's FFN hook (661-667)."""
    from oracle import norm_shift as NS
    g = _variant(key)
    scale = g["scale"] if "scale" in g else torch.tensor(1.0)
    y = NS.norm_shift_fwd(g["h"], g["shift"], scale)
    assert rel(y, g["y"]) < 1e-6
    # norm preserved per token row
    assert torch.allclose(y.norm(dim=-1), g["h"].norm(dim=-1), rtol=1e-5)
    dh, dshift, dscale = NS.norm_shift_bwd(g["h"], g["shift"], scale, g["dy"])
    assert rel(dh, g["dh"]) < 5e-6
    assert rel(dshift, g["dshift"]) < 5e-6
    if "dscale" in g:
        assert rel(dscale, g["dscale"]) < 5e-6
    # closed form == autograd of the forward restatement, fp64
    h, sh, sc = g["h"].double().requires_grad_(True), g["shift"].double().requires_grad_(True), scale.double().requires_grad_(True)
    ag = torch.autograd.grad(NS.norm_shift_fwd(h, sh, sc), [h, sh, sc], g["dy"].double())
    ex = NS.norm_shift_bwd(h.detach(), sh.detach(), sc.detach(), g["dy"].double())
    for a_, e_ in zip(ag, ex):
        assert rel(e_, a_) < 1e-12
