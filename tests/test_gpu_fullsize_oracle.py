"""tcgen05 kernels at BASELINE's FULL sizes against the ORACLE (not against the repo's own SIMT kernel): the kernels run
the whole shape (their real grid, work lists and L2 behaviour), the fp32 oracle restates a SLICE of it — one or two
heads — which it finishes in seconds (two heads of T = 4096 are two 67 MB score matrices).  With H == Hkv every head's
dq / dk / dv depends on that head alone, and with the whole batch in the slice so do dshift[h], dw[h], db[h].
Tolerance: north-star 1e-2 rel-L2 for bf16 I/O, on outputs AND gradients."""
import pytest
import torch

from oracle import shifted_attention as OA

pytestmark = pytest.mark.gpu
TOL = 1e-2


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def _rand(shape, g, scale=1.0):
    return (torch.randn(*shape, device="cuda", generator=g) * scale)


def test_context_pass_full_shape_vs_oracle_slice():
    """BASELINE config 2, context pass: B=2, H=32, Hkv=8, S=3200, d=128, causal + left padding + one PAD separator hole,
    inactive handle.  Heads 5 and 31 of the padded batch row against the oracle."""
    from mimic_b200 import ops
    B, H, Hkv, S, d = 2, 32, 8, 3200, 128
    g = torch.Generator(device="cuda").manual_seed(21)
    q, k, v = _rand((B, H, S, d), g).bfloat16(), _rand((B, Hkv, S, d), g).bfloat16(), _rand((B, Hkv, S, d), g).bfloat16()
    km = torch.ones(B, S, dtype=torch.bool, device="cuda")
    km[1, :11] = False
    km[1, 1700] = False
    out, lse_c, _, _, _ = ops.shifted_attention_forward(q, k, v, km, None, None, None, causal=True, active=False,
                                                        scale=d ** -0.5, impl="tcgen05")
    heads = [5, 31]
    kvh = [h // (H // Hkv) for h in heads]
    want = OA.shifted_attention_fwd(q[1:2, heads].float().cpu(), k[1:2, kvh].float().cpu(), v[1:2, kvh].float().cpu(),
                                    causal=True, key_mask=km[1:2].cpu(), active=False)
    valid = torch.ones(S, dtype=torch.bool)
    valid[:11] = False                                     # padded query rows see no key: SA defined as 0, lse -inf
    assert rel(out[1, :, heads][valid], want["out"][0][valid]) < TOL
    assert float(out[1, :11].float().abs().max()) == 0.0
    assert rel(lse_c[1, heads][:, valid], want["lse_c"][0][:, valid]) < 1e-4


@pytest.mark.parametrize("T,B,heads", [(4096, 4, [0, 17]), (8192, 2, [9])], ids=["T4096", "T8192"])
def test_sweep_shape_active_fwd_bwd_vs_oracle_slice(T, B, heads):
    """BASELINE config 5 (T = S, H = Hkv = 32, d = 128, active, causal): whole-batch slices of one or two heads, forward and
    backward, every output of the op."""
    from mimic_b200 import ops
    H, d = 32, 128
    g = torch.Generator(device="cuda").manual_seed(3407 + T)
    q, k, v = (_rand((B, H, T, d), g).bfloat16() for _ in range(3))
    do = _rand((B, T, H, d), g).bfloat16()
    sh, w, b = _rand((H, d), g), _rand((H, d), g, 0.1), _rand((H,), g)       # trained-like: mu is not ~0
    kw = dict(causal=True, active=True, scale=d ** -0.5, impl="tcgen05")
    out, lse_c, lse_f, mu, saved = ops.shifted_attention_forward(q, k, v, None, sh, w, b, **kw)
    dq, dk, dv, dsh, dw, db = ops.shifted_attention_backward(saved, out, do, lse_c, lse_f, mu, **kw)
    c = lambda t: t.float().cpu()
    for h in heads:
        args = (c(q[:, h:h + 1]), c(k[:, h:h + 1]), c(v[:, h:h + 1]), c(sh[h:h + 1]), c(w[h:h + 1]), c(b[h:h + 1]))
        want = OA.shifted_attention_fwd(*args, causal=True)
        assert rel(out[:, :, h], want["out"][:, :, 0]) < TOL, f"out h={h}"
        assert rel(lse_f[:, h], want["lse_full"][:, 0]) < 1e-4, f"lse_full h={h}"
        assert rel(mu[:, h], want["mu"][:, 0]) < 2e-3, f"mu h={h}"
        wg = OA.shifted_attention_bwd(*args, c(do[:, :, h:h + 1]), causal=True)
        for name, got in (("dq", dq[:, h]), ("dk", dk[:, h]), ("dv", dv[:, h])):
            assert rel(got, wg[name][:, 0]) < TOL, f"{name} h={h}"
        for name, got in (("dshift", dsh[h]), ("dw", dw[h]), ("db", db[h])):
            assert rel(got, wg[name][0]) < TOL, f"{name} h={h}"


def test_query_pass_shape_gqa_vs_oracle():
    """BASELINE config 2, shifted query pass: B=2, H=32, Hkv=8, T=S=128 with left padding — the whole op."""
    from mimic_b200 import ops
    B, H, Hkv, T, d = 2, 32, 8, 128, 128
    g = torch.Generator(device="cuda").manual_seed(5)
    q, k, v = _rand((B, H, T, d), g).bfloat16(), _rand((B, Hkv, T, d), g).bfloat16(), _rand((B, Hkv, T, d), g).bfloat16()
    km = torch.ones(B, T, dtype=torch.bool, device="cuda")
    km[1, :9] = False
    do = _rand((B, T, H, d), g).bfloat16()
    do[1, :9] = 0
    sh, w, b = _rand((H, d), g, 0.5), _rand((H, d), g, 0.1), _rand((H,), g)
    kw = dict(causal=True, active=True, scale=d ** -0.5, impl="tcgen05")
    out, lse_c, lse_f, mu, saved = ops.shifted_attention_forward(q, k, v, km, sh, w, b, **kw)
    got = ops.shifted_attention_backward(saved, out, do, lse_c, lse_f, mu, **kw)
    c = lambda t: t.float().cpu()
    args = (c(q), c(k), c(v), c(sh), c(w), c(b))
    want = OA.shifted_attention_fwd(*args, causal=True, key_mask=km.cpu())
    wg = OA.shifted_attention_bwd(*args, c(do), causal=True, key_mask=km.cpu())
    assert rel(out, want["out"]) < TOL
    for name, gt in zip(("dq", "dk", "dv", "dshift", "dw", "db"), got):
        assert rel(gt, wg[name]) < TOL, name
