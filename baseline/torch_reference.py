"""The reference's PyTorch path, restated in eager torch for TIMING ON THE SAME GPU (bench.py `sweep.gpu_reference_ms`
and `gpu_reference_step`).  BENCH BASELINE ONLY: nothing in mimic_b200/ imports this file, and it imports nothing from
oracle/ or from the CUDA library.

What the reference runs per decoder layer (src/shift_encoder.py): F.scaled_dot_product_attention (352-359 / 449-456 /
518-532) on repeat_kv'd keys (431-432), then AttnApproximator.do_shift (682-732) — a SECOND QK^T, an unmasked
logsumexp over the materialised [B,H,T,S] scores, MultiheadLinear (556), mu (712), the rank-1 add (724) — all under
Lightning's mixed-precision autocast (config/train.yaml: 16-mixed; bf16 here).  Around it: HF rotary embedding
(elementwise), record hooks that keep the layer outputs, masked_select + mse_loss (src/shift_model.py:121-179),
F.cross_entropy on fp32 logits, torch AdamW + clip_grad_norm_.

`reference_ops(ops)` swaps the C-ABI front-ends of mimic_b200.ops for these eager restatements, so the SAME backbone,
hooks, batches and trainer run "the reference way" on the GPU.  Two things stay favourable to the reference and are
said so in bench.py's output: the record path keeps our row plan (the reference stacks every layer's full [B,S,D]
activations and masked_selects them, src/shift_model.py:128-141), and SDPA may pick its flash backend when no padding
mask is present."""
from __future__ import annotations

import contextlib
import math

import torch
import torch.nn.functional as F


def repeat_kv(x, groups):
    if groups == 1:
        return x
    B, Hkv, S, d = x.shape
    return x[:, :, None].expand(B, Hkv, groups, S, d).reshape(B, Hkv * groups, S, d)


def do_shift(q, k, attn_output, shift, w, b):
    """src/shift_encoder.py:682-732, MULTI_HEAD | LEARNABLE_SHIFT_SCALE | VECTOR_SHIFT.  q [B,H,T,d], k [B,H,S,d],
    attn_output [B,T,H,d]."""
    d = q.shape[-1]
    q_t = q.transpose(1, 2)
    log_z2 = torch.logsumexp(torch.matmul(q, k.transpose(-2, -1)) / (d ** 0.5), dim=-1).transpose(-2, -1)   # 697-703
    log_z1 = torch.einsum("btnd,nd->btn", q_t, w.to(q_t.dtype)) + b                                        # 556, 709
    mu = torch.exp(log_z1 - torch.logaddexp(log_z1.float(), log_z2.float())).unsqueeze(-1)                 # 712-715
    return attn_output + mu * shift[None, None]                                                             # 724


def sdpa_do_shift(q, k, v, shift=None, w=None, bias=None, key_mask=None, *, causal=True, active=True, scale=None,
                  **_ignored):
    """Drop-in for ops.shifted_attention: [B,H,T,d] x [B,Hkv,S,d] -> [B,T,H,d]."""
    B, H, T, d = q.shape
    S = k.shape[2]
    G = H // k.shape[1]
    with torch.autocast("cuda", dtype=torch.bfloat16):
        kk, vv = repeat_kv(k, G), repeat_kv(v, G)
        if key_mask is None:
            a = F.scaled_dot_product_attention(q, kk, vv, is_causal=causal and T > 1, scale=scale)
        else:
            m = key_mask.bool()[:, None, None, :].expand(B, 1, T, S)
            if causal:
                m = m & torch.ones(T, S, dtype=torch.bool, device=q.device).tril(S - T)[None, None]
            # HF's 4-D additive mask: finfo.min at hidden positions (rows without any visible key stay finite)
            add = torch.zeros(B, 1, T, S, dtype=q.dtype, device=q.device).masked_fill_(~m, torch.finfo(q.dtype).min)
            a = F.scaled_dot_product_attention(q, kk, vv, attn_mask=add, scale=scale)
        a = a.transpose(1, 2)
        if active:
            a = do_shift(q, kk, a, shift, w, bias)
    return a.to(q.dtype)


def _rot(x):
    x1, x2 = x[..., : x.shape[-1] // 2], x[..., x.shape[-1] // 2:]
    return torch.cat((-x2, x1), dim=-1)


def rope_qk(q, k, cos, sin):
    """HF apply_rotary_pos_emb on token-major q [B,T,H,d], k [B,T,Hkv,d]."""
    if cos.dim() == 2:
        cos, sin = cos[None], sin[None]
    cos, sin = cos.unsqueeze(2).to(q.dtype), sin.unsqueeze(2).to(q.dtype)
    return q * cos + _rot(q) * sin, k * cos + _rot(k) * sin


def record_gather(src, src_row, dst):
    ok = src_row >= 0                                   # the row plan is padded with -1 up to its capacity
    dst.copy_(src.reshape(-1, src.shape[-1]).index_select(0, src_row.clamp_min(0).long()) * ok[:, None])


def align_mse(hs, hp, row_to_compact, row_w, grad_scale=1.0, assume_folded=False):
    """mean over samples of mse_loss over [L, Nq_i, D] (src/shift_model.py:145-179): stack, boolean select, fp32."""
    sel = row_to_compact >= 0
    H = torch.stack([h.reshape(-1, h.shape[-1]) for h in hs])[:, sel].float()        # [L, n, D]
    n = H.shape[1]
    Hp = hp[:, :n].float()
    w = row_w[:n]
    return (((H - Hp) ** 2).sum(-1) * w[None]).sum()


def ce_loss(logits, labels, grad_scale=1.0, assume_folded=False, inplace=False):
    V = logits.shape[-1]
    return F.cross_entropy(logits.reshape(-1, V).float(), labels.reshape(-1), ignore_index=-100)


def adamw_step(param, grad, exp_avg, exp_avg_sq, weight_decay, hyper, *, betas=(0.9, 0.999), eps=1e-8, grad_mul=1.0,
               max_norm=0.0, sumsq=None, lr_mul=None):
    """clip_grad_norm_ + AdamW, eager elementwise torch on the flat buffers."""
    g = grad * grad_mul
    if max_norm > 0:
        norm = g.norm()
        g = g * torch.clamp(max_norm / (norm + 1e-6), max=1.0)
    lr, bc1, bc2s = hyper[0], hyper[1], hyper[2]
    lr_e = lr if lr_mul is None else lr * lr_mul
    param.mul_(1 - lr_e * weight_decay)
    exp_avg.mul_(betas[0]).add_(g, alpha=1 - betas[0])
    exp_avg_sq.mul_(betas[1]).addcmul_(g, g, value=1 - betas[1])
    denom = exp_avg_sq.sqrt() / bc2s + eps
    param.sub_(lr_e / bc1 * exp_avg / denom)


@contextlib.contextmanager
def reference_ops(ops):
    """Swap mimic_b200.ops' kernel front-ends for the eager restatements above; restores them on exit."""
    names = ("shifted_attention", "rope_qk", "record_gather", "align_mse", "ce_loss", "adamw_step")
    saved = {n: getattr(ops, n) for n in names}
    try:
        for n in names:
            setattr(ops, n, globals()[n if n != "shifted_attention" else "sdpa_do_shift"])
        yield
    finally:
        for n, f in saved.items():
            setattr(ops, n, f)


def attn_fwd_bwd_ms(T, *, active=True, H=32, d=128, total=16384, bwd=True, reps=5, warm=2, seed=3407):
    """SDPA + do_shift forward (and backward) at a BASELINE config-5 shape, CUDA events; None on OOM."""
    dev = "cuda"
    B = max(1, total // T)
    g = torch.Generator(device=dev).manual_seed(seed + T)
    q, k, v = (torch.randn(B, H, T, d, device=dev, generator=g).bfloat16().requires_grad_(bwd) for _ in range(3))
    sh, w = torch.randn(H, d, device=dev, generator=g), torch.randn(H, d, device=dev, generator=g) * 0.1
    b = torch.randn(H, device=dev, generator=g)
    do = torch.randn(B, T, H, d, device=dev, generator=g).bfloat16()

    def fn():
        o = sdpa_do_shift(q, k, v, sh, w, b, None, causal=True, active=active, scale=1.0 / math.sqrt(d))
        if bwd:
            torch.autograd.grad(o, (q, k, v), do)
    try:
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        tot = 0.0
        for _ in range(reps):
            a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            e.record()
            torch.cuda.synchronize()
            tot += a.elapsed_time(e)
        return tot / reps
    except torch.OutOfMemoryError:
        return None
    finally:
        del q, k, v
        torch.cuda.empty_cache()
