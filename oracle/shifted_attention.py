"""Oracle: MimIC shifted self-attention (CPU, torch).  TEST INFRASTRUCTURE ONLY.

Restates, formula by formula,

  * the SDPA call of the three reference attention forwards
    (src/shift_encoder.py:352-359, 449-456, 518-532): causal / key-range masked
    softmax attention, and
  * `AttnApproximator.do_shift` (src/shift_encoder.py:682-732) for the
    `VECTOR_SHIFT | LEARNABLE_SHIFT_SCALE | MULTI_HEAD` strategy of
    `config/encoder/mimic.yaml`, plus the non-MULTI_HEAD variant (689-692,
    705-707).

and the explicit backward used by the CUDA kernels,
which the tests cross-check against torch autograd of the forward restatement.

Visibility rule for key s seen from query t (batch b), T queries, S keys:

    visible(b,t,s) = (not causal or s <= t + (S - T)) and key_mask[b,s]

(`key_mask` is the HF 2-D attention mask: False at padding AND at the PAD separators that
ShiftModel.forward puts between demonstrations, query and answer, src/shift_model.py:204-222.)

A query row with no visible key gives SA = 0 (the reference's behaviour there is
HF-version dependent and such rows never reach a loss).

`log_Z2` is a log-sum-exp over ALL keys: no causal mask, no padding mask
(src/shift_encoder.py:697-703).
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch


def visibility(T: int, S: int, B: int, causal: bool, key_mask: Optional[torch.Tensor], device=None) -> torch.Tensor:
    """Boolean [B,1,T,S] mask of visible keys (True = attend)."""
    if device is None and key_mask is not None:
        device = key_mask.device
    s = torch.arange(S, device=device)[None, None, None, :]
    t = torch.arange(T, device=device)[None, None, :, None]
    vis = torch.ones(B, 1, T, S, dtype=torch.bool, device=device)
    if causal:
        vis = vis & (s <= t + (S - T))
    if key_mask is not None:
        vis = vis & key_mask.bool().view(B, 1, 1, S)
    return vis


def repeat_kv(x: torch.Tensor, groups: int) -> torch.Tensor:
    """[B,Hkv,S,d] -> [B,Hkv*groups,S,d]; HF `repeat_kv`, called at src/shift_encoder.py:431-432."""
    if groups == 1:
        return x
    B, Hkv, S, d = x.shape
    return x[:, :, None].expand(B, Hkv, groups, S, d).reshape(B, Hkv * groups, S, d)


def masked_attention(q, k, v, vis, scale) -> Tuple[torch.Tensor, torch.Tensor]:
    """softmax(QK^T*scale + mask) V  (SDPA semantics, src/shift_encoder.py:352-359).

    q [B,H,T,d], k/v [B,H,S,d], vis [B,1,T,S] bool.  Returns (SA [B,H,T,d], lse_c [B,H,T]).
    Rows with no visible key: SA = 0, lse_c = -inf.
    """
    s = torch.matmul(q, k.transpose(-2, -1)) * scale
    s = s.masked_fill(~vis, float("-inf"))
    lse = torch.logsumexp(s, dim=-1)
    p = torch.exp(s - torch.where(torch.isinf(lse), torch.zeros_like(lse), lse)[..., None])
    p = torch.where(vis, p, torch.zeros_like(p))
    return torch.matmul(p, v), lse


def do_shift(q, k, attn_output, attn_shift_l, w_l, b_l, head_dim: Optional[int] = None,
             multi_head: bool = True, active: bool = True, learnable_scale: bool = True):
    """Restatement of AttnApproximator.do_shift (src/shift_encoder.py:682-732).

    q [B,H,T,d], k [B,H,S,d] (post-RoPE, post-repeat_kv), attn_output [B,T,H,d].
    MULTI_HEAD: attn_shift_l [H,d], w_l [H,d], b_l [H]  (MultiheadLinear, 547-556).
    non-MULTI_HEAD: attn_shift_l [H*d], w_l [1,H*d], b_l [1] (nn.Linear(D,1), 604).
    Returns [B,T,H,d] (multi-head) or [B,T,H*d].
    """
    B, H, T, d = q.shape
    if head_dim is None:
        head_dim = d                                   # line 683: lmm_hidden_dim // lmm_num_head
    if not active:                                     # line 685 / 732
        return attn_output
    q_t = q.transpose(1, 2)                            # line 687
    if not multi_head:                                 # lines 689-692
        q_t = q_t.reshape(B, T, -1)
        attn_output = attn_output.reshape(B, T, -1)
    if not learnable_scale:                            # lines 717-729: VECTOR_SHIFT alone adds the bare vector
        return attn_output + attn_shift_l[None, None, :]
    # lines 697-703: UNMASKED log-sum-exp over keys
    log_z2 = torch.logsumexp(torch.matmul(q, k.transpose(-2, -1)) / (head_dim ** 0.5), dim=-1).transpose(-2, -1)
    if not multi_head:                                 # lines 705-707
        log_z2 = log_z2.mean(-1, keepdim=True)
        log_z1 = q_t @ w_l.t() + b_l                   # nn.Linear(D,1)
    else:
        log_z1 = torch.einsum("btnd,nd->btn", q_t, w_l) + b_l      # line 556 / 709
    mu = torch.exp(log_z1 - torch.logaddexp(log_z1, log_z2))       # line 712
    if multi_head:
        mu = mu.unsqueeze(-1)                          # line 715
    shift = attn_shift_l[None, None, :]                # line 719
    return attn_output + mu * shift                    # line 724


def shifted_attention_fwd(q, k, v, attn_shift_l=None, w_l=None, b_l=None, *, causal=True,
                          key_mask=None, scale=None, active=True):
    """The fused op the CUDA kernels implement: SDPA followed by do_shift.

    q [B,H,T,d]; k, v [B,Hkv,S,d] (GQA heads are repeated like src/shift_encoder.py:431-432).
    Returns dict(out [B,T,H,d], lse_c [B,H,T], lse_full [B,H,T], mu [B,H,T]).
    """
    B, H, T, d = q.shape
    Hkv, S = k.shape[1], k.shape[2]
    if scale is None:
        scale = 1.0 / math.sqrt(d)
    kk, vv = repeat_kv(k, H // Hkv), repeat_kv(v, H // Hkv)
    vis = visibility(T, S, B, causal, key_mask, device=q.device)
    sa, lse_c = masked_attention(q, kk, vv, vis, scale)
    sa_t = sa.transpose(1, 2)                          # [B,T,H,d]  (src/shift_encoder.py:367 / 458)
    res = {"sa": sa_t, "lse_c": lse_c}
    if not active:
        res.update(out=sa_t, lse_full=None, mu=None)
        return res
    s_full = torch.matmul(q, kk.transpose(-2, -1)) * scale
    lse_full = torch.logsumexp(s_full, dim=-1)         # [B,H,T]
    z1 = torch.einsum("bhtd,hd->bht", q, w_l) + b_l[None, :, None]
    mu = torch.exp(z1 - torch.logaddexp(z1, lse_full))
    out = sa_t + mu.transpose(1, 2)[..., None] * attn_shift_l[None, None]
    res.update(out=out, lse_full=lse_full, mu=mu, z1=z1)
    return res


def shifted_attention_bwd(q, k, v, attn_shift_l, w_l, b_l, d_out, *, causal=True,
                          key_mask=None, scale=None, active=True):
    """Explicit backward, independent of autograd.

    d_out [B,T,H,d].  Returns dict(dq [B,H,T,d], dk, dv [B,Hkv,S,d], dshift [H,d], dw [H,d], db [H]).
    """
    B, H, T, d = q.shape
    Hkv, S = k.shape[1], k.shape[2]
    G = H // Hkv
    if scale is None:
        scale = 1.0 / math.sqrt(d)
    kk, vv = repeat_kv(k, G), repeat_kv(v, G)
    vis = visibility(T, S, B, causal, key_mask, device=q.device)
    s = torch.matmul(q, kk.transpose(-2, -1)) * scale                     # [B,H,T,S]
    sm = s.masked_fill(~vis, float("-inf"))
    lse_c = torch.logsumexp(sm, dim=-1)
    lse_c_safe = torch.where(torch.isinf(lse_c), torch.zeros_like(lse_c), lse_c)
    p_c = torch.where(vis, torch.exp(s - lse_c_safe[..., None]), torch.zeros_like(s))
    g = d_out.transpose(1, 2)                                             # [B,H,T,d]
    sa = torch.matmul(p_c, vv)
    delta = (g * sa).sum(-1)                                              # Δ_t = g_t · SA_t
    dp = torch.matmul(g, vv.transpose(-2, -1))                            # g_t · v_s
    ds = p_c * (dp - delta[..., None])
    dv_rep = torch.matmul(p_c.transpose(-2, -1), g)
    res = {}
    dq = torch.zeros_like(q)
    if active:
        lse_full = torch.logsumexp(s, dim=-1)
        z1 = torch.einsum("bhtd,hd->bht", q, w_l) + b_l[None, :, None]
        mu = torch.sigmoid(z1 - lse_full)                                 # == exp(z1 - logaddexp(z1, lse_full))
        dmu = torch.einsum("bhtd,hd->bht", g, attn_shift_l)
        dz = mu * (1.0 - mu) * dmu
        p_f = torch.exp(s - lse_full[..., None])
        ds = ds - dz[..., None] * p_f
        dq = dq + dz[..., None] * w_l[None, :, None, :]
        res["dshift"] = torch.einsum("bht,bhtd->hd", mu, g)
        res["dw"] = torch.einsum("bht,bhtd->hd", dz, q)
        res["db"] = dz.sum(dim=(0, 2))
    dq = dq + scale * torch.matmul(ds, kk)
    dk_rep = scale * torch.matmul(ds.transpose(-2, -1), q)
    res["dq"] = dq
    res["dk"] = dk_rep.view(B, Hkv, G, S, d).sum(2)
    res["dv"] = dv_rep.view(B, Hkv, G, S, d).sum(2)
    return res


def shifted_attention_autograd(q, k, v, attn_shift_l, w_l, b_l, d_out, **kw):
    """Forward restatement + torch autograd: the cross-check for `shifted_attention_bwd`."""
    leaves = [t.detach().clone().requires_grad_(True) for t in (q, k, v, attn_shift_l, w_l, b_l)]
    out = shifted_attention_fwd(*leaves, **kw)["out"]
    grads = torch.autograd.grad(out, leaves, d_out, allow_unused=True)
    names = ["dq", "dk", "dv", "dshift", "dw", "db"]
    return {n: g for n, g in zip(names, grads)}, out.detach()
