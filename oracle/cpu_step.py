"""CPU baseline: the reference's path for one MimIC train step, on host cores (torch CPU fp32).
TEST / BENCH-BASELINE INFRASTRUCTURE ONLY (bench.py `cpu_baseline` and `--impl reference`).

The reference is Python calling torch; what it runs per decoder layer is
F.scaled_dot_product_attention (src/shift_encoder.py:449-456) + AttnApproximator.do_shift (682-732),
then ShiftModel's stack/masked_select/mse_loss (src/shift_model.py:121-179) and the LM CE.  This file
drives exactly those oracle restatements around a plain functional decoder layer of the named
shape.  A full 32-layer 8 B-parameter fp32 model does not fit a bounded CPU sample, so the step is
timed on `n_layers` layers at the FULL token shape and the per-layer time is scaled to the model's
depth; the LM head + CE part is timed once.  `cpu_tokens_per_s` states the arithmetic.
"""
from __future__ import annotations

import time
from typing import Dict

import torch
import torch.nn.functional as F

from . import losses as OL
from . import shifted_attention as OA


def _rms(x, w, eps):
    return F.rms_norm(x, (x.shape[-1],), w, eps)


def _rot(x):
    a, b = x[..., : x.shape[-1] // 2], x[..., x.shape[-1] // 2:]
    return torch.cat((-b, a), -1)


def make_layer_weights(D, inter, H, Hkv, d, gen, std=0.02):
    mk = lambda *s: torch.randn(*s, generator=gen) * std
    return dict(q=mk(H * d, D), k=mk(Hkv * d, D), v=mk(Hkv * d, D), o=mk(D, H * d), gate=mk(inter, D),
                up=mk(inter, D), down=mk(D, inter), ln1=torch.ones(D), ln2=torch.ones(D))


def layer_forward(x, wt, key_mask, cos, sin, H, Hkv, d, shift=None, w=None, b=None, active=False, eps=1e-5):
    """One decoder layer, reference attention path: SDPA then do_shift.  Returns (x_out, mlp_out)."""
    B, T, D = x.shape
    h = _rms(x, wt["ln1"], eps)
    q = (h @ wt["q"].t()).view(B, T, H, d).transpose(1, 2)
    k = (h @ wt["k"].t()).view(B, T, Hkv, d).transpose(1, 2)
    v = (h @ wt["v"].t()).view(B, T, Hkv, d).transpose(1, 2)
    q, k = q * cos + _rot(q) * sin, k * cos + _rot(k) * sin
    kk, vv = OA.repeat_kv(k, H // Hkv), OA.repeat_kv(v, H // Hkv)           # shift_encoder.py:431-432
    mask = key_mask[:, None, None, :] & torch.ones(T, T, dtype=torch.bool).tril()[None, None]
    attn = F.scaled_dot_product_attention(q, kk, vv, attn_mask=mask)         # 449-456
    attn = attn.transpose(1, 2).contiguous()                                 # 458
    attn = OA.do_shift(q, kk, attn, shift, w, b, active=active)              # 461-464 -> 682-732
    x = x + attn.reshape(B, T, H * d) @ wt["o"].t()
    h2 = _rms(x, wt["ln2"], eps)
    mlp = (F.silu(h2 @ wt["gate"].t()) * (h2 @ wt["up"].t())) @ wt["down"].t()
    return x + mlp, mlp


def cpu_mimic_step(full_ids, qa_ids, *, D=4096, inter=14336, H=32, Hkv=8, vocab=32003, n_layers=1, model_layers=32,
                   pad_id=0, bos_id=1, seed=3407, theta=1e4) -> Dict[str, float]:
    """Times one MimIC train step (context pass, shifted query pass, losses, backward, AdamW) on
    `n_layers` layers; returns seconds for the layer part and the head part."""
    d = D // H
    gen = torch.Generator().manual_seed(seed)
    layers = [make_layer_weights(D, inter, H, Hkv, d, gen) for _ in range(n_layers)]
    emb = torch.randn(vocab, D, generator=gen) * 0.02
    head = torch.randn(vocab, D, generator=gen) * 0.02
    shift = (torch.randn(n_layers, H, d, generator=gen) * 1e-3).requires_grad_(True)
    w = (torch.randn(n_layers, H, d, generator=gen) * 0.02).requires_grad_(True)
    b = torch.zeros(n_layers, H, requires_grad=True)
    opt = torch.optim.AdamW([{"params": [b], "weight_decay": 0.0}, {"params": [shift, w], "weight_decay": 1e-3}], lr=5e-3)

    def rope(T):
        inv = 1.0 / (theta ** (torch.arange(0, d, 2, dtype=torch.float32) / d))
        ang = torch.arange(T, dtype=torch.float32)[:, None] * inv
        e = torch.cat((ang, ang), -1)
        return e.cos()[None, None], e.sin()[None, None]

    B, S = full_ids.shape
    T = qa_ids.shape[1]
    t0 = time.perf_counter()
    # step 1: context pass, no grad, record mlp outputs (shift_model.py:225-233)
    km = full_ids != pad_id
    cos, sin = rope(S)
    recorded = []
    with torch.no_grad():
        x = emb[full_ids]
        for wt in layers:
            x, mlp = layer_forward(x, wt, km, cos, sin, H, Hkv, d)
            recorded.append(mlp)
    hp = OL.get_hidden_states(recorded, OL.generate_label_mask(full_ids, pad_id, 1))
    del recorded
    # step 2: shifted query pass (237-243)
    kmq = qa_ids != pad_id
    cos, sin = rope(T)
    x = emb[qa_ids]
    rec_q = []
    for l, wt in enumerate(layers):
        x, mlp = layer_forward(x, wt, kmq, cos, sin, H, Hkv, d, shift[l], w[l], b[l], active=True)
        rec_q.append(mlp)
    hq = OL.get_hidden_states(rec_q, kmq & (qa_ids != bos_id))
    mse = OL.layer_wise_mse(hq, hp)
    t1 = time.perf_counter()
    logits = F.rms_norm(x, (D,)) @ head.t()
    ce = OL.lm_ce_loss(logits, qa_ids, kmq)
    t2 = time.perf_counter()
    loss = 0.5 * ce + 1.0 * mse
    opt.zero_grad()
    loss.backward()
    torch.nn.utils.clip_grad_norm_([shift, w, b], 1.0)
    opt.step()
    t3 = time.perf_counter()
    # backward time is dominated by the layer part; attribute it there
    return {"layers_s": (t1 - t0) + (t3 - t2), "head_s": t2 - t1, "loss": float(loss.detach()),
            "n_layers": n_layers, "model_layers": model_layers, "tokens": int(B * (S + T))}


def cpu_tokens_per_s(step: Dict[str, float]) -> float:
    """tokens / (per-layer seconds x model depth + head seconds)."""
    per_layer = step["layers_s"] / step["n_layers"]
    return step["tokens"] / (per_layer * step["model_layers"] + step["head_s"])
