"""torch.autograd front-ends over the C ABI (include/mimic_b200.h).  PyTorch is plumbing here:
device memory, the current stream, autograd bookkeeping.  All arithmetic is in libmimic_b200.so;
there is no eager / CPU fallback — CPU tensors raise."""
from __future__ import annotations

import math
from typing import List, Optional, Sequence

import torch

from . import _lib
from ._lib import AttnDesc, MimicError, check, lib

_DT = {torch.bfloat16: _lib.BF16, torch.float32: _lib.F32}

# launch counter: every kernel family launch made through this module (bench.py reports it)
launch_count = 0


def _count(n: int) -> None:
    global launch_count
    launch_count += n


class KernelTimer:
    """Optional CUDA-event brackets around every C-ABI call, on the launching stream (bench.py uses
    them to report per-kernel device time inside the timed region)."""

    def __init__(self):
        self.enabled = False
        self.records = {}

    def start(self, key):
        if not self.enabled:
            return None
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        self.records.setdefault(key, []).append((a, b))
        return b

    @staticmethod
    def stop(ev):
        if ev is not None:
            ev.record()

    def summary(self):
        """key -> (launches, mean ms, total ms); call after a device synchronize."""
        out = {}
        for k, evs in self.records.items():
            ms = [a.elapsed_time(b) for a, b in evs]
            out[k] = (len(ms), sum(ms) / len(ms), sum(ms))
        return out

    def reset(self):
        self.records = {}


timer = KernelTimer()


_zero_ws_cache = {}
graph_keepalive = []          # pinned host staging captured by CUDA graphs must outlive the autograd graph


def _zero_ws(dev, nbytes: int, tag: str) -> torch.Tensor:
    """Persistent zero-initialised workspace (the single-launch loss kernels keep a self-resetting arrival
    counter in it).  One per (device, kernel); grown on demand."""
    key = (dev, tag)
    ws = _zero_ws_cache.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.zeros(max(nbytes, 4096), dtype=torch.uint8, device=dev)
        _zero_ws_cache[key] = ws
    return ws


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _require_cuda(*ts) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise MimicError("mimic_b200 ops run on CUDA tensors only (no CPU fallback)")


def _rows_ok(t: torch.Tensor) -> torch.Tensor:
    """last dim contiguous, other strides multiples of 8 elements, 16-byte aligned base"""
    if t.stride(-1) != 1 or any(s % 8 for s in t.stride()[:-1]) or t.data_ptr() % 16:
        return t.contiguous()
    return t


def _make_desc(q, k, v, out, scale, causal, active, impl) -> AttnDesc:
    B, H, T, d = q.shape
    Hkv, S = k.shape[1], k.shape[2]
    return AttnDesc(B=B, H=H, Hkv=Hkv, T=T, S=S, d=d,
                    q_sb=q.stride(0), q_sh=q.stride(1), q_st=q.stride(2),
                    k_sb=k.stride(0), k_sh=k.stride(1), k_ss=k.stride(2),
                    v_sb=v.stride(0), v_sh=v.stride(1), v_ss=v.stride(2),
                    o_sb=out.stride(0), o_st=out.stride(1), o_sh=out.stride(2),
                    scale=float(scale), causal=int(bool(causal)), active=int(bool(active)),
                    dtype=_DT[q.dtype], impl=_lib.IMPL_BY_NAME[impl])


def shifted_attention_forward(q, k, v, key_mask, shift, w, bias, *, causal, active, scale, impl="auto"):
    """Raw forward: returns (out [B,T,H,d], lse_c, lse_full, mu) with no autograd."""
    _require_cuda(q, k, v, key_mask, shift, w, bias)
    if q.dtype not in _DT or k.dtype != q.dtype or v.dtype != q.dtype:
        raise MimicError(f"q/k/v must share dtype bf16 or fp32 (got {q.dtype}, {k.dtype}, {v.dtype})")
    q, k, v = _rows_ok(q), _rows_ok(k), _rows_ok(v)
    B, H, T, d = q.shape
    out = torch.empty(B, T, H, d, dtype=q.dtype, device=q.device)
    lse_c = torch.empty(B, H, T, dtype=torch.float32, device=q.device)
    lse_full = torch.empty_like(lse_c) if active else None
    mu = torch.empty_like(lse_c) if active else None
    if key_mask is not None:
        key_mask = key_mask.to(torch.uint8).contiguous()
        if key_mask.shape != (B, k.shape[2]):
            raise MimicError(f"key_mask must be [B,S]={B, k.shape[2]}, got {tuple(key_mask.shape)}")
    if active:
        shift, w, bias = (t.float().contiguous() for t in (shift, w, bias))
    desc = _make_desc(q, k, v, out, scale, causal, active, impl)
    _ev = timer.start(f"attn_fwd_{'active' if active else 'inactive'}_T{T}_S{k.shape[2]}")
    check(lib.mimic_shifted_attn_fwd(desc, _ptr(q), _ptr(k), _ptr(v), _ptr(key_mask), _ptr(shift), _ptr(w),
                                     _ptr(bias), _ptr(out), _ptr(lse_c), _ptr(lse_full), _ptr(mu), _stream()),
          "mimic_shifted_attn_fwd")
    timer.stop(_ev)
    _count(1)
    return out, lse_c, lse_full, mu, (q, k, v, key_mask, shift, w, bias)


def shifted_attention_backward(saved, out, d_out, lse_c, lse_full, mu, *, causal, active, scale, impl="auto",
                               dshift=None, dw=None, dbias=None, d_lse_full=None):
    """Raw backward.  dshift/dw/dbias (fp32) are ACCUMULATED INTO when given, else allocated zeroed.
    d_lse_full [B,H,T] fp32: optional gradient arriving at lse_full from outside the op."""
    q, k, v, key_mask, shift, w, bias = saved
    B, H, T, d = q.shape
    if d_out.stride(-1) != 1 or any(s % 8 for s in d_out.stride()[:-1]) or d_out.data_ptr() % 16:
        d_out = d_out.contiguous()
    dq = torch.empty(B, H, T, d, dtype=q.dtype, device=q.device)
    dk = torch.empty(k.shape, dtype=q.dtype, device=q.device)
    dv = torch.empty(v.shape, dtype=q.dtype, device=q.device)
    if active:
        if dshift is None:
            dshift = torch.zeros(H, d, dtype=torch.float32, device=q.device)
            dw = torch.zeros(H, d, dtype=torch.float32, device=q.device)
            dbias = torch.zeros(H, dtype=torch.float32, device=q.device)
    desc = _make_desc(q, k, v, out, scale, causal, active, impl)
    nbytes = lib.mimic_shifted_attn_bwd_workspace(desc)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=q.device)
    _ev = timer.start(f"attn_bwd_{'active' if active else 'inactive'}_T{T}_S{k.shape[2]}")
    if d_lse_full is not None:
        d_lse_full = d_lse_full.float().contiguous()
    check(lib.mimic_shifted_attn_bwd_lse(desc, _ptr(q), _ptr(k), _ptr(v), _ptr(key_mask), _ptr(shift), _ptr(w),
                                         _ptr(bias), _ptr(out), _ptr(d_out), d_out.stride(0), d_out.stride(1),
                                         d_out.stride(2), _ptr(lse_c), _ptr(lse_full), _ptr(mu), _ptr(d_lse_full),
                                         _ptr(dq), _ptr(dk), _ptr(dv), _ptr(dshift), _ptr(dw), _ptr(dbias), _ptr(ws),
                                         nbytes, _stream()),
          "mimic_shifted_attn_bwd_lse")
    timer.stop(_ev)
    _count(3)
    return dq, dk, dv, dshift, dw, dbias


class GradSink:
    """Views into the flat fp32 parameter-gradient buffer for one layer: the attention backward
    accumulates straight into them (no per-layer temporaries, one all-reduce over the flat buffer)."""

    def __init__(self, dshift: torch.Tensor, dw: torch.Tensor, dbias: torch.Tensor):
        self.dshift, self.dw, self.dbias = dshift, dw, dbias


class _ShiftedAttention(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v, shift, w, bias, key_mask, causal, active, scale, impl, sink):
        out, lse_c, lse_full, mu, saved = shifted_attention_forward(
            q, k, v, key_mask, shift, w, bias, causal=causal, active=active, scale=scale, impl=impl)
        # inputs AND the op's own outputs go through save_for_backward: no grad_fn <-> output reference cycle (the
        # activations are freed by reference counting), and saved-tensor hooks / in-place version checks see them
        ctx.save_for_backward(*saved, out, lse_c, lse_full, mu)
        ctx.cfg = (causal, active, scale, impl)
        ctx.sink = sink
        ctx.param_dtypes = tuple(None if t is None else t.dtype for t in (shift, w, bias))
        ctx.mark_non_differentiable(lse_c)
        return out, lse_c

    @staticmethod
    def backward(ctx, d_out, _d_lse):
        causal, active, scale, impl = ctx.cfg
        *saved, out, lse_c, lse_full, mu = ctx.saved_tensors
        sink = ctx.sink
        dq, dk, dv, dshift, dw, dbias = shifted_attention_backward(
            tuple(saved), out, d_out, lse_c, lse_full, mu, causal=causal, active=active, scale=scale, impl=impl,
            dshift=None if sink is None else sink.dshift, dw=None if sink is None else sink.dw,
            dbias=None if sink is None else sink.dbias)
        if sink is not None or not active:
            dshift = dw = dbias = None
        else:
            dshift, dw, dbias = (g.to(dt) for g, dt in zip((dshift, dw, dbias), ctx.param_dtypes))
        return dq, dk, dv, dshift, dw, dbias, None, None, None, None, None, None


class _AttentionWithFullLse(torch.autograd.Function):
    """(SA [B,T,H,d], lse_full [B,H,T]) from ONE pass of the fused kernel, both differentiable: the op is run active
    with shift = w = bias = 0, so its own mu term vanishes and out == SA; the gradient that arrives at lse_full goes
    back in through mimic_shifted_attn_bwd_lse."""

    @staticmethod
    def forward(ctx, q, k, v, key_mask, causal, scale, impl):
        B, H, T, d = q.shape
        zero_hd = torch.zeros(H, d, dtype=torch.float32, device=q.device)
        zero_h = torch.zeros(H, dtype=torch.float32, device=q.device)
        out, lse_c, lse_full, mu, saved = shifted_attention_forward(
            q, k, v, key_mask, zero_hd, zero_hd, zero_h, causal=causal, active=True, scale=scale, impl=impl)
        ctx.save_for_backward(*saved, out, lse_c, lse_full, mu)
        ctx.cfg = (causal, scale, impl)
        return out, lse_full

    @staticmethod
    def backward(ctx, d_out, d_lse_full):
        causal, scale, impl = ctx.cfg
        *saved, out, lse_c, lse_full, mu = ctx.saved_tensors
        if d_out is None:
            d_out = torch.zeros_like(out)
        dq, dk, dv, _, _, _ = shifted_attention_backward(
            tuple(saved), out, d_out, lse_c, lse_full, mu, causal=causal, active=True, scale=scale, impl=impl,
            d_lse_full=d_lse_full)
        return dq, dk, dv, None, None, None, None


def attention_with_full_lse(q, k, v, key_mask=None, *, causal=True, scale=None, impl="auto"):
    """Masked attention output plus the UNMASKED log-sum-exp of the scores (log_Z2 of src/shift_encoder.py:697-703)
    per head — what the non-MULTI_HEAD variant of do_shift averages over the heads (705-707)."""
    if scale is None:
        scale = 1.0 / math.sqrt(q.shape[-1])
    return _AttentionWithFullLse.apply(q, k, v, key_mask, causal, scale, impl)


# ------------------------------------------------------------------------------------------------
class _NormShift(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h, shift, scale):
        D = h.shape[-1]
        h2 = h.reshape(-1, D)
        if h2.stride(-1) != 1 or h2.stride(0) % 8 or h2.data_ptr() % 16:
            h2 = h2.contiguous()
        sh = shift.detach().float().contiguous()
        if sh.data_ptr() % 16:                 # a row of a parameter that sits at an odd offset of the flat optimizer buffer
            sh = sh.clone()
        sc = None if scale is None else scale.detach().float().reshape(1).contiguous()
        y = torch.empty(h2.shape, dtype=h.dtype, device=h.device)
        if h2.shape[0]:
            _ev = timer.start("norm_shift_fwd")
            check(lib.mimic_norm_shift_fwd(_ptr(h2), h2.stride(0), _DT[h.dtype], _ptr(sh), _ptr(sc), _ptr(y),
                                           h2.shape[0], D, _stream()), "mimic_norm_shift_fwd")
            timer.stop(_ev)
            _count(1)
        ctx.saved = (h2, sh, sc)
        ctx.meta = (h.shape, shift.dtype, None if scale is None else (scale.dtype, scale.shape))
        return y.view(h.shape)

    @staticmethod
    def backward(ctx, dy):
        h2, sh, sc = ctx.saved
        shape, shift_dtype, scale_meta = ctx.meta
        D = shape[-1]
        dy2 = dy.reshape(-1, D)
        if dy2.dtype != h2.dtype:
            dy2 = dy2.to(h2.dtype)
        if dy2.stride(-1) != 1 or dy2.stride(0) % 8 or dy2.data_ptr() % 16:
            dy2 = dy2.contiguous()
        dh = torch.empty(h2.shape, dtype=h2.dtype, device=h2.device)
        gsum = torch.zeros(D, dtype=torch.float32, device=h2.device)
        if h2.shape[0]:
            _ev = timer.start("norm_shift_bwd")
            check(lib.mimic_norm_shift_bwd(_ptr(h2), h2.stride(0), _ptr(dy2), dy2.stride(0), _DT[h2.dtype], _ptr(sh),
                                           _ptr(sc), _ptr(dh), _ptr(gsum), h2.shape[0], D, _stream()),
                  "mimic_norm_shift_bwd")
            timer.stop(_ev)
            _count(1)
        dshift = (gsum if sc is None else gsum * sc).to(shift_dtype)
        dscale = None
        if scale_meta is not None:
            dscale = torch.dot(gsum, sh).to(scale_meta[0]).reshape(scale_meta[1])
        return dh.view(shape), dshift, dscale


def norm_shift(h: torch.Tensor, shift: torch.Tensor, scale: Optional[torch.Tensor] = None) -> torch.Tensor:
    """y = u / |u| * |h| with u = h + scale * shift: the norm-preserving shift hook of AttnFFNShift
    (src/shift_encoder.py:233-240) and of AttnApproximator's FFN hook (661-667), one launch each way.
    h [..., D] bf16/fp32, shift [D], scale a one-element tensor or None (= 1).  Output in h's dtype."""
    _require_cuda(h, shift, scale)
    if h.dtype not in _DT:
        raise MimicError(f"norm_shift: h must be bf16 or fp32 (got {h.dtype})")
    D = h.shape[-1]
    if D % 8 or D > 8192 or shift.shape != (D,):
        raise MimicError(f"norm_shift: hidden size {D} unsupported (multiple of 8, at most 8192) or shift shape "
                         f"{tuple(shift.shape)} != ({D},)")
    return _NormShift.apply(h, shift, scale)


def shifted_attention(q, k, v, shift=None, w=None, bias=None, key_mask=None, *, causal=True, active=True,
                      scale=None, impl="auto", grad_sink: Optional[GradSink] = None, return_lse=False):
    """out[B,T,H,d] = SDPA(q,k,v) + mu(q,k) * shift   — the fused replacement of
    F.scaled_dot_product_attention + AttnApproximator.do_shift (src/shift_encoder.py:352-359, 682-732).

    q [B,H,T,d], k/v [B,Hkv,S,d], shift/w [H,d], bias [H], key_mask [B,S] (True/1 = attend)."""
    if scale is None:
        scale = 1.0 / math.sqrt(q.shape[-1])
    out, lse_c = _ShiftedAttention.apply(q, k, v, shift, w, bias, key_mask, causal, active, scale, impl, grad_sink)
    return (out, lse_c) if return_lse else out


# ------------------------------------------------------------------------------------------------
def _rope_launch(q, k, cos, sin, inverse: bool):
    """q [B,T,H,d], k [B,T,Hkv,d] (any strides with unit head-dim stride) -> contiguous rotated copies."""
    B, T, H, d = q.shape
    Hkv = k.shape[2]
    if q.stride(3) != 1:
        q = q.contiguous()
    if k.stride(3) != 1:
        k = k.contiguous()
    cos = cos.to(q.dtype)
    sin = sin.to(q.dtype)
    if cos.dim() == 2:
        cos, sin = cos[None], sin[None]
    if cos.stride(-1) != 1 or sin.stride() != cos.stride():
        cos, sin = cos.contiguous(), sin.contiguous()
    c_sb = 0 if cos.shape[0] == 1 else cos.stride(0)
    q_out = torch.empty((B, T, H, d), dtype=q.dtype, device=q.device)
    k_out = torch.empty((B, T, Hkv, d), dtype=k.dtype, device=k.device)
    _ev = timer.start("rope_qk")
    check(lib.mimic_rope_qk(_ptr(q), q.stride(0), q.stride(1), q.stride(2), _ptr(k), k.stride(0), k.stride(1), k.stride(2),
                            _ptr(cos), _ptr(sin), c_sb, cos.stride(1), _ptr(q_out), _ptr(k_out), _DT[q.dtype],
                            B, T, H, Hkv, d, int(inverse), _stream()), "mimic_rope_qk")
    timer.stop(_ev)
    _count(1)
    return q_out, k_out


class _RopeQK(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, cos, sin):
        ctx.save_for_backward(cos, sin)
        return _rope_launch(q, k, cos, sin, False)

    @staticmethod
    def backward(ctx, gq, gk):
        cos, sin = ctx.saved_tensors
        dq, dk = _rope_launch(gq, gk, cos, sin, True)
        return dq, dk, None, None


def rope_qk(q: torch.Tensor, k: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor):
    """HF apply_rotary_pos_emb on q [B,T,H,d] and k [B,T,Hkv,d] (token-major views of the projection outputs) with
    cos/sin [B,T,d] (or [T,d]): ONE launch instead of ~10 elementwise ones; returns contiguous [B,T,H,d], [B,T,Hkv,d].
    Differentiable w.r.t. q and k (the backward is the transposed rotation, same kernel)."""
    _require_cuda(q, k, cos, sin)
    return _RopeQK.apply(q, k, cos, sin)


def record_gather(src: torch.Tensor, src_row: torch.Tensor, dst: torch.Tensor) -> None:
    """dst[n,:] = src.view(-1,D)[src_row[n],:]   (kernel R; src_row int32 on device)."""
    _require_cuda(src, src_row, dst)
    D = src.shape[-1]
    src2 = src.reshape(-1, D)
    if src2.stride(1) != 1:
        src2 = src2.contiguous()
    N = src_row.numel()
    assert dst.shape == (N, D) and dst.is_contiguous() and src_row.dtype == torch.int32
    _ev = timer.start("record_gather")
    check(lib.mimic_record_gather(_ptr(src2), src2.stride(0), _DT[src2.dtype], _ptr(src_row), _ptr(dst), _DT[dst.dtype],
                                  N, D, _stream()), "mimic_record_gather")
    timer.stop(_ev)
    _count(1)


class RowPlan:
    """Outputs of kernel P (mimic_row_plan) for one batch of token ids; None where not requested."""
    __slots__ = ("attn_mask", "label_mask", "rows", "row_to_compact", "row_w", "lm_labels", "paired", "count", "row_counts",
                 "sep_counts")


def row_plan(ids: torch.Tensor, *, pad_id: int, bos_id: int, left_padding: bool, num_separator: int = 0, select: str = "label",
             keep_bos: bool = False, cap: int = 0, w_scale: float = 0.0, want_attn_mask: bool = True, want_label_mask: bool = False,
             want_row_to_compact: bool = False, want_row_w: bool = False, want_lm_labels: bool = False,
             pair_src: Optional[torch.Tensor] = None) -> RowPlan:
    """ONE launch for the integer side of a MimIC pass (include/mimic_b200.h, mimic_row_plan): attention mask, label mask
    (ShiftModel.generate_label_mask), compact row plan (ascending flat indices, -1 padded to `cap`), row -> compact index,
    per-row loss weights, shifted LM labels, pairing with another plan.  ids [B,S] int64 on the device; no host sync."""
    _require_cuda(ids, pair_src)
    if ids.dtype != torch.int64 or ids.dim() != 2:
        raise MimicError(f"row_plan: ids must be int64 [B,S] (got {ids.dtype}, {tuple(ids.shape)})")
    ids = ids.contiguous()
    B, S = ids.shape
    dev = ids.device
    out = RowPlan()
    out.attn_mask = torch.empty(B, S, dtype=torch.bool, device=dev) if want_attn_mask else None
    out.label_mask = torch.empty(B, S, dtype=torch.bool, device=dev) if want_label_mask else None
    out.rows = torch.empty(cap, dtype=torch.int32, device=dev) if cap > 0 else None
    out.row_to_compact = torch.empty(B * S, dtype=torch.int32, device=dev) if want_row_to_compact else None
    out.row_w = torch.empty(cap, dtype=torch.float32, device=dev) if (want_row_w and cap > 0) else None
    out.lm_labels = torch.empty(B, S, dtype=torch.int64, device=dev) if want_lm_labels else None
    out.paired = torch.empty(B * S, dtype=torch.int32, device=dev) if pair_src is not None else None
    counts = torch.empty(1 + 2 * B, dtype=torch.int32, device=dev)
    out.count, out.row_counts, out.sep_counts = counts[:1], counts[1:1 + B], counts[1 + B:]
    if pair_src is not None:
        pair_src = pair_src.to(torch.int32).contiguous()
    _ev = timer.start("row_plan")
    check(lib.mimic_row_plan(_ptr(ids), B, S, int(pad_id), int(bos_id), int(bool(left_padding)), int(num_separator),
                             {"label": 0, "non_pad_non_bos": 1}[select], int(bool(keep_bos)), float(w_scale),
                             _ptr(out.attn_mask), _ptr(out.label_mask), _ptr(out.rows), int(cap), _ptr(out.row_to_compact),
                             _ptr(out.row_w), _ptr(out.lm_labels), _ptr(pair_src), 0 if pair_src is None else pair_src.numel(),
                             _ptr(out.paired), counts.data_ptr(), counts.data_ptr() + 4, counts.data_ptr() + 4 * (1 + B),
                             _stream()), "mimic_row_plan")
    timer.stop(_ev)
    _count(1)
    return out


class _AlignMSE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, hp, row_to_compact, row_w, grad_scale, assume_folded, *hs):
        return _align_forward(ctx, "mse", hp, row_to_compact, row_w, grad_scale, assume_folded, *hs)

    @staticmethod
    def backward(ctx, g):
        return _align_backward(ctx, g)


class _AlignCos(torch.autograd.Function):
    @staticmethod
    def forward(ctx, hp, row_to_compact, row_w, grad_scale, assume_folded, *hs):
        return _align_forward(ctx, "cos", hp, row_to_compact, row_w, grad_scale, assume_folded, *hs)

    @staticmethod
    def backward(ctx, g):
        return _align_backward(ctx, g)


def _align_forward(ctx, kind, hp, row_to_compact, row_w, grad_scale, assume_folded, *hs):
    L = len(hs)
    h0 = hs[0]
    D = h0.shape[-1]
    rows = h0.numel() // D
    # rows are read as 16-byte vectors: contiguous and 16-byte aligned (a contiguous view at an odd offset is copied)
    hs = [(h if h.data_ptr() % 16 == 0 else h.clone()) if h.is_contiguous() else h.contiguous() for h in hs]
    N = hp.shape[1]
    dev = h0.device
    need_grad = any(ctx.needs_input_grad[5:])
    ptrs = [h.data_ptr() for h in hs]
    dh = torch.empty((L, rows, D), dtype=h0.dtype, device=dev) if need_grad else None
    if need_grad:
        step = rows * D * dh.element_size()
        ptrs += [dh.data_ptr() + l * step for l in range(L)]
    ptr_host = torch.tensor(ptrs, dtype=torch.int64).pin_memory()
    ptr_t = torch.empty(len(ptrs), dtype=torch.int64, device=dev)
    ptr_t.copy_(ptr_host, non_blocking=True)
    if torch.cuda.is_current_stream_capturing():
        graph_keepalive.append(ptr_host)
    loss = torch.empty(1, dtype=torch.float32, device=dev)
    nbytes = lib.mimic_mse_workspace(L, rows, D)
    ws = _zero_ws(dev, nbytes, "mse")
    _ev = timer.start(f"{kind}_fwd_bwd")
    fn = lib.mimic_mse_fwd_bwd if kind == "mse" else lib.mimic_cos_fwd_bwd
    check(fn(ptr_t.data_ptr(), D, _DT[h0.dtype], _ptr(hp), _DT[hp.dtype], _ptr(row_to_compact),
             _ptr(row_w), (ptr_t.data_ptr() + 8 * L) if need_grad else None, float(grad_scale),
             L, rows, N, D, _ptr(loss), _ptr(ws), nbytes, _stream()), f"mimic_{kind}_fwd_bwd")
    timer.stop(_ev)
    _count(1)
    ctx.dh = dh
    ctx.shapes = [h.shape for h in hs]
    ctx.grad_scale = grad_scale
    ctx.assume_folded = assume_folded
    ctx.keepalive = (ptr_t, ptr_host, hs)
    return loss[0]


def _unfold(g: torch.Tensor, grad_scale: float) -> torch.Tensor:
    """Factor that turns a gradient pre-scaled by `grad_scale` into one scaled by the upstream gradient `g`
    (a zero weight wrote zeros: keep them)."""
    return g / grad_scale if grad_scale != 0.0 else torch.zeros_like(g)


def _align_backward(ctx, g):
    dh = ctx.dh
    if dh is None:
        return (None,) * (5 + len(ctx.shapes))
    if not ctx.assume_folded:
        dh = dh * _unfold(g, ctx.grad_scale).to(dh.dtype)
    return (None, None, None, None, None) + tuple(dh[l].view(s) for l, s in enumerate(ctx.shapes))


def align_mse(hs: Sequence[torch.Tensor], hp: torch.Tensor, row_to_compact: torch.Tensor, row_w: torch.Tensor,
              grad_scale: float = 1.0, assume_folded: bool = False) -> torch.Tensor:
    """Layer-wise MSE (src/shift_model.py:145-179) over L recorded tensors [*, D] against the compact
    H' buffer [L,N,D]; one fused forward+backward launch.  With `assume_folded`, the caller promises
    the upstream gradient of the returned scalar is exactly `grad_scale` (it was folded into dH)."""
    _require_cuda(hp, row_to_compact, row_w, *hs)
    return _AlignMSE.apply(hp, row_to_compact, row_w, grad_scale, assume_folded, *hs)


def align_cos(hs: Sequence[torch.Tensor], hp: torch.Tensor, row_to_compact: torch.Tensor, row_w: torch.Tensor,
              grad_scale: float = 1.0, assume_folded: bool = False) -> torch.Tensor:
    """Layer-wise cosine-similarity loss (Strategy.LAYER_WISE_COS_SIM, src/shift_model.py:152-160): same calling
    convention as align_mse, with row_w[n] = 1 / (n_samples * L * Nq_sample(n)) (no 1/D)."""
    _require_cuda(hp, row_to_compact, row_w, *hs)
    return _AlignCos.apply(hp, row_to_compact, row_w, grad_scale, assume_folded, *hs)


def _grad_like(x: torch.Tensor) -> torch.Tensor:
    """Gradient buffer for a [R, V] row view: SAME row stride and the same address modulo 32 bytes as `x` (the loss kernels
    address logits and gradient with one stride and vector-align their chunks on the logits row)."""
    if x.is_contiguous() and x.data_ptr() % 32 == 0:
        return torch.empty_like(x)
    R, V = x.shape
    es = x.element_size()
    n = (R - 1) * x.stride(0) + V
    buf = torch.empty(n + 32 // es, dtype=x.dtype, device=x.device)
    off = ((x.data_ptr() - buf.data_ptr()) % 32) // es
    return buf.as_strided((R, V), (x.stride(0), 1), off)


class _KLLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, y, y_row, grad_scale, assume_folded, inplace):
        V = z.shape[-1]
        zz = z.reshape(-1, V)
        yy = y.reshape(-1, V)
        if zz.stride(1) != 1:
            zz = zz.contiguous()
        if yy.stride(1) != 1:
            yy = yy.contiguous()
        R = zz.shape[0]
        y_row = y_row.reshape(-1).to(torch.int32).contiguous()
        assert y_row.numel() == R and yy.dtype == zz.dtype
        dev = zz.device
        need_grad = ctx.needs_input_grad[0]
        dz = (zz if inplace else _grad_like(zz)) if need_grad else None
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        n_rows = torch.empty(1, dtype=torch.int32, device=dev)
        nbytes = lib.mimic_ce_workspace(R)
        ws = _zero_ws(dev, nbytes, "kl")
        _ev = timer.start("kl_fwd_bwd")
        check(lib.mimic_kl_fwd_bwd(_ptr(zz), zz.stride(0), _ptr(yy), yy.stride(0), _DT[zz.dtype], _ptr(y_row), _ptr(dz),
                                   float(grad_scale), R, V, _ptr(loss), _ptr(n_rows), _ptr(ws), nbytes, _stream()),
              "mimic_kl_fwd_bwd")
        timer.stop(_ev)
        _count(1)
        ctx.dz = dz
        ctx.shape = z.shape
        ctx.grad_scale = grad_scale
        ctx.assume_folded = assume_folded
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        d = ctx.dz
        if d is None:
            return (None,) * 6
        if not ctx.assume_folded:
            d = d * _unfold(g, ctx.grad_scale).to(d.dtype)
        return d.reshape(ctx.shape), None, None, None, None, None


def kl_loss(shift_logits: torch.Tensor, prefix_logits: torch.Tensor, prefix_row: torch.Tensor, grad_scale: float = 1.0,
            assume_folded: bool = False, inplace: bool = False) -> torch.Tensor:
    """KL(softmax(prefix) || softmax(shift)), batchmean over the selected rows (Strategy.LOGITS_KL_DIV,
    src/shift_model.py:181-191), forward+backward in one launch.  prefix_row[i] = row of prefix_logits.view(-1, V)
    paired with row i of shift_logits.view(-1, V), or -1 when row i is not selected.  The prefix logits get no gradient
    (they come from the no-grad context pass)."""
    _require_cuda(shift_logits, prefix_logits, prefix_row)
    return _KLLoss.apply(shift_logits, prefix_logits, prefix_row, grad_scale, assume_folded, inplace)


class _CELoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels, grad_scale, assume_folded, inplace):
        V = logits.shape[-1]
        lg = logits.reshape(-1, V)
        if lg.stride(1) != 1:
            lg = lg.contiguous()
        R = lg.shape[0]
        labels = labels.reshape(-1).to(torch.int64).contiguous()
        dev = lg.device
        need_grad = ctx.needs_input_grad[0]
        dlog = None
        if need_grad:
            dlog = lg if inplace else _grad_like(lg)
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        n_kept = torch.empty(1, dtype=torch.int32, device=dev)
        nbytes = lib.mimic_ce_workspace(R)
        ws = _zero_ws(dev, nbytes, "ce")
        _ev = timer.start("ce_fwd_bwd")
        check(lib.mimic_ce_fwd_bwd(_ptr(lg), lg.stride(0), _DT[lg.dtype], _ptr(labels), _ptr(dlog), float(grad_scale),
                                   R, V, _ptr(loss), _ptr(n_kept), _ptr(ws), nbytes, _stream()), "mimic_ce_fwd_bwd")
        timer.stop(_ev)
        _count(1)
        ctx.dlog = dlog
        ctx.shape = logits.shape
        ctx.grad_scale = grad_scale
        ctx.assume_folded = assume_folded
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        d = ctx.dlog
        if d is None:
            return None, None, None, None, None
        if not ctx.assume_folded:
            d = d * _unfold(g, ctx.grad_scale).to(d.dtype)
        return d.reshape(ctx.shape), None, None, None, None


def ce_loss(logits: torch.Tensor, labels: torch.Tensor, grad_scale: float = 1.0, assume_folded: bool = False,
            inplace: bool = False) -> torch.Tensor:
    """Mean CE over rows with label >= 0 (fp32 math), forward+backward in one launch.
    logits [..., V] (bf16/fp32), labels [...] int64 (-100 = ignore)."""
    _require_cuda(logits, labels)
    return _CELoss.apply(logits, labels, grad_scale, assume_folded, inplace)


def adamw_hyper(lr: float, step: int, betas=(0.9, 0.999)) -> torch.Tensor:
    """Host-side [lr, 1-beta1^step, sqrt(1-beta2^step)] (pinned) — copy it into the device `hyper` tensor."""
    return torch.tensor([lr, 1.0 - betas[0] ** step, math.sqrt(1.0 - betas[1] ** step)], dtype=torch.float32)


def adamw_workspace(device) -> torch.Tensor:
    """Zeroed fp32 workspace of the clip + AdamW step (mimic_adamw_workspace() bytes); [0] = squared gradient norm after a step."""
    return torch.zeros(lib.mimic_adamw_workspace() // 4, dtype=torch.float32, device=device)


def adamw_step(param, grad, exp_avg, exp_avg_sq, weight_decay, hyper: torch.Tensor, *, betas=(0.9, 0.999),
               eps=1e-8, grad_mul=1.0, max_norm=0.0, sumsq: Optional[torch.Tensor] = None,
               lr_mul: Optional[torch.Tensor] = None) -> None:
    """Clip (global norm) + AdamW on flat fp32 buffers, on device (replaces DeepSpeedCPUAdam).
    hyper: device fp32 [3] = adamw_hyper(lr, step) (device-resident so a CUDA graph can replay the step).
    lr_mul: optional fp32 [n] per-element learning-rate multiplier (the reference's scale_lr group)."""
    _require_cuda(param, grad, exp_avg, exp_avg_sq, weight_decay, hyper, lr_mul)
    if sumsq is None:
        sumsq = adamw_workspace(param.device)
    check(lib.mimic_adamw_step_groups(_ptr(param), _ptr(grad), _ptr(exp_avg), _ptr(exp_avg_sq), _ptr(weight_decay),
                                      _ptr(lr_mul), param.numel(), _ptr(hyper), betas[0], betas[1], eps,
                                      float(grad_mul), float(max_norm), _ptr(sumsq), _stream()),
          "mimic_adamw_step_groups")
    _count(2)
