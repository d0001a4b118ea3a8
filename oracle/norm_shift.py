"""CPU restatement of the norm-preserving shift hook, TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows AttnFFNShift._shift_hook (src/shift_encoder.py:233-240) and the FFN hook of AttnApproximator
(src/shift_encoder.py:661-667, the same arithmetic with the scale fixed at 1).  Pinned by tests/golden/variants.npz
(keys ns_attn.*, ns_ffn.*, ns_approx.*), which oracle/gen_golden.py made by calling the reference's own hooks.
"""
from __future__ import annotations

import torch


def norm_shift_fwd(h: torch.Tensor, shift: torch.Tensor, scale=1.0) -> torch.Tensor:
    """h [..., D], shift [D]:  (h + scale*shift) / |h + scale*shift| * |h|   (lines 235-240)."""
    shifted = h + scale * shift
    return shifted / shifted.norm(dim=-1, keepdim=True) * h.norm(dim=-1, keepdim=True)


def norm_shift_bwd(h, shift, scale, dy):
    """Closed-form gradients of the above (what the CUDA backward computes), checked against autograd of
    norm_shift_fwd and against the reference-run fixture.  Returns (dh, dshift [D], dscale scalar).

    With u = h + s v, a = |h|, c = |u|, p = dy.u (per row):
        dL/du = (a/c) dy - (a p / c^3) u          dh = dL/du + (p / (a c)) h
        dshift = s * sum_rows dL/du               dscale = sum_rows dL/du . v
    """
    s = torch.as_tensor(scale, dtype=h.dtype)
    u = h + s * shift
    a = h.norm(dim=-1, keepdim=True)
    c = u.norm(dim=-1, keepdim=True)
    p = (dy * u).sum(-1, keepdim=True)
    gu = (a / c) * dy - (a * p / c ** 3) * u
    dh = gu + torch.where(a > 0, p / (a * c), torch.zeros_like(a)) * h
    gsum = gu.reshape(-1, h.shape[-1]).sum(0)
    return dh, s * gsum, (gsum * shift).sum()
