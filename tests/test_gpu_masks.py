"""The integer side of the path ON THE DEVICE: label masks, compact row plans and query/context row
pairing computed from CUDA token tensors, bit-exact against the reference-made fixture (tests/golden/loss_pieces.npz,
src/shift_model.py:75-110, 121-143, 185-186) and against the oracle on random ragged batches."""
import os
import types

import numpy as np
import pytest
import torch

from oracle import losses as OL

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MIMIC = ("ShiftStrategy.VECTOR_SHIFT | ShiftStrategy.LEARNABLE_SHIFT_SCALE | ShiftStrategy.MULTI_HEAD",
         "ShiftStrategy.RECORD_HIDDEN_STATES")


def _shift_model(side="left"):
    from tests.test_host_logic import _tiny_lmm
    from mimic_b200 import AttnApproximator, ShiftModel, Strategy
    lmm = _tiny_lmm()
    lmm.processor = types.SimpleNamespace(tokenizer=types.SimpleNamespace(pad_token_id=0, bos_token_id=1, padding_side=side))
    return ShiftModel(types.SimpleNamespace(), AttnApproximator(lmm, *MIMIC), Strategy.LM_LOSS | Strategy.LAYER_WISE_MSE)


def test_generate_label_mask_on_cuda_bit_exact():
    z = np.load(os.path.join(GOLDEN, "loss_pieces.npz"))
    sm = _shift_model()
    ids = torch.from_numpy(z["ids"]).cuda()
    for n in (1, 2):
        got = sm.generate_label_mask({"input_ids": ids}, n)
        assert got.is_cuda and torch.equal(got.cpu(), torch.from_numpy(z[f"label_mask_sep{n}"]).bool())
    assert torch.equal(sm.generate_label_mask({"input_ids": ids}, 1, keep_bos=True).cpu(),
                       torch.from_numpy(z["label_mask_sep1_keep_bos"]).bool())
    with pytest.raises(ValueError):
        sm.generate_label_mask({"input_ids": ids}, 3)
    g = torch.Generator().manual_seed(5)
    for side in ("left", "right"):
        sm = _shift_model(side)
        for _ in range(20):
            ids = torch.randint(0, 4, (3, 17), generator=g)
            ids[:, 8] = 0
            ids[:, 12] = 0
            ids[:, 3] = 2
            want = OL.generate_label_mask(ids, 0, 2, padding_side=side)
            assert torch.equal(sm.generate_label_mask({"input_ids": ids.cuda()}, 2).cpu(), want)


def test_row_plans_on_cuda_bit_exact():
    """_selected_rows == nonzero() of the mask (ascending, -1 padded), _paired_rows == the k-th selected query row paired
    with the k-th selected context row (boolean-mask indexing order, src/shift_model.py:185-186)."""
    from mimic_b200.shift_model import _paired_rows, _selected_rows
    z = np.load(os.path.join(GOLDEN, "loss_pieces.npz"))
    for key in ("mask_ctx", "mask_q", "kl2_ma", "kl2_mb"):
        m = torch.from_numpy(z[key]).bool()
        rows, count = _selected_rows(m.cuda(), m.numel() + 3)
        want = m.reshape(-1).nonzero().flatten()
        assert int(count) == want.numel()
        assert torch.equal(rows.cpu()[: want.numel()].long(), want) and bool((rows.cpu()[want.numel():] == -1).all())
    g = torch.Generator().manual_seed(9)
    for _ in range(10):
        B, T, S = 3, 40, 170
        mq = torch.rand(B, T, generator=g) < 0.3
        mp_ = torch.zeros(B, S, dtype=torch.bool)
        sel = torch.randperm(B * S, generator=g)[: int(mq.sum())]
        mp_.view(-1)[sel] = True                                   # same number of selected rows on both sides
        got = _paired_rows(mq.cuda(), mp_.cuda()).cpu()
        want = torch.full((B * T,), -1, dtype=torch.int32)
        want[mq.reshape(-1)] = mp_.reshape(-1).nonzero().flatten().to(torch.int32)
        assert torch.equal(got, want)


def test_step_row_plan_matches_reference_get_hidden_states():
    """The rows kernel R gathers in the context pass and the rows the alignment loss reads in the query pass are the
    reference's masked_select rows: mask_ctx / mask_q of the fixture -> compact [L, Nq, D] tensors hp0/hp1."""
    from mimic_b200 import ops
    from mimic_b200.shift_model import _selected_rows
    z = np.load(os.path.join(GOLDEN, "loss_pieces.npz"))
    ctx = torch.from_numpy(z["ctx"]).cuda()                        # [L, B, S, D] recorded context-pass activations
    mask = torch.from_numpy(z["mask_ctx"]).bool().cuda()
    L, B, S, D = ctx.shape
    rows, count = _selected_rows(mask, B * S)
    n = int(count)
    buf = torch.zeros(L, B * S, D, device="cuda")
    for l in range(L):
        ops.record_gather(ctx[l], rows, buf[l])
    want = torch.cat([torch.from_numpy(z["hp0"]), torch.from_numpy(z["hp1"])], dim=1)     # per-sample compact tensors
    assert torch.equal(buf[:, :n].cpu(), want)


def _plan_reference(ids, pad, bos, left, num_sep, select, cap, w_scale):
    """What kernel P must produce, from the oracle's label mask and plain torch on the CPU."""
    B, S = ids.shape
    non_pad = ids != pad
    # rows with enough separators: the oracle's restatement of the reference loop (which raises on the others — the host
    # side does that from sep_counts); rows without: nothing is selected
    first0 = non_pad.long().argmax(1)
    label = torch.zeros_like(non_pad)
    for b in range(B):
        pads = (~non_pad[b]).nonzero().flatten()
        if left:
            pads = pads[pads > first0[b]]
        if num_sep >= 1 and len(pads) >= num_sep:
            label[b] = OL.generate_label_mask(ids[b:b + 1], pad, num_sep, padding_side="left" if left else "right")[0]
        elif num_sep == 0:
            label[b] = non_pad[b]
    sel = label if select == "label" else (non_pad & (ids != bos))
    flat = sel.reshape(-1)
    rows_all = flat.nonzero().flatten().to(torch.int32)
    rows = torch.full((cap,), -1, dtype=torch.int32)
    rows[: min(cap, rows_all.numel())] = rows_all[:cap]
    r2c = torch.full((B * S,), -1, dtype=torch.int32)
    r2c[flat] = torch.arange(rows_all.numel(), dtype=torch.int32)
    n_b = sel.sum(1)
    row_w = torch.zeros(cap)
    k = 0
    for b in range(B):
        for _ in range(int(n_b[b])):
            if k < cap:
                row_w[k] = w_scale / max(int(n_b[b]), 1)
            k += 1
    labels = torch.full((B, S), -100, dtype=torch.int64)
    labels[:, :-1] = torch.where(non_pad[:, 1:], ids[:, 1:], torch.full_like(ids[:, 1:], -100))
    first = non_pad.long().argmax(1)
    pos = torch.arange(S)[None]
    sep = (~non_pad) & ((pos > first[:, None]) if left else torch.ones_like(non_pad))
    return dict(attn_mask=non_pad, label_mask=label, rows=rows, row_to_compact=r2c, row_w=row_w, lm_labels=labels,
                count=int(flat.sum()), row_counts=n_b.to(torch.int32), sep_counts=sep.sum(1).to(torch.int32))


@pytest.mark.parametrize("left", [True, False], ids=["left-pad", "right-pad"])
@pytest.mark.parametrize("S", [12, 1024, 3200], ids=["S12", "S1024", "S3200"])
def test_row_plan_kernel_bit_exact(S, left):
    """mimic_row_plan against the oracle on ragged batches: leading / trailing padding, separators, an all-pad row, a row
    without separators; every output, bit for bit (integer work) — row weights to fp32 rounding."""
    from mimic_b200 import ops
    g = torch.Generator().manual_seed(S + int(left))
    B = 4
    ids = torch.randint(3, 50, (B, S), generator=g)
    for b in range(B):
        n_pad = int(torch.randint(0, max(2, S // 4), (1,), generator=g))
        if left:
            ids[b, :n_pad] = 0
            ids[b, n_pad] = 1
        else:
            ids[b, 0] = 1
            ids[b, S - n_pad:] = 0
        seps = torch.randint(2, S - 1, (3,), generator=g)
        ids[b, seps] = 0
    ids[2] = 0                                             # a row of padding only
    ids[3][ids[3] == 0] = 7                                # a row without any pad
    if left:
        ids[3, 0] = 1
    for num_sep, select in ((1, "label"), (2, "label"), (0, "non_pad_non_bos")):
        cap = B * S // 2 + 5
        want = _plan_reference(ids, 0, 1, left, num_sep, select, cap, 0.25)
        got = ops.row_plan(ids.cuda(), pad_id=0, bos_id=1, left_padding=left, num_separator=num_sep, select=select, cap=cap,
                           w_scale=0.25, want_label_mask=True, want_row_to_compact=True, want_row_w=True, want_lm_labels=True)
        for name in ("attn_mask", "label_mask", "rows", "row_to_compact", "lm_labels", "row_counts", "sep_counts"):
            assert torch.equal(getattr(got, name).cpu(), want[name]), (name, num_sep, select)
        assert int(got.count) == want["count"]
        assert torch.allclose(got.row_w.cpu(), want["row_w"], rtol=1e-6, atol=0)


def test_row_plan_pairing_matches_boolean_indexing():
    """paired[i] = the context row boolean-mask indexing pairs with query row i (src/shift_model.py:185-186), from ids."""
    from mimic_b200 import ops
    z = np.load(os.path.join(GOLDEN, "loss_pieces.npz"))
    ids = torch.from_numpy(z["ids"])                        # [2, 12] with two separators per row
    ctx = ops.row_plan(ids.cuda(), pad_id=0, bos_id=1, left_padding=True, num_separator=2, select="label", cap=ids.numel(),
                       want_label_mask=True)
    assert torch.equal(ctx.label_mask.cpu(), torch.from_numpy(z["label_mask_sep2"]).bool())
    q = ops.row_plan(ids.cuda(), pad_id=0, bos_id=1, left_padding=True, num_separator=2, select="label", pair_src=ctx.rows)
    want = torch.full((ids.numel(),), -1, dtype=torch.int32)
    m = torch.from_numpy(z["label_mask_sep2"]).bool().reshape(-1)
    want[m] = m.nonzero().flatten().to(torch.int32)
    assert torch.equal(q.paired.cpu(), want)
