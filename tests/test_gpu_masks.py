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
