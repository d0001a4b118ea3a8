"""Oracle: alignment loss, LM-loss CE, label masks, row gathering (CPU, torch).
TEST INFRASTRUCTURE ONLY.  Restates src/shift_model.py of the reference.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F


def generate_label_mask(input_ids: torch.Tensor, pad_token_id: int, num_separator: int,
                        padding_side: str = "left", keep_bos: bool = False) -> torch.Tensor:
    """ShiftModel.generate_label_mask (src/shift_model.py:75-110).

    True for non-pad tokens after the `num_separator`-th PAD *separator* of each row (with left
    padding, pads before the first non-pad token are padding, not separators).
    """
    B, S = input_ids.shape
    pad_mask = input_ids == pad_token_id
    non_pad = ~pad_mask
    label_mask = torch.zeros_like(input_ids, dtype=torch.bool)
    bos_position = non_pad.long().argmax(dim=1) if padding_side == "left" else None   # line 85
    for i in range(B):
        pads = pad_mask[i].nonzero(as_tuple=False).squeeze(-1)                        # line 88
        if padding_side == "left":
            pads = pads[pads > bos_position[i]]                                       # lines 90-93
        if len(pads) < num_separator:                                                 # lines 95-99
            raise ValueError(
                f"Sequence {i} has fewer pad tokens ({len(pads)}) than num_separator ({num_separator})")
        sep = int(pads[num_separator - 1])                                            # line 101
        label_mask[i, sep + 1:] = True                                                # line 102
    label_mask = label_mask & non_pad                                                 # line 104
    if keep_bos:                                                                      # lines 105-108
        label_mask[torch.arange(B), bos_position] = True
    return label_mask


def get_hidden_states(recorded: Sequence[torch.Tensor], mask: torch.Tensor) -> List[torch.Tensor]:
    """ShiftModel.get_hidden_states for one record list (src/shift_model.py:121-143).

    recorded: L tensors [B,S,D]; mask [B,S] bool.  Returns B tensors [L, Nq_i, D].
    """
    hs = torch.stack(list(recorded)).transpose(0, 1)          # [B,L,S,D]   line 131
    B, L, S, D = hs.shape
    return [h.masked_select(m[None, :, None]).view(L, -1, D) for h, m in zip(hs, mask)]   # 133-136


def layer_wise_mse(shift_hs: Sequence[torch.Tensor], prefix_hs: Sequence[torch.Tensor]) -> torch.Tensor:
    """Strategy.LAYER_WISE_MSE branch of calculate_layer_wise_loss (src/shift_model.py:145-179):
    mean over samples of mean_{l,t,d} (H_i - H'_i)^2."""
    return torch.mean(torch.stack([F.mse_loss(a, b, reduction="mean") for a, b in zip(shift_hs, prefix_hs)]))


def layer_wise_cos_sim(shift_hs, prefix_hs) -> torch.Tensor:
    """Strategy.LAYER_WISE_COS_SIM branch (src/shift_model.py:152-160, 171-178).  The reference's
    per-sample value is a length-L vector (mean over tokens only); stack+mean averages it all."""
    return torch.mean(torch.stack(
        [1 - torch.mean(F.cosine_similarity(a, b, dim=-1), dim=1) for a, b in zip(shift_hs, prefix_hs)]))


def lm_ce_loss(logits: torch.Tensor, labels: torch.Tensor, attention_mask: Optional[torch.Tensor],
               ignore_index: int = -100) -> torch.Tensor:
    """LM loss computed inside the HF model when `labels=` is passed (call site
    src/shift_model.py:238-250, `labels=input_ids`, no -100 masking by the caller).

    The arithmetic lives in third-party `transformers==4.45.2` (requirements.txt:21), which is not
    in the reference tree and not installed (installed: 5.5.0, whose rule differs).  PARITY UNPINNED
    at this boundary.  The published 4.45-era rule for Idefics/Idefics2/Llava is restated here:

        shift_logits = logits[..., :-1, :], shift_labels = labels[..., 1:]
        keep position t iff attention_mask[..., 1:][t] != 0 (and label != ignore_index)
        loss = mean over kept positions of CE(fp32 logits, label)
    """
    B, T, V = logits.shape
    sl = logits[:, :-1, :].float().reshape(-1, V)
    tl = labels[:, 1:].reshape(-1).clone()
    if attention_mask is not None:
        keep = attention_mask[:, 1:].reshape(-1) != 0
        tl[~keep] = ignore_index
    return F.cross_entropy(sl, tl, ignore_index=ignore_index, reduction="mean")


def logits_kl(shift_logits_rows: torch.Tensor, prefix_logits_rows: torch.Tensor) -> torch.Tensor:
    """calculate_logits_kl_loss (src/shift_model.py:181-191) on already-selected rows [N,V]."""
    return F.kl_div(shift_logits_rows.log_softmax(-1), prefix_logits_rows.softmax(-1),
                    reduction="batchmean", log_target=False)


def mimic_loss(ce: Optional[torch.Tensor], layer_losses: Dict[str, torch.Tensor],
               ce_loss_weight: float = 0.5, align_loss_weight: float = 1.0, lm_only: bool = False):
    """Loss composition of ShiftModel.forward (src/shift_model.py:245-250, 265-270)."""
    loss = 0.0
    if ce is not None:
        loss = loss + (1.0 if lm_only else ce_loss_weight) * ce
    if layer_losses:
        loss = loss + align_loss_weight * sum(layer_losses.values())
    return loss
