"""Drop-in mirror of the reference's training module (src/shift_model.py) for the hot path:
`Strategy`, `ShiftModel.forward(...) -> loss_dict` with keys "loss", "ce_loss", "ffn_mse_loss", and
the reference's helper methods (`generate_label_mask`, `get_hidden_states`,
`calculate_layer_wise_loss`, `remove_hooks`) — with the record / alignment-loss / LM-loss
arithmetic on the CUDA kernels of libmimic_b200.so.

The Lightning shell (optimizer wiring, logging, checkpoint callbacks) is replaced by
`mimic_b200.trainer`; this class is a plain nn.Module.
"""
from __future__ import annotations

import enum
from contextlib import nullcontext
from functools import reduce
from typing import Dict, List, Optional

import torch
from torch import nn

from . import ops
from .shift_encoder import RecordPlan


class Strategy(enum.IntFlag):            # values are API: src/shift_model.py:21-25
    LAYER_WISE_MSE = 2
    LAYER_WISE_COS_SIM = 64
    LOGITS_KL_DIV = 4
    LM_LOSS = 8

    def has_layer_wise(self):
        try:
            self.layer_wise_strategy()
            return True
        except ValueError:
            return False

    def validate(self):
        layer_wise = [Strategy.LAYER_WISE_MSE, Strategy.LAYER_WISE_COS_SIM]
        if bin(self & reduce(lambda x, y: x | y, layer_wise)).count("1") > 1:
            raise ValueError(f"{[e.name for e in layer_wise]} are mutually exclusive.")

    def layer_wise_strategy(self):
        if Strategy.LAYER_WISE_MSE in self:
            return "mse_loss"
        if Strategy.LAYER_WISE_COS_SIM in self:
            return "cos_sim"
        raise ValueError("None of layer wise loss strategy is enabled")


def _selected_rows(mask: torch.Tensor, capacity: int):
    """Flat indices of the True entries of `mask` [B,S], ascending, padded with -1 up to `capacity`,
    and their count — without a host sync (stable sort instead of nonzero)."""
    flat = mask.reshape(-1)
    order = torch.sort((~flat).to(torch.uint8), stable=True).indices[:capacity]
    count = flat.sum()
    valid = torch.arange(order.numel(), device=mask.device) < count
    rows = torch.where(valid, order, torch.full_like(order, -1)).to(torch.int32)
    return rows, count


def _paired_rows(query_mask: torch.Tensor, prefix_mask: torch.Tensor) -> torch.Tensor:
    """int32 [B*T]: flat row of the context-pass tensor paired with each query row (k-th selected <-> k-th selected,
    the pairing boolean-mask indexing gives at src/shift_model.py:185-186), -1 on unselected rows.  No host sync."""
    qf = query_mask.reshape(-1)
    rows_p, _ = _selected_rows(prefix_mask, qf.numel())
    rank = (qf.long().cumsum(0) - 1).clamp_min(0)
    return torch.where(qf, rows_p[rank], torch.full_like(rows_p[rank], -1)).to(torch.int32)


class ShiftModel(nn.Module):
    def __init__(self, cfg, shift_encoder, strategy: Strategy, save_checkpoint_when=None) -> None:
        super().__init__()
        self.lmm = shift_encoder.lmm
        self.cfg = cfg
        self.shift_encoder = shift_encoder
        strategy.validate()
        self.strategy = strategy
        self.save_checkpoint_when = save_checkpoint_when
        self._hp_buffers: Dict[str, torch.Tensor] = {}

    # ---- tokenizer facts -----------------------------------------------------------------------
    @property
    def _tok(self):
        return self.lmm.processor.tokenizer

    @property
    def device(self):
        return next(self.lmm.model.parameters()).device

    @property
    def dtype(self):
        return next(self.lmm.model.parameters()).dtype

    # ---- src/shift_model.py:75-110: kernel P (one launch, no per-sample .item()) ---------------------
    def generate_label_mask(self, inputs, num_separator, keep_bos=False):
        input_ids = inputs["input_ids"]
        tok = self._tok
        left = tok.padding_side == "left"
        if keep_bos and not left:
            raise NameError("bos_position is only defined for left padding (reference behaviour)")
        plan = ops.row_plan(input_ids, pad_id=tok.pad_token_id, bos_id=getattr(tok, "bos_token_id", -1) or -1, left_padding=left,
                            num_separator=num_separator, select="label", keep_bos=keep_bos, want_attn_mask=False,
                            want_label_mask=True)
        if getattr(self, "check_separators", True):                # the reference's ValueError (97-100): one host read
            sep = plan.sep_counts.tolist()
            for i, n in enumerate(sep):
                if n < num_separator:
                    raise ValueError(f"Sequence {i} has fewer pad tokens ({n}) than num_separator ({num_separator})")
        return plan.label_mask

    def remove_hooks(self, hooks):                       # src/shift_model.py:112-119
        for name, handles in hooks.items():
            if isinstance(handles, list):
                for handle in handles:
                    handle.remove()
            else:
                handles.remove()

    # ---- reference-shaped helpers (lists of per-sample tensors), CUDA kernels underneath ---------
    def get_hidden_states(self, query_label_mask):       # src/shift_model.py:121-143
        out = {}
        counts = query_label_mask.sum(dim=1).tolist()
        rows = query_label_mask.reshape(-1).nonzero().flatten().to(torch.int32)
        for name, attr in vars(self.shift_encoder).items():
            if "hidden_states" not in name:
                continue
            L, D = len(attr), attr[0].shape[-1]
            buf = torch.empty(L, rows.numel(), D, dtype=attr[0].dtype, device=attr[0].device)
            grads = any(t.requires_grad for t in attr)
            if grads:                                    # keep autograd: differentiable gather
                for l, t in enumerate(attr):
                    buf[l] = t.reshape(-1, D).index_select(0, rows.long())
            else:
                for l, t in enumerate(attr):
                    ops.record_gather(t, rows, buf[l])
            out[name] = list(buf.split(counts, dim=1))
        if not out:
            raise RuntimeError("Layer wise loss requires to record hidden states, but no any "
                               "*_hidden_states in shift encoder.")
        return out

    def calculate_layer_wise_loss(self, shift_hidden_states, prefix_hidden_states):   # 145-179
        cos = Strategy.LAYER_WISE_MSE not in self.strategy     # validate(): exactly one of MSE / COS_SIM (src 147-160)
        if cos and Strategy.LAYER_WISE_COS_SIM not in self.strategy:
            raise ValueError("None of layer wise loss strategy is enabled")
        layer_loss = {}
        for (name, shift_list), (_, prefix_list) in zip(shift_hidden_states.items(), prefix_hidden_states.items()):
            per_sample = []
            for hs, hp in zip(shift_list, prefix_list):
                L, n, D = hs.shape
                hs_c = hs.contiguous()
                idx = torch.arange(n, dtype=torch.int32, device=hs.device)
                w = torch.full((n,), 1.0 / (L * n * (1 if cos else D)), dtype=torch.float32, device=hs.device)
                fn = ops.align_cos if cos else ops.align_mse
                per_sample.append(fn([hs_c[l] for l in range(L)], hp.contiguous(), idx, w))
            layer_loss[name.replace("hidden_states", self.strategy.layer_wise_strategy())] = \
                torch.mean(torch.stack(per_sample))
        return layer_loss

    def calculate_logits_kl_loss(self, shift_logits, prefix_logits, query_label_inputs, prefix_label_mask):   # 181-191
        """F.kl_div(log_softmax(shift_logits[mq]), softmax(prefix_logits[mp]), "batchmean") without the two masked
        copies: the k-th selected query row is paired with the k-th selected context row inside kernel KL."""
        return {"logits_kl_loss": ops.kl_loss(shift_logits, prefix_logits,
                                              _paired_rows(query_label_inputs, prefix_label_mask))}

    # ---- the hot path ------------------------------------------------------------------------------
    def forward(self, prefix_texts, query_texts, answers, images):
        """Reference signature (src/shift_model.py:193): strings + images -> loss_dict.  Builds the
        two prompts exactly like lines 204-222 and hands token tensors to `forward_inputs`."""
        tok = self._tok
        pad, eos = tok.pad_token, tok.eos_token
        query_answer = [q + pad + a + eos for q, a in zip(query_texts, answers)]
        query_images = [img[-self.cfg.data.num_image_in_query:] for img in images]
        query_inputs = self.lmm.process_input(query_images, query_answer).to(device=self.device, dtype=self.dtype)
        inputs = None
        if self.strategy != Strategy.LM_LOSS:
            full_text = [ice + pad + q + pad + a + eos for ice, q, a in zip(prefix_texts, query_texts, answers)]
            inputs = self.lmm.process_input(images, full_text).to(device=self.device, dtype=self.dtype)
        return self.forward_inputs(inputs, query_inputs)

    def _model_forward_params(self):
        import inspect
        if not hasattr(self, "_fwd_params"):
            self._fwd_params = set(inspect.signature(self.lmm.model.forward).parameters)
        return self._fwd_params

    def _hp_buffer(self, name, L, N, D, dtype, device):
        buf = self._hp_buffers.get(name)
        if buf is None or buf.shape != (L, N, D) or buf.dtype != dtype or buf.device != device:
            buf = torch.empty(L, N, D, dtype=dtype, device=device)
            self._hp_buffers[name] = buf
        return buf

    def forward_inputs(self, inputs: Optional[dict], query_inputs: dict, fold_scale: float = 1.0,
                       assume_folded: bool = False):
        """Token-level MimIC step (src/shift_model.py:200-285).  `inputs` = full-context batch
        ([bos] ICE [PAD] query [PAD] answer [eos]), `query_inputs` = query-only batch; both dicts of
        model kwargs containing "input_ids" (attention masks are rebuilt from the pad id like lines
        212/222).

        Gradient contract: the loss kernels write dH / dlogits in the same launch as the loss value, pre-scaled by
        `weight * fold_scale`.  By default (`assume_folded=False`) the autograd backward rescales them by the gradient
        that actually arrives at each loss term, so `loss_dict["loss"]` behaves like any differentiable scalar (loss
        scaling, GradScaler, extra terms all work).  `assume_folded=True` is a promise that the caller backpropagates
        exactly `fold_scale * loss_dict["loss"]` (MimicTrainer: fold_scale = 1/accumulate_grad_batches) and skips that
        rescale; any other use of it gives wrong gradients."""
        enc, tok = self.shift_encoder, self._tok
        want_kl = Strategy.LOGITS_KL_DIV in self.strategy
        cos = Strategy.LAYER_WISE_COS_SIM in self.strategy
        prefix_logits = prefix_rows_kl = None
        loss_dict = {"loss": 0.0}
        hooks = enc.register_record_hooks()
        q_ids = query_inputs["input_ids"]
        B, T = q_ids.shape
        layer_wise = self.strategy.has_layer_wise()
        record_names = [n for n in vars(enc) if "hidden_states" in n]
        N_cap = B * T
        D, L = enc.lmm_hidden_dim, enc.lmm_layers
        ids_kw = dict(pad_id=tok.pad_token_id, bos_id=tok.bos_token_id, left_padding=tok.padding_side == "left")
        # kernel P, query pass: attention mask, the query rows of the alignment loss with their compact index and weight
        # 1 / (B L [D] N_b) (lines 253-257, 145-179), the shifted LM labels — one launch
        plan_q = ops.row_plan(q_ids, select="non_pad_non_bos", cap=N_cap if layer_wise else 0,
                              w_scale=1.0 / (B * L * (1 if cos else D)), want_row_to_compact=layer_wise,
                              want_row_w=layer_wise, want_lm_labels=Strategy.LM_LOSS in self.strategy, **ids_kw)
        query_inputs = dict(query_inputs)
        query_inputs["attention_mask"] = plan_q.attn_mask

        if self.strategy != Strategy.LM_LOSS:
            inputs = dict(inputs)
            # kernel P, context pass: attention mask + the rows after the first PAD separator (generate_label_mask(inputs, 1))
            plan_c = ops.row_plan(inputs["input_ids"], num_separator=1, select="label", cap=N_cap, **ids_kw)
            inputs["attention_mask"] = plan_c.attn_mask
            if layer_wise:
                if not record_names:
                    raise RuntimeError("Layer wise loss requires to record hidden states, but no any "
                                       "*_hidden_states in shift encoder.")
                # one [L, N_cap, D] H' buffer; with two record points the hook order is attn then mlp per layer
                plans = {n: RecordPlan(plan_c.rows, self._hp_buffer(n, L, N_cap, D, self.dtype, q_ids.device))
                         for n in record_names}
                enc.record_plan = plans
            # step 1: full-context forward, no grad, adapters off (line 225-226); hooks gather H' rows
            enc.key_mask_override = inputs["attention_mask"]
            if not want_kl and "logits_to_keep" in self._model_forward_params():
                inputs.setdefault("logits_to_keep", 1)      # prefix logits are only read by LOGITS_KL_DIV
            with torch.no_grad(), getattr(self.lmm.model, "disable_adapter", nullcontext)():
                out1 = self.lmm.model(**inputs)
            enc.record_plan = None
            if want_kl:
                prefix_logits = out1["logits"] if not isinstance(out1, torch.Tensor) else out1
                # answer + [EOS] rows of the context pass (line 235), as a compact row list for the pairing below
                prefix_rows_kl = ops.row_plan(inputs["input_ids"], num_separator=2, select="label", cap=N_cap,
                                              want_attn_mask=False, **ids_kw).rows

        # step 2: shifted query forward (lines 237-243)
        hooks.update(enc.register_shift_hooks())
        enc.key_mask_override = query_inputs["attention_mask"]
        out = self.lmm.model(**query_inputs)
        enc.key_mask_override = None
        logits = out["logits"] if not isinstance(out, torch.Tensor) else out

        if Strategy.LM_LOSS in self.strategy:
            w = 1.0 if self.strategy == Strategy.LM_LOSS else float(self.cfg.ce_loss_weight)
            ce = ops.ce_loss(logits, plan_q.lm_labels, grad_scale=w * fold_scale, assume_folded=assume_folded)
            loss_dict["ce_loss"] = ce.detach()
            loss_dict["loss"] = loss_dict["loss"] + w * ce

        if layer_wise:
            aw = float(self.cfg.align_loss_weight)
            layer_loss = {}
            for n in record_names:
                hs = list(getattr(enc, n))
                align = (ops.align_cos if cos else ops.align_mse)(hs, self._hp_buffers[n], plan_q.row_to_compact, plan_q.row_w,
                                                                  grad_scale=aw * fold_scale, assume_folded=assume_folded)
                layer_loss[n.replace("hidden_states", self.strategy.layer_wise_strategy())] = align
            for k_, v_ in layer_loss.items():
                loss_dict[k_] = v_.detach()
            loss_dict["loss"] = loss_dict["loss"] + aw * sum(layer_loss.values())

        if want_kl:                                                               # step 4, lines 272-283
            aw = float(self.cfg.align_loss_weight)
            # k-th selected query row (generate_label_mask(query_inputs, 1)) <-> k-th selected context row: kernel P pairs them
            paired = ops.row_plan(q_ids, num_separator=1, select="label", want_attn_mask=False, pair_src=prefix_rows_kl,
                                  **ids_kw).paired
            kl = ops.kl_loss(logits, prefix_logits, paired, grad_scale=aw * fold_scale, assume_folded=assume_folded)
            loss_dict["logits_kl_loss"] = kl.detach()
            loss_dict["loss"] = loss_dict["loss"] + aw * kl

        self.remove_hooks(hooks)
        return loss_dict
