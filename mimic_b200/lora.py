"""LoRA adapters for the MimIC + LoRA runs (scripts/run_mimic_lora.sh, src/config/peft/mini_lora.yaml: r = 1,
lora_alpha = 2 r, dropout 0.05, target `...self_attn.o_proj$`).  The reference gets them from `peft.get_peft_model`
(src/utils.py:50-63); `peft` is a third-party wrapper, so this is the small part of it the path needs, with peft's
module and parameter names (`base_layer`, `lora_A.default.weight`, `lora_B.default.weight`) so adapter checkpoints
keep their keys, and `model.disable_adapter()` for the context pass (src/shift_model.py:225).

The adapters sit on `o_proj`, i.e. directly behind the fused shifted-attention op: their gradients flow through plain
cuBLAS GEMMs, the attention backward kernel is unchanged.  Trainable adapter weights join the flat fp32 optimizer
buffer of `trainer.FlatShiftParams` through the encoder's `lmm.*` parameter names, exactly like the reference's
optimizer sees them (src/shift_model.py:338-340)."""
from __future__ import annotations

import math
import re
from contextlib import contextmanager
from typing import List

import torch
import torch.nn.functional as F
from torch import nn


class _Default(nn.Module):
    """`lora_A.default` / `lora_B.default`: peft keeps one sub-module per adapter name."""

    def __init__(self, out_features: int, in_features: int):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(out_features, in_features))


class LoraLinear(nn.Module):
    def __init__(self, base: nn.Linear, r: int, lora_alpha: float, lora_dropout: float = 0.0):
        super().__init__()
        self.base_layer = base
        base.requires_grad_(False)
        self.r, self.scaling = r, lora_alpha / r
        self.lora_dropout = nn.Dropout(lora_dropout) if lora_dropout > 0 else nn.Identity()
        self.lora_A = nn.ModuleDict({"default": _Default(r, base.in_features)})
        self.lora_B = nn.ModuleDict({"default": _Default(base.out_features, r)})
        # adapter weights are fp32 master parameters (they live in the optimizer's flat fp32 buffer); peft init:
        # A kaiming-uniform(a=sqrt(5)), B zeros
        for m in (self.lora_A["default"], self.lora_B["default"]):
            m.to(device=base.weight.device, dtype=torch.float32)
        nn.init.kaiming_uniform_(self.lora_A["default"].weight, a=math.sqrt(5))
        nn.init.zeros_(self.lora_B["default"].weight)
        self.disable_adapters = False

    @property
    def weight(self):                # code that inspects `.weight` of the wrapped projection keeps working
        return self.base_layer.weight

    @property
    def in_features(self):
        return self.base_layer.in_features

    @property
    def out_features(self):
        return self.base_layer.out_features

    def forward(self, x):
        y = self.base_layer(x)
        if self.disable_adapters:
            return y
        a = self.lora_A["default"].weight.to(x.dtype)
        b = self.lora_B["default"].weight.to(x.dtype)
        return y + F.linear(F.linear(self.lora_dropout(x), a), b) * self.scaling


def apply_lora(model: nn.Module, target_modules: str, r: int = 1, lora_alpha: float = 2.0,
               lora_dropout: float = 0.05) -> List[str]:
    """Wrap every nn.Linear whose qualified name matches the regex `target_modules` (peft's string form of
    LoraConfig.target_modules is a full-match regex; the reference's configs are suffix-anchored patterns, so `search`
    is used) and give `model` a `disable_adapter()` context manager.  Returns the wrapped module names."""
    pat = re.compile(target_modules)
    hits = [(n, m) for n, m in model.named_modules() if isinstance(m, nn.Linear) and pat.search(n)]
    if not hits:
        raise ValueError(f"Target modules {target_modules} not found in the base model.")      # peft's message
    for name, lin in hits:
        parent = model.get_submodule(name.rsplit(".", 1)[0]) if "." in name else model
        setattr(parent, name.rsplit(".", 1)[-1], LoraLinear(lin, r, lora_alpha, lora_dropout))

    @contextmanager
    def disable_adapter():
        mods = [m for m in model.modules() if isinstance(m, LoraLinear)]
        old = [m.disable_adapters for m in mods]
        for m in mods:
            m.disable_adapters = True
        try:
            yield
        finally:
            for m, o in zip(mods, old):
                m.disable_adapters = o

    model.disable_adapter = disable_adapter
    return [n for n, _ in hits]


def lora_state_dict(model: nn.Module):
    """Adapter weights only, peft key layout (what PeftModel.save_pretrained writes, src/utils.py:96-102)."""
    return {k: v.detach().clone() for k, v in model.state_dict().items() if ".lora_A." in k or ".lora_B." in k}
