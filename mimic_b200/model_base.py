"""The slice of the reference's model-wrapper boundary that the hot path is reached through:
regex/list-addressed hook registration and forward-method replacement
(testbed/models/model_base.py:65-88, 169-322, 649-739; testbed/utils/__init__.py:11-49).

`HookedModel` wraps ANY nn.Module tree (an HF model or the synthetic backbones of
`mimic_b200.backbone`) — it does not load checkpoints or processors; those stay with the caller.
A reference `ModelBase` instance can be passed to `AttnApproximator` instead: the encoder only
uses the methods mirrored here.
"""
from __future__ import annotations

import inspect
import re
from functools import partial
from types import MethodType
from typing import Callable, Dict, List, Union

import torch.nn as nn


def try_inject_params(fn: Callable, **kwargs) -> Callable:
    """Bind those of `kwargs` that `fn` can take (all of them if it has **kwargs); see
    testbed/utils/__init__.py:11-49."""
    params = inspect.signature(fn).parameters
    if any(p.kind == inspect.Parameter.VAR_KEYWORD for p in params.values()):
        chosen = kwargs
    else:
        chosen = {k: v for k, v in kwargs.items() if k in params}
    return partial(fn, **chosen) if chosen else fn


class HookedModel(nn.Module):
    def __init__(self, model: nn.Module, model_name: str, processor=None):
        super().__init__()
        self.model = model
        self.model_name = model_name
        self.processor = processor
        self.config = getattr(model, "config", None)

    # -- module addressing ----------------------------------------------------------------
    def _match(self, module_name: Union[str, List[str], type]) -> Dict[str, nn.Module]:
        named = self.model.named_modules()
        if isinstance(module_name, str):
            return {n: m for n, m in named if re.search(module_name, n)}
        if isinstance(module_name, list):
            return {n: m for n, m in named if n in module_name}
        if isinstance(module_name, type):
            return {n: m for n, m in named if isinstance(m, module_name)}
        raise TypeError(f"module_name should be str or list of str, but got {type(module_name).__name__}")

    def _register_hook(self, register_fn_name, module_name, hook, **kwargs):
        if not isinstance(module_name, (str, list)):
            raise TypeError(f"module_name should be str or list of str, but got {type(module_name).__name__}")
        return [getattr(m, register_fn_name)(try_inject_params(hook, module_name=n), **kwargs)
                for n, m in self._match(module_name).items()]

    def register_forward_hook(self, module_name, hook, **kwargs):
        return self._register_hook("register_forward_hook", module_name, hook, **kwargs)

    def register_forward_pre_hook(self, module_name, hook, **kwargs):
        return self._register_hook("register_forward_pre_hook", module_name, hook, **kwargs)

    def register_full_backward_hook(self, module_name, hook, **kwargs):
        return self._register_hook("register_full_backward_hook", module_name, hook, **kwargs)

    def replace_module_method(self, module_name_or_type, method_name: str, new_method: Callable, *,
                              strict: bool = True):
        matched = self._match(module_name_or_type)
        if not matched:
            raise ValueError(f"No modules found matching {module_name_or_type}")
        for name, module in matched.items():
            old = getattr(module, method_name, None)
            if old is None:
                raise ValueError(f"Module '{name}' has no method '{method_name}'")
            if strict:
                orig = list(inspect.signature(old).parameters)
                new = list(inspect.signature(new_method).parameters)
                if "self" not in orig and isinstance(old, MethodType):
                    orig.insert(0, "self")
                if orig != new[: len(orig)]:
                    raise ValueError(
                        f"The first few parameters of the new function do not match the original method "
                        f"'{method_name}' of module '{name}'; set strict=False to skip this check.")
            setattr(module, method_name,
                    MethodType(try_inject_params(new_method, module_name=name, old_method=old), module))

    def forward(self, *args, **kwargs):
        return self.model(*args, **kwargs)
