"""Import the UNMODIFIED reference (mercurystraw/MimIC) by path, with stub modules for its absent
dependencies.  TEST INFRASTRUCTURE ONLY; works only where /root/reference exists (the authoring
container), i.e. for `oracle/gen_golden.py` and the `needs_reference` CPU tests.  Nothing that
runs on the GPU box may call this.  Recipe: SURVEY.md Appendix C.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("MIMIC_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "src", "shift_encoder.py"))


def _load(name: str, path: str):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load_shift_encoder():
    """src/shift_encoder.py needs only torch."""
    if "ref_shift_encoder" in sys.modules:
        return sys.modules["ref_shift_encoder"]
    return _load("ref_shift_encoder", os.path.join(REF_ROOT, "src", "shift_encoder.py"))


def load_model_base():
    """testbed/models/model_base.py + testbed/utils/__init__.py, package __init__s bypassed."""
    if "testbed.models.model_base" in sys.modules:
        return sys.modules["testbed.models.model_base"]
    sys.modules.setdefault("bytecode", types.SimpleNamespace(Bytecode=None, Instr=None))
    for pkg, sub in (("testbed", "testbed"), ("testbed.models", "testbed/models")):
        if pkg not in sys.modules:
            m = types.ModuleType(pkg)
            m.__path__ = [os.path.join(REF_ROOT, sub)]
            sys.modules[pkg] = m
    _load("testbed.utils", os.path.join(REF_ROOT, "testbed", "utils", "__init__.py"))
    return _load("testbed.models.model_base", os.path.join(REF_ROOT, "testbed", "models", "model_base.py"))


def load_shift_model():
    """src/shift_model.py with pytorch_lightning / deepspeed / omegaconf / peft / paths / utils stubbed."""
    if "ref_shift_model" in sys.modules:
        return sys.modules["ref_shift_model"]
    import torch
    from torch import nn
    from transformers import get_cosine_schedule_with_warmup  # noqa: F401  (before the peft stub)

    class LightningModule(nn.Module):
        @property
        def device(self):
            return torch.device("cpu")

        @property
        def dtype(self):
            return torch.float32

    stubs = {
        "pytorch_lightning": types.SimpleNamespace(LightningModule=LightningModule),
        "deepspeed": types.ModuleType("deepspeed"),
        "deepspeed.ops": types.ModuleType("deepspeed.ops"),
        "deepspeed.ops.adam": types.SimpleNamespace(DeepSpeedCPUAdam=object),
        "omegaconf": types.SimpleNamespace(
            OmegaConf=types.SimpleNamespace(register_new_resolver=lambda *a, **k: None,
                                            to_container=lambda c, **k: c)),
        "paths": types.SimpleNamespace(result_dir="/tmp/mimic_ref_results"),
        "utils": types.SimpleNamespace(save_pretrained=lambda *a, **k: None,
                                       get_expand_runname=lambda cfg: "oracle"),
    }
    saved = {k: sys.modules.get(k) for k in stubs}
    sys.modules.update(stubs)
    try:
        mod = _load("ref_shift_model", os.path.join(REF_ROOT, "src", "shift_model.py"))
    finally:
        for k, v in saved.items():       # do not leave fake `utils` / `paths` visible to anyone else
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod


MIMIC_ATTN = "ShiftStrategy.VECTOR_SHIFT | ShiftStrategy.LEARNABLE_SHIFT_SCALE | ShiftStrategy.MULTI_HEAD"
MIMIC_FFN = "ShiftStrategy.RECORD_HIDDEN_STATES"


def make_fake_lmm(hidden: int, layers: int, heads: int, model_name: str = "idefics2-8b-base"):
    """Minimal `lmm` accepted by BaseHookEncoder.__init__ (src/shift_encoder.py:42-61)."""
    from torch import nn

    class FakeLMM(nn.Module):
        def __init__(self):
            super().__init__()
            tc = types.SimpleNamespace(hidden_size=hidden, num_hidden_layers=layers, num_attention_heads=heads)
            self.model = types.SimpleNamespace(config=types.SimpleNamespace(
                text_config=tc, hidden_size=hidden, num_hidden_layers=layers, num_attention_heads=heads))
            self.model_name = model_name

    return FakeLMM()


def make_ref_encoder(hidden: int, layers: int, heads: int, attn_strategy: str = MIMIC_ATTN,
                     ffn_strategy: str = MIMIC_FFN, model_name: str = "idefics2-8b-base", lmm=None):
    se = load_shift_encoder()
    # the strategy strings are eval'd inside the reference module's namespace (lines 30-39)
    return se.AttnApproximator(lmm if lmm is not None else make_fake_lmm(hidden, layers, heads, model_name),
                               attn_strategy=attn_strategy, ffn_strategy=ffn_strategy)
