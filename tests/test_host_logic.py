"""CPU tests: the C-ABI library loads and exports every symbol include/mimic_b200.h declares, and
the host-side mirror of the reference interface (strategy parsing, hook API, label masks, state
dict keys, error behaviour) matches the reference / oracle.  No GPU compute calls."""
import os
import re
import types

import numpy as np
import pytest
import torch
from torch import nn

from oracle import losses as OL
from oracle import ref_loader

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MIMIC = ("ShiftStrategy.VECTOR_SHIFT | ShiftStrategy.LEARNABLE_SHIFT_SCALE | ShiftStrategy.MULTI_HEAD",
         "ShiftStrategy.RECORD_HIDDEN_STATES")


def test_library_exports_every_declared_symbol():
    import ctypes
    from mimic_b200 import _lib
    header = open(os.path.join(ROOT, "include", "mimic_b200.h")).read()
    declared = set(re.findall(r"MIMIC_API\s+[\w\s\*]+?\b(mimic_\w+)\s*\(", header))
    assert len(declared) >= 11
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    assert lib.mimic_abi_version() == _lib.ABI_VERSION == 2


def test_ops_refuse_cpu_tensors():
    from mimic_b200 import ops
    from mimic_b200._lib import MimicError
    t = torch.zeros(1, 1, 8, 32)
    with pytest.raises(MimicError):
        ops.shifted_attention(t, t, t, active=False)
    with pytest.raises(MimicError):
        ops.ce_loss(torch.zeros(2, 5), torch.zeros(2, dtype=torch.long))


def _tiny_lmm(name="llava-interleave-qwen-7b-hf", layers=2):
    from mimic_b200 import HookedModel

    class Attn(nn.Module):
        def forward(self, x, **kw):
            return x * 2, None

    class Layer(nn.Module):
        def __init__(self):
            super().__init__()
            self.self_attn, self.mlp = Attn(), nn.Linear(8, 8)

        def forward(self, x):
            return self.mlp(self.self_attn(x)[0])

    class Inner(nn.Module):
        def __init__(self):
            super().__init__()
            self.layers = nn.ModuleList(Layer() for _ in range(layers))

        def forward(self, x):
            for l in self.layers:
                x = l(x)
            return x

    class LM(nn.Module):
        def __init__(self):
            super().__init__()
            self.model = Inner()

    class Top(nn.Module):
        def __init__(self):
            super().__init__()
            self.language_model = LM()
            tc = types.SimpleNamespace(hidden_size=8, num_hidden_layers=layers, num_attention_heads=2)
            self.config = types.SimpleNamespace(text_config=tc)

        def forward(self, x):
            return self.language_model.model(x)

    return HookedModel(Top(), name)


def test_strategy_parsing_matches_reference_semantics():
    from mimic_b200.shift_encoder import ShiftStrategy, parse_strategy_flags
    assert parse_strategy_flags(MIMIC[0]) == ShiftStrategy(1 | 8 | 16)
    assert parse_strategy_flags("null") == ShiftStrategy(0) and parse_strategy_flags(None) == ShiftStrategy(0)
    assert parse_strategy_flags(ShiftStrategy.VECTOR_SHIFT) == ShiftStrategy.VECTOR_SHIFT
    assert parse_strategy_flags("ShiftStrategy(0)") == ShiftStrategy(0)
    with pytest.raises(AttributeError):           # reference: eval raises AttributeError too (attn_shift_ffn_mse.yaml:5)
        parse_strategy_flags("ShiftStrategy.LEARNABLE_SCALE")
    assert [int(x) for x in (ShiftStrategy.VECTOR_SHIFT, ShiftStrategy.RECORD_HIDDEN_STATES,
                             ShiftStrategy.LEARNABLE_SHIFT_SCALE, ShiftStrategy.MULTI_HEAD)] == [1, 4, 8, 16]


def test_encoder_contract_state_dict_and_errors():
    from mimic_b200 import AttnApproximator, Strategy
    lmm = _tiny_lmm()
    enc = AttnApproximator(lmm, *MIMIC)
    keys = {k for k in enc.state_dict() if not k.startswith("lmm")}
    assert keys == {"attn_shift", "log_Z1_lin.0.weight", "log_Z1_lin.0.bias", "log_Z1_lin.1.weight", "log_Z1_lin.1.bias"}
    assert enc.attn_shift.shape == (2, 2, 4) and enc.log_Z1_lin[0].weight.shape == (2, 4)
    assert isinstance(enc.ffn_hidden_states, list) and len(enc.ffn_hidden_states) == 2 and not hasattr(enc, "attn_hidden_states")
    with pytest.raises(ValueError):
        AttnApproximator(_tiny_lmm("gpt2"), *MIMIC)
    with pytest.raises(ValueError):
        AttnApproximator(lmm, "ShiftStrategy.LEARNABLE_SHIFT_SCALE", None)
    assert [int(s) for s in (Strategy.LAYER_WISE_MSE, Strategy.LOGITS_KL_DIV, Strategy.LM_LOSS, Strategy.LAYER_WISE_COS_SIM)] == [2, 4, 8, 64]
    with pytest.raises(ValueError):
        (Strategy.LAYER_WISE_MSE | Strategy.LAYER_WISE_COS_SIM).validate()


@pytest.mark.needs_reference
def test_state_dict_keys_equal_reference():
    from mimic_b200 import AttnApproximator
    ref = ref_loader.make_ref_encoder(8, 2, 2)
    mine = AttnApproximator(_tiny_lmm(), *MIMIC)
    rk = {k: tuple(v.shape) for k, v in ref.state_dict().items() if not k.startswith("lmm")}
    mk = {k: tuple(v.shape) for k, v in mine.state_dict().items() if not k.startswith("lmm")}
    assert rk == mk
    mine.load_state_dict({k: v for k, v in ref.state_dict().items() if not k.startswith("lmm")}, strict=False)


def test_hook_api_record_and_replace():
    from mimic_b200 import AttnApproximator
    lmm = _tiny_lmm()
    enc = AttnApproximator(lmm, *MIMIC)
    hooks = enc.register_record_hooks()
    assert list(hooks) == ["ffn_record_hook"] and len(hooks["ffn_record_hook"]) == 2
    x = torch.randn(1, 3, 8)
    lmm.model(x)
    assert all(isinstance(t, torch.Tensor) and t.shape == (1, 3, 8) for t in enc.ffn_hidden_states)
    for h in hooks["ffn_record_hook"]:
        h.remove()
    # replace_module_method injects module_name / old_method like the reference (model_base.py:730-739)
    seen = {}

    def new_forward(self, x, module_name=None, old_method=None, **kw):
        seen[module_name] = old_method
        return old_method(x)

    lmm.replace_module_method(enc.decoder_self_attn_name, "forward", new_forward, strict=False)
    lmm.model(x)
    assert sorted(seen) == ["language_model.model.layers.0.self_attn", "language_model.model.layers.1.self_attn"]
    with pytest.raises(ValueError):
        lmm.replace_module_method(r"no_such_module$", "forward", new_forward)
    reg = enc.register_shift_hooks()
    assert list(reg) == ["attn_hook"] and all(h.active for h in reg["attn_hook"])
    for h in reg["attn_hook"]:
        h.remove()
    assert not any(h.active for h in enc.attn_shift_handles)


def test_generate_label_mask_is_cuda_only():
    """The mask / row-plan arithmetic is kernel P (mimic_row_plan): like every op it refuses CPU tensors — no fallback.
    (Bit-exact parity with the reference-made masks: tests/test_gpu_masks.py; the oracle restatement against the same
    fixture: tests/test_oracle_golden.py.)"""
    from mimic_b200 import AttnApproximator, ShiftModel, Strategy
    from mimic_b200._lib import MimicError
    lmm = _tiny_lmm()
    lmm.processor = types.SimpleNamespace(tokenizer=types.SimpleNamespace(pad_token_id=0, bos_token_id=1, padding_side="left"))
    sm = ShiftModel(types.SimpleNamespace(), AttnApproximator(lmm, *MIMIC), Strategy.LM_LOSS | Strategy.LAYER_WISE_MSE)
    with pytest.raises(MimicError):
        sm.generate_label_mask({"input_ids": torch.zeros(2, 8, dtype=torch.int64)}, 1)


def test_selected_rows_without_sync():
    from mimic_b200.shift_model import _selected_rows
    m = torch.tensor([[0, 1, 1, 0], [1, 0, 0, 1]], dtype=torch.bool)
    rows, count = _selected_rows(m, 6)
    assert rows.tolist() == [1, 2, 4, 7, -1, -1] and int(count) == 4


def test_key_mask_from_4d():
    from mimic_b200.shift_encoder import key_mask_from_4d
    B, T, S = 2, 3, 5
    km = torch.tensor([[0, 1, 1, 1, 1], [1, 1, 0, 1, 1]], dtype=torch.bool)
    causal = torch.tril(torch.ones(T, S, dtype=torch.bool), diagonal=S - T)
    m4 = (km[:, None, None, :] & causal[None, None])
    c, k = key_mask_from_4d(m4, T, S)
    assert c and torch.equal(k, km)
    add = torch.zeros(B, 1, T, S).masked_fill(~m4, torch.finfo(torch.float32).min)
    c, k = key_mask_from_4d(add, T, S)
    assert c and torch.equal(k, km)
    assert key_mask_from_4d(None, 1, S) == (False, None)


def test_lora_wrapper_semantics_cpu():
    """mimic_b200.lora on CPU tensors (plain torch modules, no kernel involved): peft's forward, names and
    disable_adapter()."""
    import torch
    from torch import nn
    from mimic_b200.lora import LoraLinear, apply_lora, lora_state_dict

    class Blk(nn.Module):
        def __init__(self):
            super().__init__()
            self.self_attn = nn.Module()
            self.self_attn.o_proj = nn.Linear(8, 8, bias=False)
            self.self_attn.q_proj = nn.Linear(8, 8, bias=False)

    class M(nn.Module):
        def __init__(self):
            super().__init__()
            self.layers = nn.ModuleList([Blk(), Blk()])

    torch.manual_seed(0)
    m = M()
    names = apply_lora(m, r"layers\.\d+\.self_attn\.o_proj$", r=1, lora_alpha=2.0, lora_dropout=0.0)
    assert names == ["layers.0.self_attn.o_proj", "layers.1.self_attn.o_proj"]
    lo = m.layers[0].self_attn.o_proj
    assert isinstance(lo, LoraLinear) and isinstance(m.layers[0].self_attn.q_proj, nn.Linear)
    x = torch.randn(3, 8)
    assert torch.equal(lo(x), lo.base_layer(x))                      # B starts at zero (peft init)
    with torch.no_grad():
        lo.lora_B["default"].weight.normal_()
    want = lo.base_layer(x) + (x @ lo.lora_A["default"].weight.t() @ lo.lora_B["default"].weight.t()) * 2.0
    assert torch.allclose(lo(x), want, atol=1e-6)
    with m.disable_adapter():
        assert torch.equal(lo(x), lo.base_layer(x))
    assert torch.allclose(lo(x), want, atol=1e-6)
    assert sorted(lora_state_dict(m)) == sorted(f"layers.{i}.self_attn.o_proj.lora_{ab}.default.weight"
                                                for i in (0, 1) for ab in "AB")
    trainable = [n for n, p in m.named_parameters() if p.requires_grad]
    assert all(".lora_" in n for n in trainable if "o_proj" in n) and not lo.base_layer.weight.requires_grad
    import pytest
    with pytest.raises(ValueError):
        apply_lora(M(), r"no_such_module$")


def test_header_is_plain_c_and_links_without_torch(tmp_path):
    """include/mimic_b200.h must be consumable from C (what a cgo / JNI / N-API stub would include) and the library must
    link into a program that knows nothing about torch: compile a C translation unit against the header, link it with
    libmimic_b200.so, run it (no GPU needed: it only asks for the ABI version and an error string)."""
    import os
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        import pytest
        pytest.skip("gcc not available")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    import mimic_b200  # noqa: F401  makes sure the library is built
    src = tmp_path / "abi.c"
    src.write_text('''
#include <stdio.h>
#include <string.h>
#include "mimic_b200.h"
int main(void) {
  mimic_attn_desc_t d;
  memset(&d, 0, sizeof d);
  /* a NULL descriptor / NULL pointers must come back as an error status with a message, not a crash */
  int rc = mimic_shifted_attn_fwd(&d, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0);
  printf("%d %d %s\\n", mimic_abi_version(), rc != 0, mimic_last_error());
  return (mimic_abi_version() == 2 && rc != 0 && strlen(mimic_last_error()) > 0) ? 0 : 1;
}
''')
    exe = tmp_path / "abi"
    libdir = os.path.join(root, "mimic_b200")
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(root, "include"), str(src), "-o", str(exe),
                        "-L", libdir, "-l:libmimic_b200.so", f"-Wl,-rpath,{libdir}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    run = subprocess.run([str(exe)], capture_output=True, text=True)
    assert run.returncode == 0, run.stdout + run.stderr


def test_flat_params_scale_lr_groups():
    """FlatShiftParams: weight decay 0 for *bias* names, lr multiplier scale_lr / lr for *log_Z1* / *scale* names
    (src/shift_model.py:306-357), for both encoder classes."""
    from mimic_b200 import AttnApproximator, AttnFFNShift
    from mimic_b200.trainer import FlatShiftParams
    from oracle.ref_loader import make_fake_lmm
    enc = AttnFFNShift(make_fake_lmm(64, 3, 4), None, "ShiftStrategy.VECTOR_SHIFT | ShiftStrategy.LEARNABLE_SHIFT_SCALE",
                       shift_scale_init_value=0.1)
    assert {n for n, _ in enc.named_parameters()} == {"ffn_shift", "ffn_shift_scale"}
    assert torch.allclose(enc.ffn_shift_scale, torch.full((3,), 0.1))
    flat = FlatShiftParams(enc, 1e-3, lr=5e-3, scale_lr=1e-4)
    assert flat.numel == 3 * 64 + 3
    by_name = dict(zip(flat.names, torch.split(flat.lr_mul, [enc.get_parameter(n).numel() for n in flat.names])))
    assert torch.all(by_name["ffn_shift"] == 1.0) and torch.allclose(by_name["ffn_shift_scale"], torch.tensor(0.02))
    assert FlatShiftParams(enc, 1e-3).lr_mul is None
    enc2 = AttnApproximator(make_fake_lmm(64, 2, 4),
                            "ShiftStrategy.VECTOR_SHIFT | ShiftStrategy.LEARNABLE_SHIFT_SCALE | ShiftStrategy.MULTI_HEAD", None)
    flat2 = FlatShiftParams(enc2, 1e-3, lr=1e-2, scale_lr=1e-3)
    sizes = [enc2.get_parameter(n).numel() for n in flat2.names]
    for n, mul, wd in zip(flat2.names, torch.split(flat2.lr_mul, sizes), torch.split(flat2.weight_decay, sizes)):
        assert torch.allclose(mul, torch.tensor(0.1 if "log_Z1" in n else 1.0)), n
        assert torch.all(wd == (0.0 if "bias" in n else 1e-3)), n
    # the other AttnFFNShift shapes: non-learnable scale is a buffer, MULTI_HEAD is refused like the reference does
    enc3 = AttnFFNShift(make_fake_lmm(64, 3, 4), "ShiftStrategy.VECTOR_SHIFT", None)
    assert {n for n, _ in enc3.named_parameters()} == {"attn_shift"} and "attn_shift_scale" in dict(enc3.named_buffers())
    with pytest.raises(ValueError):
        AttnFFNShift(make_fake_lmm(64, 3, 4), "ShiftStrategy.VECTOR_SHIFT | ShiftStrategy.MULTI_HEAD", None)
    assert AttnFFNShift(make_fake_lmm(64, 3, 4), None, None).register_shift_hooks() == {}


def test_every_c_symbol_is_documented_in_integration_md():
    """include/mimic_b200.h is the drop-in boundary: each entry point must appear in INTEGRATION.md's table of what it
    replaces in the reference, and in the ctypes binding table (mimic_b200/_lib.py)."""
    import re
    from mimic_b200 import _lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = open(os.path.join(root, "include", "mimic_b200.h")).read()
    names = set(re.findall(r"MIMIC_API\s+[\w\s\*]+?\b(mimic_\w+)\s*\(", header))
    assert len(names) >= 20
    doc = open(os.path.join(root, "INTEGRATION.md")).read()
    assert not [n for n in sorted(names) if n not in doc]
    assert names == set(_lib.EXPORTS)


def test_graft_entry_build_runs():
    """`__graft_entry__.build()` (the driver's "does it build" check) compiles / loads the library and agrees with the
    front-end on the ABI version — it asserted a stale version number once."""
    import importlib
    ge = importlib.import_module("__graft_entry__")
    ge.build()


def test_grad_like_keeps_row_stride_and_alignment():
    """Gradient buffers of the loss ops: the row stride of the logits view and its address modulo 32 bytes (the kernels
    address logits and gradient with ONE stride and align their 16-byte chunks on the logits row)."""
    import torch
    from mimic_b200.ops import _grad_like
    wide = torch.zeros(7, 50, dtype=torch.bfloat16)
    for view in (wide, wide[:, :37], wide[1:, 3:40], wide[2:, 5:6]):
        g = _grad_like(view)
        assert g.shape == view.shape and g.dtype == view.dtype
        if view.shape[0] > 1:
            assert g.stride(0) == view.stride(0)
        assert g.stride(1) == 1 and g.data_ptr() % 32 == view.data_ptr() % 32
        g.fill_(1)                                           # every element addressable inside its own storage
        assert int(g.float().sum()) == view.numel()
