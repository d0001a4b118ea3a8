"""BASELINE config 1: the full MimIC step on the tiny random-init decoder, CUDA path vs what the
REFERENCE produced through its own hooks (tests/golden/tiny_step.npz, made by oracle/gen_golden.py)."""
import os
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel(a, b):
    a, b = a.detach().double().cpu(), torch.as_tensor(b).double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def _tiny_setup(z, strategy):
    from oracle.gen_golden import _Tok, tiny_qwen2
    from mimic_b200 import AttnApproximator, HookedModel, ShiftModel
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    model = tiny_qwen2()
    lmm = HookedModel(model.cuda(), "llava-interleave-qwen-7b-hf", processor=types.SimpleNamespace(tokenizer=_Tok()))
    enc = AttnApproximator(lmm, "ShiftStrategy.VECTOR_SHIFT | ShiftStrategy.LEARNABLE_SHIFT_SCALE | ShiftStrategy.MULTI_HEAD",
                           "ShiftStrategy.RECORD_HIDDEN_STATES").cuda()
    sd = {k[5:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("init.")}
    enc.load_state_dict(sd, strict=False)
    cfg = types.SimpleNamespace(ce_loss_weight=0.5, align_loss_weight=1.0,
                                data=types.SimpleNamespace(num_image_in_query=1))
    return enc, ShiftModel(cfg, enc, strategy)


def test_tiny_step_cos_sim_and_logits_kl_match_reference():
    """Same step under Strategy.LM_LOSS | LAYER_WISE_COS_SIM | LOGITS_KL_DIV: loss terms and the
    shift-parameter gradients against what the reference produced (tiny_step.npz, keys cosk.*)."""
    from mimic_b200 import Strategy
    z = np.load(os.path.join(GOLDEN, "tiny_step.npz"))
    enc, sm = _tiny_setup(z, Strategy.LM_LOSS | Strategy.LAYER_WISE_COS_SIM | Strategy.LOGITS_KL_DIV)
    full_ids, qa_ids = torch.from_numpy(z["full_ids"]).cuda(), torch.from_numpy(z["qa_ids"]).cuda()
    loss_dict = sm.forward_inputs({"input_ids": full_ids}, {"input_ids": qa_ids})
    assert set(loss_dict) == {"loss", "ce_loss", "ffn_cos_sim", "logits_kl_loss"}
    assert abs(float(loss_dict["ffn_cos_sim"]) - float(z["cosk.ffn_cos_sim"])) < 1e-4 * float(z["cosk.ffn_cos_sim"])
    assert abs(float(loss_dict["logits_kl_loss"]) - float(z["cosk.logits_kl_loss"])) < 1e-3 * float(z["cosk.logits_kl_loss"])
    assert abs(float(loss_dict["loss"]) - float(z["cosk.loss"])) < 2e-5 * float(z["cosk.loss"])
    params = {n: p for n, p in enc.named_parameters() if p.requires_grad and not n.startswith("lmm")}
    grads = torch.autograd.grad(loss_dict["loss"], list(params.values()))
    for (n, p), g in zip(params.items(), grads):
        assert rel(g, z["cosk.grad." + n]) < 5e-4, n


def test_tiny_step_with_lora_matches_reference():
    """BASELINE config 4 in miniature (MimIC + LoRA r=1 on self_attn.o_proj, context pass with adapters disabled):
    losses, shift-parameter gradients and adapter gradients against the reference-made fixture, then the adapters
    join the flat optimizer buffer of the on-device trainer."""
    from oracle.gen_golden import _Tok, tiny_qwen2
    from mimic_b200 import AttnApproximator, HookedModel, ShiftModel, Strategy
    from mimic_b200.lora import LoraLinear, apply_lora, lora_state_dict
    from mimic_b200.trainer import FlatShiftParams
    z = np.load(os.path.join(GOLDEN, "tiny_step_lora.npz"))
    torch.backends.cuda.matmul.allow_tf32 = False
    model = tiny_qwen2().cuda()
    names = apply_lora(model, r"language_model\.model\.layers\.\d+\.self_attn\.o_proj$", r=1, lora_alpha=2.0, lora_dropout=0.0)
    assert len(names) == 4
    with torch.no_grad():
        for n in names:
            m = model.get_submodule(n)
            assert isinstance(m, LoraLinear) and not m.base_layer.weight.requires_grad
            m.lora_A["default"].weight.copy_(torch.from_numpy(z[f"lora.{n}.A"]))
            m.lora_B["default"].weight.copy_(torch.from_numpy(z[f"lora.{n}.B"]))
    lmm = HookedModel(model, "llava-interleave-qwen-7b-hf", processor=types.SimpleNamespace(tokenizer=_Tok()))
    enc = AttnApproximator(lmm, "ShiftStrategy.VECTOR_SHIFT | ShiftStrategy.LEARNABLE_SHIFT_SCALE | ShiftStrategy.MULTI_HEAD",
                           "ShiftStrategy.RECORD_HIDDEN_STATES").cuda()
    enc.load_state_dict({k[5:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("init.")}, strict=False)
    cfg = types.SimpleNamespace(ce_loss_weight=0.5, align_loss_weight=1.0,
                                data=types.SimpleNamespace(num_image_in_query=1))
    sm = ShiftModel(cfg, enc, Strategy.LM_LOSS | Strategy.LAYER_WISE_MSE)
    full_ids, qa_ids = torch.from_numpy(z["full_ids"]).cuda(), torch.from_numpy(z["qa_ids"]).cuda()
    loss_dict = sm.forward_inputs({"input_ids": full_ids}, {"input_ids": qa_ids})
    assert abs(float(loss_dict["ce_loss"]) - float(z["ce_loss"])) < 2e-5 * float(z["ce_loss"])
    assert abs(float(loss_dict["ffn_mse_loss"]) - float(z["ffn_mse_loss"])) < 1e-4 * float(z["ffn_mse_loss"])
    assert abs(float(loss_dict["loss"]) - float(z["loss"])) < 2e-5 * float(z["loss"])
    params = {n: p for n, p in enc.named_parameters() if p.requires_grad}
    shift_names = [n for n in params if not n.startswith("lmm")]
    lora_names = [n for n in params if n.startswith("lmm")]
    assert len(lora_names) == 8                                   # A and B of four o_proj, reached through enc.lmm
    grads = dict(zip(params, torch.autograd.grad(loss_dict["loss"], list(params.values()))))
    for n in shift_names:
        assert rel(grads[n], z["grad." + n]) < 2e-4, n
    for n in names:
        assert rel(grads[f"lmm.model.{n}.lora_A.default.weight"], z[f"grad.{n}.A"]) < 2e-4, n
        assert rel(grads[f"lmm.model.{n}.lora_B.default.weight"], z[f"grad.{n}.B"]) < 2e-4, n
    assert set(k.split(".lora_")[1][0] for k in lora_state_dict(model)) == {"A", "B"}
    # the adapters are optimizer state like the shift parameters (src/shift_model.py:338-340)
    flat = FlatShiftParams(enc, weight_decay=1e-3)
    assert flat.numel == sum(p.numel() for p in params.values())
    assert all(n in flat.views for n in lora_names)


def test_tiny_step_matches_reference():
    from oracle.gen_golden import _Tok, tiny_qwen2
    from mimic_b200 import AttnApproximator, HookedModel, ShiftModel, Strategy, ops
    z = np.load(os.path.join(GOLDEN, "tiny_step.npz"))
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    model = tiny_qwen2()
    wsum = sum(float(p.double().abs().sum()) for p in model.parameters())
    assert abs(wsum - float(z["weights_checksum"])) < 1e-6 * wsum, "random-init backbone differs from the fixture's"
    lmm = HookedModel(model.cuda(), "llava-interleave-qwen-7b-hf", processor=types.SimpleNamespace(tokenizer=_Tok()))
    enc = AttnApproximator(lmm, "ShiftStrategy.VECTOR_SHIFT | ShiftStrategy.LEARNABLE_SHIFT_SCALE | ShiftStrategy.MULTI_HEAD",
                           "ShiftStrategy.RECORD_HIDDEN_STATES").cuda()
    sd = {k[5:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("init.")}
    missing = enc.load_state_dict(sd, strict=False)
    assert not [k for k in missing.missing_keys if not k.startswith("lmm")]
    cfg = types.SimpleNamespace(ce_loss_weight=0.5, align_loss_weight=1.0,
                                data=types.SimpleNamespace(num_image_in_query=1))
    sm = ShiftModel(cfg, enc, Strategy.LM_LOSS | Strategy.LAYER_WISE_MSE)
    full_ids, qa_ids = torch.from_numpy(z["full_ids"]).cuda(), torch.from_numpy(z["qa_ids"]).cuda()
    n0 = ops.launch_count
    loss_dict = sm.forward_inputs({"input_ids": full_ids}, {"input_ids": qa_ids})
    assert set(loss_dict) == {"loss", "ce_loss", "ffn_mse_loss"}
    assert abs(float(loss_dict["ce_loss"]) - float(z["ce_loss"])) < 2e-5 * float(z["ce_loss"])
    assert abs(float(loss_dict["ffn_mse_loss"]) - float(z["ffn_mse_loss"])) < 1e-4 * float(z["ffn_mse_loss"])
    assert abs(float(loss_dict["loss"]) - float(z["loss"])) < 2e-5 * float(z["loss"])
    params = {n: p for n, p in enc.named_parameters() if p.requires_grad and not n.startswith("lmm")}
    grads = torch.autograd.grad(loss_dict["loss"], list(params.values()))
    # first step: the context pass still runs the stock HF attention (forward is replaced at
    # register_shift_hooks, src/shift_encoder.py:639-645): 4 gathers + 4 fused fwd + mse + ce + 4 bwd(3)
    assert ops.launch_count - n0 >= 4 + 4 + 1 + 1 + 4 * 3
    for (n, p), g in zip(params.items(), grads):
        assert rel(g, z["grad." + n]) < 2e-4, n
    # handles are inactive again, hooks removed (src/shift_model.py:262)
    assert not any(h.active for h in enc.attn_shift_handles)


def test_tiny_step_licv_matches_reference():
    """The LIVE baseline in miniature (src/config/encoder/licv.yaml: AttnFFNShift, learnable norm-preserving FFN shift,
    Strategy.LM_LOSS | LOGITS_KL_DIV): loss terms and the gradients of ffn_shift / ffn_shift_scale against what the
    reference produced through its own hooks (tests/golden/tiny_step_licv.npz)."""
    from oracle.gen_golden import _Tok, tiny_qwen2
    from mimic_b200 import AttnFFNShift, HookedModel, ShiftModel, Strategy, ops
    z = np.load(os.path.join(GOLDEN, "tiny_step_licv.npz"))
    torch.backends.cuda.matmul.allow_tf32 = False
    lmm = HookedModel(tiny_qwen2().cuda(), "llava-interleave-qwen-7b-hf", processor=types.SimpleNamespace(tokenizer=_Tok()))
    enc = AttnFFNShift(lmm, None, "ShiftStrategy.VECTOR_SHIFT | ShiftStrategy.LEARNABLE_SHIFT_SCALE",
                       shift_scale_init_value=0.1).cuda()
    assert set(n for n, p in enc.named_parameters() if not n.startswith("lmm")) == {"ffn_shift", "ffn_shift_scale"}
    assert float(enc.ffn_shift_scale[0]) == pytest.approx(0.1)
    enc.load_state_dict({k[5:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("init.")}, strict=False)
    cfg = types.SimpleNamespace(ce_loss_weight=0.5, align_loss_weight=1.0,
                                data=types.SimpleNamespace(num_image_in_query=1))
    sm = ShiftModel(cfg, enc, Strategy.LM_LOSS | Strategy.LOGITS_KL_DIV)
    full_ids, qa_ids = torch.from_numpy(z["full_ids"]).cuda(), torch.from_numpy(z["qa_ids"]).cuda()
    n0 = ops.launch_count
    loss_dict = sm.forward_inputs({"input_ids": full_ids}, {"input_ids": qa_ids})
    assert set(loss_dict) == {"loss", "ce_loss", "logits_kl_loss"}
    assert abs(float(loss_dict["ce_loss"]) - float(z["ce_loss"])) < 2e-5 * float(z["ce_loss"])
    assert abs(float(loss_dict["logits_kl_loss"]) - float(z["logits_kl_loss"])) < 1e-3 * float(z["logits_kl_loss"])
    assert abs(float(loss_dict["loss"]) - float(z["loss"])) < 2e-5 * float(z["loss"])
    params = {n: p for n, p in enc.named_parameters() if p.requires_grad and not n.startswith("lmm")}
    grads = torch.autograd.grad(loss_dict["loss"], list(params.values()))
    assert ops.launch_count - n0 >= 4 + 4 + 2                 # 4 hook launches each way, CE and KL
    for (n, p), g in zip(params.items(), grads):
        assert rel(g, z["grad." + n]) < 2e-4, n


def test_tiny_step_trainer_update_matches_reference_adamw():
    """f1 end to end: ONE MimicTrainer step (flat gradient buffer written by the kernels -> global-norm clip -> on-device
    AdamW with the reference's two weight-decay groups) lands on the parameters the REFERENCE step produced with
    clip_grad_norm_(1.0) + torch.optim.AdamW(lr 5e-3, wd 1e-3 / 0 on bias) — keys `after.*` of tiny_step.npz
    (src/shift_model.py:306-372, train.yaml grad_clip_val)."""
    from mimic_b200 import Strategy
    from mimic_b200.trainer import MimicTrainer
    z = np.load(os.path.join(GOLDEN, "tiny_step.npz"))
    enc, sm = _tiny_setup(z, Strategy.LM_LOSS | Strategy.LAYER_WISE_MSE)
    # constant lr 5e-3 at the first step: no warm-up, cosine(0) = 1
    tr = MimicTrainer(sm, lr=5e-3, weight_decay=1e-3, warmup_step=0, total_steps=10 ** 9, grad_clip_val=1.0)
    full_ids, qa_ids = torch.from_numpy(z["full_ids"]).cuda(), torch.from_numpy(z["qa_ids"]).cuda()
    ld = tr.micro_step({"input_ids": full_ids}, {"input_ids": qa_ids})
    assert tr.opt_step == 1
    assert abs(float(ld["loss"]) - float(z["loss"])) < 2e-5 * float(z["loss"])
    params = {n: p for n, p in enc.named_parameters() if p.requires_grad and not n.startswith("lmm")}
    for n, p in params.items():
        moved = float(np.abs(z["after." + n] - z["init." + n]).max())
        assert moved > 0, n                                          # the fixture's step really moved this parameter
        # compare the UPDATE (after - init): the parameters themselves are dominated by their initial values
        upd, want = p.detach().cpu().double() - torch.from_numpy(z["init." + n]).double(), \
            torch.from_numpy(z["after." + n]).double() - torch.from_numpy(z["init." + n]).double()
        assert float((upd - want).norm() / want.norm()) < 2e-3, n
        assert float((p.detach().cpu() - torch.from_numpy(z["after." + n])).abs().max()) < 2e-6, n
