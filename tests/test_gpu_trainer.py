"""The on-device step shell on the tiny synthetic backbone: flat-buffer gradients written by the
attention backward == autograd's, loss goes down, checkpoint keys are the reference's."""
import types

import pytest
import torch

pytestmark = pytest.mark.gpu
MIMIC = ("ShiftStrategy.VECTOR_SHIFT | ShiftStrategy.LEARNABLE_SHIFT_SCALE | ShiftStrategy.MULTI_HEAD",
         "ShiftStrategy.RECORD_HIDDEN_STATES")


def _setup(dtype, seed=0):
    from mimic_b200 import AttnApproximator, HookedModel, ShiftModel, Strategy
    from mimic_b200.backbone import build_synthetic, synthetic_batch, synthetic_processor
    model, name = build_synthetic("tiny", dtype=dtype, device="cuda")
    lmm = HookedModel(model, name, processor=synthetic_processor())
    torch.manual_seed(seed)
    enc = AttnApproximator(lmm, *MIMIC).cuda()
    with torch.no_grad():
        enc.attn_shift.normal_(0, 0.05)
        for lin in enc.log_Z1_lin:
            lin.weight.normal_(0, 0.1)
            lin.bias.normal_(0, 0.5)
    cfg = types.SimpleNamespace(ce_loss_weight=0.5, align_loss_weight=1.0, data=types.SimpleNamespace(num_image_in_query=1))
    sm = ShiftModel(cfg, enc, Strategy.LM_LOSS | Strategy.LAYER_WISE_MSE)
    full, qa = synthetic_batch(2, 160, 48, 512, max_left_pad=6)
    return sm, enc, full.cuda(), qa.cuda()


def test_flat_grads_equal_autograd():
    from mimic_b200.trainer import FlatShiftParams
    sm, enc, full, qa = _setup(torch.float32)
    ld = sm.forward_inputs({"input_ids": full}, {"input_ids": qa})
    params = {n: p for n, p in enc.named_parameters() if p.requires_grad}
    want = dict(zip(params, torch.autograd.grad(ld["loss"], list(params.values()))))
    flat = FlatShiftParams(enc, 1e-3)
    assert flat.numel == sum(p.numel() for p in params.values()) == 4 * 8 * 32 * 2 + 4 * 8
    flat.grad_and_scalars.zero_()
    ld2 = sm.forward_inputs({"input_ids": full}, {"input_ids": qa})
    ld2["loss"].backward()
    assert abs(float(ld2["loss"]) - float(ld["loss"])) < 1e-6
    for n, p in params.items():
        err = float((p.grad - want[n]).norm() / (want[n].norm() + 1e-30))
        assert err < 1e-5, (n, err)
    assert torch.equal(flat.weight_decay == 0, torch.cat([torch.full((p.numel(),), "bias" in n) for n, p in params.items()]).cuda())


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
def test_training_reduces_loss(dtype):
    from mimic_b200.trainer import MimicTrainer
    sm, enc, full, qa = _setup(dtype)
    tr = MimicTrainer(sm, lr=5e-3, warmup_step=0, total_steps=100, accumulate_grad_batches=2)
    losses = []
    for i in range(12):
        ld = tr.micro_step({"input_ids": full}, {"input_ids": qa})
        losses.append(float(ld["ffn_mse_loss"]))
    assert tr.opt_step == 6
    assert all(l == l for l in losses) and losses[-1] < losses[0], losses
    sc = tr.logged_scalars()
    assert set(sc) == {"loss", "ce_loss", "ffn_mse_loss"} and sc["loss"] > 0
    ck = tr.state_dict_for_checkpoint()
    assert set(ck) == {"attn_shift"} | {f"log_Z1_lin.{l}.{k}" for l in range(4) for k in ("weight", "bias")}


def test_graph_replay_equals_eager():
    """The whole-step CUDA graph (trainer.capture / graph_step) trains exactly like the eager step."""
    from mimic_b200.trainer import MimicTrainer
    finals = []
    for mode in ("eager", "graph"):
        sm, enc, full, qa = _setup(torch.float32, seed=7)
        tr = MimicTrainer(sm, lr=5e-3, warmup_step=2, total_steps=50, accumulate_grad_batches=1)
        batches = [(full.roll(i, 1), qa) for i in range(6)]
        for i in range(2):                                   # hooks installed, forward replaced, buffers allocated
            tr.micro_step({"input_ids": batches[i][0]}, {"input_ids": batches[i][1]})
        if mode == "graph":
            tr.capture({"input_ids": full}, {"input_ids": qa})
        for i in range(2, 6):
            if mode == "graph":
                tr.graph_step([({"input_ids": batches[i][0]}, {"input_ids": batches[i][1]})])
            else:
                tr.micro_step({"input_ids": batches[i][0]}, {"input_ids": batches[i][1]})
        torch.cuda.synchronize()
        assert tr.opt_step == 6
        finals.append((tr.flat.param.clone(), tr.logged_scalars()))
    (pe, se), (pg, sg) = finals
    assert float((pe - pg).norm() / pe.norm()) < 1e-5
    assert abs(se["loss"] - sg["loss"]) < 1e-4 * abs(se["loss"])


def test_licv_trainer_graph_replay_equals_eager():
    """The LIVE baseline (AttnFFNShift with a learnable norm-preserving FFN shift, Strategy.LM_LOSS | LOGITS_KL_DIV,
    second learning rate for the scale parameters) through the same on-device step shell: eager and whole-step CUDA graph
    train identically, the KL term rides in the logged scalars, the loss goes down."""
    from mimic_b200 import AttnFFNShift, HookedModel, ShiftModel, Strategy
    from mimic_b200.backbone import build_synthetic, synthetic_batch, synthetic_processor
    from mimic_b200.trainer import MimicTrainer
    finals = []
    for mode in ("eager", "graph"):
        model, name = build_synthetic("tiny", dtype=torch.float32, device="cuda")
        lmm = HookedModel(model, name, processor=synthetic_processor())
        torch.manual_seed(11)
        enc = AttnFFNShift(lmm, None, "ShiftStrategy.VECTOR_SHIFT | ShiftStrategy.LEARNABLE_SHIFT_SCALE",
                           shift_scale_init_value=0.1).cuda()
        cfg = types.SimpleNamespace(ce_loss_weight=0.5, align_loss_weight=1.0, data=types.SimpleNamespace(num_image_in_query=1))
        sm = ShiftModel(cfg, enc, Strategy.LM_LOSS | Strategy.LOGITS_KL_DIV)
        full, qa = synthetic_batch(2, 160, 48, 512, max_left_pad=6)
        full, qa = full.cuda(), qa.cuda()
        tr = MimicTrainer(sm, lr=5e-3, warmup_step=0, total_steps=50, scale_lr=5e-2)
        assert tr.flat.lr_mul is not None and float(tr.flat.lr_mul.max()) == pytest.approx(10.0)
        first = None
        for i in range(2):
            ld = tr.micro_step({"input_ids": full}, {"input_ids": qa})
            first = first if first is not None else float(ld["logits_kl_loss"])
        if mode == "graph":
            tr.capture({"input_ids": full}, {"input_ids": qa})
        for i in range(2, 8):
            if mode == "graph":
                tr.graph_step([({"input_ids": full}, {"input_ids": qa})])
            else:
                tr.micro_step({"input_ids": full}, {"input_ids": qa})
        torch.cuda.synchronize()
        sc = tr.logged_scalars()
        assert set(sc) == {"loss", "ce_loss", "logits_kl_loss"}
        assert sc["logits_kl_loss"] < first, (sc, first)
        assert set(tr.state_dict_for_checkpoint()) == {"ffn_shift", "ffn_shift_scale"}
        finals.append((tr.flat.param.clone(), sc))
    (pe, se), (pg, sg) = finals
    assert float((pe - pg).norm() / pe.norm()) < 1e-5
    assert abs(se["loss"] - sg["loss"]) < 1e-4 * abs(se["loss"])
