"""Data-parallel path on CPU, world_size 2, gloo: ONE all-reduce over the flat gradient buffer (+ loss
scalars at its tail), 1/world scaling, identical parameters on every rank, equal to the single-process
step on the mean gradient.  The AdamW kernel itself is CUDA-only (tests/test_gpu_losses.py); here it is
replaced by its torch restatement so that the host-side DP logic can run without a GPU."""
import os
import types

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

MIMIC = ("ShiftStrategy.VECTOR_SHIFT | ShiftStrategy.LEARNABLE_SHIFT_SCALE | ShiftStrategy.MULTI_HEAD",
         "ShiftStrategy.RECORD_HIDDEN_STATES")


def _adamw_torch(param, grad, exp_avg, exp_avg_sq, weight_decay, hyper, *, betas=(0.9, 0.999), eps=1e-8, grad_mul=1.0,
                 max_norm=0.0, sumsq=None, lr_mul=None):
    """torch restatement of mimic_adamw_step[_groups] (clip + AdamW), test-only."""
    lr, bc1, bc2s = [float(x) for x in hyper]
    lr = lr * (lr_mul if lr_mul is not None else torch.ones_like(param))
    grad.mul_(grad_mul)
    if max_norm > 0:
        grad.mul_(min(1.0, max_norm / (float(grad.norm()) + 1e-6)))
    param.mul_(1 - lr * weight_decay)
    exp_avg.mul_(betas[0]).add_(grad, alpha=1 - betas[0])
    exp_avg_sq.mul_(betas[1]).addcmul_(grad, grad, value=1 - betas[1])
    param.sub_(lr / bc1 * exp_avg / (exp_avg_sq.sqrt() / bc2s + eps))


def _make_trainer(seed=0):
    """NOTE: replaces ops.adamw_step with the torch restatement for the calling process — callers in the pytest
    process go through the `torch_adamw` fixture, which restores the CUDA front-end afterwards."""
    from tests.test_host_logic import _tiny_lmm
    from mimic_b200 import AttnApproximator, ShiftModel, Strategy, ops
    from mimic_b200.trainer import MimicTrainer
    torch.manual_seed(seed)
    lmm = _tiny_lmm()
    lmm.processor = types.SimpleNamespace(tokenizer=types.SimpleNamespace(pad_token_id=0, bos_token_id=1, padding_side="left"))
    enc = AttnApproximator(lmm, *MIMIC)
    sm = ShiftModel(types.SimpleNamespace(ce_loss_weight=0.5, align_loss_weight=1.0), enc, Strategy.LM_LOSS | Strategy.LAYER_WISE_MSE)
    ops.adamw_step = _adamw_torch
    return MimicTrainer(sm, lr=1e-2, warmup_step=0, total_steps=10, grad_clip_val=1.0)


def _rank_grad(rank, n):
    g = torch.Generator().manual_seed(100 + rank)
    return torch.randn(n, generator=g)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    tr = _make_trainer(seed=rank)          # ranks initialise DIFFERENTLY: the trainer must broadcast rank 0's state
    assert tr.world == world
    n = tr.flat.numel
    tr.flat.grad.copy_(_rank_grad(rank, n))
    tr.flat.scalars[:3] = torch.tensor([1.0 + rank, 2.0 + rank, 3.0 + rank])
    tr.optimizer_step()
    out[rank] = (tr.flat.param.clone(), tr.logged_scalars())
    dist.destroy_process_group()


@pytest.fixture
def torch_adamw():
    from mimic_b200 import ops
    saved = ops.adamw_step
    yield
    ops.adamw_step = saved


def test_two_rank_step_equals_mean_gradient_step(torch_adamw):
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    p0, s0 = out[0]
    p1, s1 = out[1]
    assert torch.equal(p0, p1), "ranks diverged"
    assert s0 == s1 and abs(s0["loss"] - 1.5) < 1e-6 and abs(s0["ffn_mse_loss"] - 3.5) < 1e-6
    # single process (rank 0's initial parameters: what the broadcast hands every rank), mean of the two rank gradients
    tr = _make_trainer(seed=0)
    n = tr.flat.numel
    tr.flat.grad.copy_((_rank_grad(0, n) + _rank_grad(1, n)) / 2)
    tr.optimizer_step()
    assert float((tr.flat.param - p0).abs().max()) < 1e-6
    # parameters are views of the flat buffer; bias entries carry weight-decay 0 (src/shift_model.py:309)
    enc = tr.sm.shift_encoder
    assert enc.attn_shift.data_ptr() == tr.flat.param.data_ptr()
    names = [n_ for n_, p in enc.named_parameters() if p.requires_grad and not n_.startswith("lmm")]
    off = 0
    for n_, p in [(n_, dict(enc.named_parameters())[n_]) for n_ in names]:
        wd = tr.flat.weight_decay[off:off + p.numel()]
        assert bool((wd == (0.0 if "bias" in n_ else 1e-3)).all()), n_
        off += p.numel()


def test_cosine_schedule_matches_transformers():
    from transformers import get_cosine_schedule_with_warmup
    from mimic_b200.trainer import cosine_with_warmup
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.SGD([p], lr=1.0)
    sch = get_cosine_schedule_with_warmup(opt, num_warmup_steps=10, num_training_steps=100)
    for step in range(100):
        assert abs(sch.get_last_lr()[0] - cosine_with_warmup(step, 10, 100)) < 1e-7
        opt.step()
        sch.step()
