"""Worker of tests/test_gpu_dp_nccl.py — launched by `python -m torch.distributed.run --nproc-per-node 2`.
Two NCCL ranks, one GPU each, the CUDA kernels everywhere (attention backward writing the flat gradient buffer, the
on-device clip + AdamW): ranks are seeded DIFFERENTLY and see different batches; after each optimizer step every rank
must hold bit-identical parameters (one all-reduce, identical update), and the first step must equal a single-process
step on the MEAN gradient (train.yaml:14 DDP semantics; src/shift_model.py:289 sync_dist for the logged losses)."""
import os
import sys
import types

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
MIMIC = ("ShiftStrategy.VECTOR_SHIFT | ShiftStrategy.LEARNABLE_SHIFT_SCALE | ShiftStrategy.MULTI_HEAD",
         "ShiftStrategy.RECORD_HIDDEN_STATES")


def make(seed, **trainer_kw):
    from mimic_b200 import AttnApproximator, HookedModel, ShiftModel, Strategy
    from mimic_b200.backbone import build_synthetic, synthetic_processor
    from mimic_b200.trainer import MimicTrainer
    model, name = build_synthetic("tiny", dtype=torch.float32, device="cuda")      # same frozen backbone everywhere
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
    return MimicTrainer(sm, lr=5e-3, weight_decay=1e-3, warmup_step=0, total_steps=100, grad_clip_val=1.0, **trainer_kw)


def batch(rank):
    from mimic_b200.backbone import synthetic_batch
    full, qa = synthetic_batch(2, 160, 48, 512, seed=100 + rank, max_left_pad=6)
    return {"input_ids": full.cuda()}, {"input_ids": qa.cuda()}


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    torch.backends.cuda.matmul.allow_tf32 = False
    tr = make(seed=rank)                                   # different initial parameters per rank ...
    gathered = [torch.empty_like(tr.flat.param) for _ in range(world)]
    dist.all_gather(gathered, tr.flat.param)
    assert all(torch.equal(gathered[0], g) for g in gathered), "initial parameters were not broadcast from rank 0"
    init = tr.flat.param.clone()
    ins, qs = batch(rank)                                  # ... and different data
    for step in range(4):                                  # (clipping is active: grad_clip_val = 1.0)
        tr.micro_step(ins, qs)
        dist.all_gather(gathered, tr.flat.param)
        if not all(torch.equal(gathered[0], g) for g in gathered):
            # say what diverged: the all-reduced (scaled) gradients, the squared norm behind the clip factor, or the update
            gg = [torch.empty_like(tr.flat.grad) for _ in range(world)]
            dist.all_gather(gg, tr.flat.grad.contiguous())
            nn = [torch.empty(1, device="cuda") for _ in range(world)]
            dist.all_gather(nn, tr.sumsq[:1].clone())
            dp = float((gathered[0] - gathered[1]).abs().max())
            dg = float((gg[0] - gg[1]).abs().max())
            raise AssertionError(f"ranks diverged after step {step + 1}: max |dparam| {dp:.3e}, max |dgrad| {dg:.3e}, "
                                 f"norm^2 per rank {[float(x) for x in nn]}, differing params "
                                 f"{int((gathered[0] != gathered[1]).sum())} of {gathered[0].numel()}")
        if step == 0:
            after_one = tr.flat.param.clone()
            logged = tr.logged_scalars()
    assert float((after_one - init).abs().max()) > 0
    if rank == 0:
        # single process, both ranks' micro-batches accumulated with 1/2 each == the mean gradient
        ref = make(seed=0, accumulate_grad_batches=world, distributed=False)
        assert ref.world == 1 and torch.equal(ref.flat.param, init)
        for r in range(world):
            ref.micro_step(*batch(r))
        assert ref.opt_step == 1
        upd, want = (after_one - init).double(), (ref.flat.param - init).double()
        err = float((upd - want).norm() / want.norm())
        assert err < 1e-4, f"DP step != mean-gradient step (rel {err:.2e})"
        ref_logged = ref.logged_scalars()
        for k, v in logged.items():
            assert abs(v - ref_logged[k]) < 1e-5 * max(1.0, abs(v)), (k, v, ref_logged[k])
        print(f"DP_NCCL_OK world={world} update_rel_err={err:.2e} loss={logged['loss']:.6f}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
