"""On-device MimIC step shell: what pl.Trainer + DeepSpeedCPUAdam do around
`ShiftModel.forward` in the reference (src/train.py:94-111, src/shift_model.py:306-390), rebuilt
for one process per GPU:

  * all trainable shift parameters live in ONE flat fp32 buffer (parameters are views), gradients
    in another; the attention backward accumulates its parameter gradients straight into it
    (ops.GradSink), so the data-parallel exchange is a single NCCL all-reduce of ~1 MB per
    optimizer step, with the logged loss scalars riding at its tail (sync_dist, shift_model.py:289);
  * global-norm clip (train.yaml grad_clip_val) + AdamW with the reference's two weight-decay
    groups ("bias" names -> 0, shift_model.py:307-336) run as two small kernels on the device;
  * cosine schedule with linear warm-up (get_cosine_schedule_with_warmup, shift_model.py:384-386),
    gradient accumulation (train.yaml accumulate_grad_batches).
"""
from __future__ import annotations

import math
from contextlib import nullcontext, contextmanager
from typing import Dict, Optional

import torch
import torch.distributed as dist

from . import ops
from .shift_model import ShiftModel


class FlatShiftParams:
    def __init__(self, encoder, weight_decay: float, n_scalars: int = 8, lr: float = 1.0,
                 scale_lr: Optional[float] = None):
        # every trainable parameter reachable from the encoder: the shift parameters and — the backbone being a
        # sub-module (`encoder.lmm`) — any LoRA adapter weights, like the reference's optimizer (src/shift_model.py:338-340)
        named = [(n, p) for n, p in encoder.named_parameters() if p.requires_grad]
        named.sort(key=lambda np_: np_[0].startswith("lmm"))      # stable: shift parameters first, adapters behind them
        if not named:
            raise AssertionError("No parameter to optimize.")
        dev = named[0][1].device
        self.names = [n for n, _ in named]
        self.numel = sum(p.numel() for _, p in named)
        self.param = torch.empty(self.numel, dtype=torch.float32, device=dev)
        # gradient buffer + tail slots for the loss scalars (one all-reduce carries both)
        self.grad_and_scalars = torch.zeros(self.numel + n_scalars, dtype=torch.float32, device=dev)
        self.grad = self.grad_and_scalars[: self.numel]
        self.scalars = self.grad_and_scalars[self.numel:]
        self.weight_decay = torch.empty(self.numel, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros_like(self.param)
        self.exp_avg_sq = torch.zeros_like(self.param)
        # cfg.peft.scale_lr (src/shift_model.py:341-355): names containing "log_Z1" or "scale" train at scale_lr
        self.lr_mul = torch.ones(self.numel, dtype=torch.float32, device=dev) if scale_lr else None
        self.views: Dict[str, torch.Tensor] = {}
        off = 0
        for n, p in named:
            k = p.numel()
            self.param[off:off + k].copy_(p.detach().reshape(-1).float())
            p.data = self.param[off:off + k].view(p.shape)
            g = self.grad[off:off + k].view(p.shape)
            p.grad = g
            self.views[n] = g
            self.weight_decay[off:off + k] = 0.0 if "bias" in n else weight_decay      # shift_model.py:309
            if self.lr_mul is not None and ("log_Z1" in n or "scale" in n):
                self.lr_mul[off:off + k] = scale_lr / lr
            off += k
        # per-layer sinks for the fused attention backward
        if hasattr(encoder, "attn_shift") and hasattr(encoder, "log_Z1_lin") and encoder._fused_ok():
            gs = self.views["attn_shift"]
            encoder.grad_sinks = [
                ops.GradSink(gs[l], self.views[f"log_Z1_lin.{l}.weight"], self.views[f"log_Z1_lin.{l}.bias"])
                for l in range(encoder.lmm_layers)]


def cosine_with_warmup(step: int, warmup: int, total: int) -> float:
    """transformers.get_cosine_schedule_with_warmup lambda (num_cycles 0.5)."""
    if step < warmup:
        return step / max(1, warmup)
    prog = (step - warmup) / max(1, total - warmup)
    return max(0.0, 0.5 * (1.0 + math.cos(math.pi * prog)))


class MimicTrainer:
    def __init__(self, shift_model: ShiftModel, *, lr: float = 5e-3, weight_decay: float = 1e-3, warmup_step=0.1,
                 total_steps: int = 1000, grad_clip_val: float = 1.0, accumulate_grad_batches: int = 1,
                 process_group=None, scale_lr: Optional[float] = None, distributed: bool = True):
        self.sm = shift_model
        self.sm.check_separators = False            # no host sync inside the step
        self.flat = FlatShiftParams(shift_model.shift_encoder, weight_decay, lr=lr, scale_lr=scale_lr)
        self.base_lr, self.total_steps = lr, total_steps
        self.warmup = int(warmup_step * total_steps) if isinstance(warmup_step, float) else int(warmup_step)
        self.clip = grad_clip_val
        self.accum = accumulate_grad_batches
        self.pg = process_group
        # distributed=False: a single-process trainer inside an initialised job (e.g. a reference step on one rank)
        self.world = (dist.get_world_size(process_group)
                      if (distributed and dist.is_available() and dist.is_initialized()) else 1)
        dev = self.flat.param.device
        if self.world > 1:
            # Lightning DDP broadcasts rank 0's module state when it wraps the model (config/train.yaml:14): ranks that
            # were seeded differently, or loaded different checkpoints, must start from identical parameters
            self.broadcast_state()
        self.hyper = torch.zeros(3, dtype=torch.float32, device=dev)       # lr, bias corrections (device: graph replay)
        self.sumsq = ops.adamw_workspace(dev) if dev.type == "cuda" else torch.zeros(1, dtype=torch.float32, device=dev)
        self.opt_step = 0
        self.micro = 0
        self._scalar_keys = ["loss", "ce_loss", "ffn_mse_loss"]     # slots at the tail of the all-reduce buffer (mimic.yaml
                                                                    # default; replaced by what the strategy produces)
        self._graph = None
        self._static = None
        # Every step (eager warm-up, capture, replay) runs on this one non-default stream: autograd binds each
        # leaf's gradient accumulator to the stream of its first use, and a capture only accepts same-stream edges.
        self.stream = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None

    def broadcast_state(self, src: int = 0) -> None:
        """Rank `src`'s parameters and AdamW moments to every rank (call again after loading a checkpoint on one rank)."""
        if self.world > 1:
            for t in (self.flat.param, self.flat.exp_avg, self.flat.exp_avg_sq):
                dist.broadcast(t, src=src, group=self.pg)

    @contextmanager
    def _on_stream(self):
        """Run on the trainer's stream, ordered against the caller's: the trainer's stream first waits for what the caller
        has enqueued (input batches, a loaded checkpoint), and on the way out the caller's stream waits for the trainer's —
        whoever reads `flat.param` / the loss tensors right after `micro_step` sees the finished step.  (Without the second
        wait a 2-rank test that all-gathered the parameters on the default stream read them while AdamW was still running
        on some boxes of the pool.)  Nested use (capture -> _captured_step) and stream capture are left alone."""
        if self.stream is None:
            yield
            return
        outer = torch.cuda.current_stream(self.stream.device)
        if outer == self.stream or torch.cuda.is_current_stream_capturing():
            with torch.cuda.stream(self.stream):
                yield
            return
        self.stream.wait_stream(outer)
        try:
            with torch.cuda.stream(self.stream):
                yield
        finally:
            outer.wait_stream(self.stream)

    def micro_step(self, inputs: Optional[dict], query_inputs: dict) -> Dict[str, torch.Tensor]:
        with self._on_stream():
            return self._micro_step(inputs, query_inputs)

    def _micro_step(self, inputs: Optional[dict], query_inputs: dict) -> Dict[str, torch.Tensor]:
        """One micro-batch: forward (2 passes) + backward.  Steps the optimizer every `accum` calls."""
        if self.micro % self.accum == 0:
            self.flat.grad_and_scalars.zero_()
        loss_dict = self.sm.forward_inputs(inputs, query_inputs, fold_scale=1.0 / self.accum, assume_folded=True)
        (loss_dict["loss"] / self.accum).backward()
        self._accumulate_scalars(loss_dict)
        self.micro += 1
        if self.micro % self.accum == 0:
            self.optimizer_step()
        return loss_dict

    def _accumulate_scalars(self, loss_dict) -> None:
        """Loss terms of the micro-batch into the tail of the all-reduce buffer.  "loss" first, then the per-term keys the
        strategy produces (ce_loss, {attn,ffn}_{mse_loss,cos_sim}, logits_kl_loss): fixed by the strategy, hence the same
        slots on every rank."""
        keys = ["loss"] + sorted(k for k in loss_dict if k != "loss")
        if keys != self._scalar_keys:
            if len(keys) > self.flat.scalars.numel():
                raise RuntimeError(f"{len(keys)} loss terms, {self.flat.scalars.numel()} scalar slots")
            self._scalar_keys = keys
        for i, key in enumerate(keys):
            self.flat.scalars[i] += loss_dict[key].detach().float() / self.accum

    def _advance_schedule(self) -> None:
        """Host side of an optimizer step: step counter, cosine-with-warm-up lr, Adam bias corrections ->
        device `hyper` (12 bytes; enqueued before the kernels / graph that read it)."""
        self.opt_step += 1
        lr = self.base_lr * cosine_with_warmup(self.opt_step - 1, self.warmup, self.total_steps)
        # pageable source: the driver stages the 12 bytes at call time, so the host may run many replays ahead
        # without racing on a shared pinned buffer
        self.hyper.copy_(ops.adamw_hyper(lr, self.opt_step))

    def _device_optimizer_step(self) -> None:
        if self.world > 1:
            _ev = ops.timer.start("nccl_allreduce_grads")                   # (events only when the bench asks for them)
            dist.all_reduce(self.flat.grad_and_scalars, group=self.pg)      # the only collective of the path
            ops.timer.stop(_ev)
        ops.adamw_step(self.flat.param, self.flat.grad, self.flat.exp_avg, self.flat.exp_avg_sq,
                       self.flat.weight_decay, self.hyper, grad_mul=1.0 / self.world,
                       max_norm=self.clip, sumsq=self.sumsq, lr_mul=self.flat.lr_mul)

    def optimizer_step(self) -> None:
        with self._on_stream():
            self._advance_schedule()
            self._device_optimizer_step()

    # ---- whole-step CUDA graph -----------------------------------------------------------------------
    def capture(self, inputs: Optional[dict], query_inputs: dict) -> None:
        """Capture one full optimizer step (accum micro-batches of: context pass, shifted query pass, losses,
        backward; then all-reduce, clip, AdamW) into a CUDA graph over static input buffers.  The eager step is
        launch-bound on a busy host (~4000 small launches); replay removes the host from the loop.  Call after
        a few eager steps (allocations, hook installation and the forward replacement must have happened)."""
        assert self.micro % self.accum == 0, "capture at an optimizer-step boundary"
        with self._on_stream():
            self._capture(inputs, query_inputs)

    def _capture(self, inputs, query_inputs) -> None:
        self._static = ([{k: v.clone() for k, v in (inputs or {}).items()} for _ in range(self.accum)],
                        [{k: v.clone() for k, v in query_inputs.items()} for _ in range(self.accum)])
        # warm-up and capture run real optimizer kernels: neutralise them (lr 0) and restore the state after
        snap = [t.clone() for t in (self.flat.param, self.flat.exp_avg, self.flat.exp_avg_sq, self.hyper)]
        self.hyper.copy_(torch.tensor([0.0, 1.0, 1.0], device=self.hyper.device))
        self._graph_body()                                  # one more warm-up, same stream
        torch.cuda.synchronize()
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph, stream=self.stream):
            self._graph_body()
        torch.cuda.synchronize()
        for dst, src in zip((self.flat.param, self.flat.exp_avg, self.flat.exp_avg_sq, self.hyper), snap):
            dst.copy_(src)

    def _graph_body(self) -> None:
        self.flat.grad_and_scalars.zero_()
        for a in range(self.accum):
            ins, qs = self._static[0][a], self._static[1][a]
            loss_dict = self.sm.forward_inputs(ins if ins else None, qs, fold_scale=1.0 / self.accum, assume_folded=True)
            (loss_dict["loss"] / self.accum).backward()
            self._accumulate_scalars(loss_dict)
        self._device_optimizer_step()

    def graph_step(self, batches) -> None:
        """Replay the captured step on `accum` new micro-batches [(inputs, query_inputs), ...] of the captured
        shapes (host or device tensors; copied into the static buffers)."""
        assert self._graph is not None, "call capture() first"
        with self._on_stream():
            self._graph_step(batches)

    def _graph_step(self, batches) -> None:
        for a, (ins, qs) in enumerate(batches):
            for k, v in (ins or {}).items():
                self._static[0][a][k].copy_(v, non_blocking=True)
            for k, v in qs.items():
                self._static[1][a][k].copy_(v, non_blocking=True)
        self._advance_schedule()
        self._graph.replay()
        self.micro += self.accum

    def logged_scalars(self) -> Dict[str, float]:
        """Mean over ranks of the loss terms of the last optimizer step (one D2H read)."""
        with self._on_stream():
            vals = (self.flat.scalars / self.world).tolist()
        return dict(zip(self._scalar_keys, vals))

    # encoder.pth, same keys as the reference's save_pretrained (src/utils.py:96-102)
    def state_dict_for_checkpoint(self):
        enc = self.sm.shift_encoder
        return {k: v.detach().clone() for k, v in enc.state_dict().items() if not k.startswith("lmm")}
