#!/usr/bin/env python
"""bench.py — MimIC train step throughput on B200 (BASELINE.json metric: train tokens/s).

    python bench.py --gpus N --steps K --warmup W            # our arm (one rank per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W   # reference CPU path (rank 0 only)

Workload at N=1 (BASELINE.json configs[1]): Idefics2-8B shapes (Mistral text tower L=32, D=4096,
H=32, KV=8, d=128, inter 14336, V=32003, random-init bf16, frozen), MimIC shifted attention +
LAYER_WISE_MSE + LM loss train step, synthetic VQAv2-like token batch: B=2 per GPU, 3200-token
32-shot context pass (no grad, record H') + 128-token shifted query pass (grad), clip, AdamW.
A "step" is one optimizer step on one such micro-batch.  tokens = B * (S_ctx + T_q) per GPU per step.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    "idefics2-8b": dict(preset="idefics2-8b", B=2, S_ctx=3200, T_q=128),
    # BASELINE config 3: 32-shot Idefics-v1, 3 text tokens per image, pixels enter through the gated cross-attention
    # blocks (stock PyTorch, untouched by MimIC): 33 x 64 image tokens in the context pass, 64 in the query pass
    "idefics-9b": dict(preset="idefics-9b", B=2, S_ctx=1024, T_q=32, n_img_ctx=33, n_img_q=1),
    # BASELINE config 4: MimIC + LoRA (r=1, alpha=2, dropout 0.05 on self_attn.o_proj: scripts/run_mimic_lora.sh)
    "llava-interleave-7b": dict(preset="llava-interleave-7b", B=2, S_ctx=8192, T_q=768,
                                lora=r"language_model\.model\.layers\.\d+\.self_attn\.o_proj$"),
    "tiny": dict(preset="tiny", B=2, S_ctx=512, T_q=64),
}


# dram__bytes_read.sum + dram__bytes_write.sum per launch of the context-pass forward at the Idefics2 shape
TRAFFIC_BYTES = 93.74e6     # 78.79 MB read + 14.95 MB written per launch
TRAFFIC_SOURCE = "profiles/r02_ncu_kernels.md (ncu --set full, attn_fwd_tc64p_kernel at the context-pass shape)"


def workload_name(key: str) -> str:
    return f"{key} shapes, MimIC shifted attention + LAYER_WISE_MSE + LM loss train step, synthetic VQAv2-like batch"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return dict(hbm=j["hbm_gbs"], tf_burst=j["bf16_tflops"], tf_sustained=j["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 9 for n, v in zip(names, r[5:9]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def _load_presets():
    """mimic_b200/presets.py by FILE PATH (torch only): importing the package would map libmimic_b200.so into the
    process, and the reference arm must not load any of the repo's native code."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("_mimic_presets", os.path.join(ROOT, "mimic_b200", "presets.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def cpu_baseline(wl, sample_layers: int, threads: int):
    """The reference's path on host cores (oracle port), bounded sample; see oracle/cpu_step.py."""
    import torch
    from oracle.cpu_step import cpu_mimic_step, cpu_tokens_per_s
    presets = _load_presets()
    torch.set_num_threads(threads)
    kw = presets.PRESETS[wl["preset"]][1]
    full, qa = presets.synthetic_batch(wl["B"], wl["S_ctx"], wl["T_q"], kw["vocab_size"])
    step = cpu_mimic_step(full, qa, D=kw["hidden_size"], inter=kw["intermediate_size"], H=kw["num_attention_heads"],
                          Hkv=kw["num_key_value_heads"], vocab=kw["vocab_size"], n_layers=sample_layers,
                          model_layers=kw["num_hidden_layers"])
    step["wall_s"] = step["layers_s"] + step["head_s"]          # the timed part of the sample (weights are built before it)
    return cpu_tokens_per_s(step), step


def run_reference(args, wl):
    """Reference arm: the reference's own (CPU, PyTorch) implementation of the path, all host threads.
    Rank 0 alone works; the other ranks exit 0.  Each "step" is a BOUNDED SAMPLE: `--cpu-layers` decoder layers of the
    model at the full token shape.  `ms_per_step` is the wall time of that sample (what was really timed);
    `value` extrapolates it to the model's depth (per-layer time x layers + LM head / CE once) and says so."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    assert "mimic_b200" not in sys.modules, "the reference arm must not load the repo's native library"
    threads = os.cpu_count() or 1
    vals, walls, last = [], [], None
    for i in range(args.warmup + args.steps):
        v, last = cpu_baseline(wl, args.cpu_layers, threads)
        if i >= args.warmup:
            vals.append(v)
            walls.append(last["wall_s"])
    value = sum(vals) / len(vals)
    sample = (f"{last['n_layers']} of {last['model_layers']} decoder layers at the full token shape "
              f"(B={wl['B']}, S_ctx={wl['S_ctx']}, T_q={wl['T_q']}), per-layer time x depth + LM head/CE once; "
              f"oracle port of SDPA+do_shift / masked_select+mse_loss / CE, torch CPU fp32")
    line = {"impl": "reference", "metric": "MimIC train tokens/s", "value": value, "unit": "tokens/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * sum(walls) / len(walls), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "extrapolated": True, "steps_timed": len(walls), "layers_timed": last["n_layers"],
            "model_layers": last["model_layers"],
            "ms_per_full_step_extrapolated": 1e3 * last["tokens"] / value,
            "config": {"workload": workload_name(args.workload), "B_per_gpu": wl["B"], "S_ctx": wl["S_ctx"], "T_q": wl["T_q"],
                       "layers": last["model_layers"], "parallelism": f"dp{args.gpus}",
                       "note": "reference CPU path: rank 0 only, the same per-GPU micro-batch; ms_per_step is the wall "
                               "time of the bounded sample, value is extrapolated to the model depth"},
            "cpu_baseline": {"value": value, "unit": "tokens/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def attention_sweep(peaks, *, reps=20, warm=3, with_reference=True):
    """BASELINE config 5 inside the driver-run line: the SHIFTED kernel (active=1, causal, no padding), forward and
    backward, T = S in {1024, 4096, 8192}, B = 16384 / T, H = 32, d = 128, bf16; each kernel timed alone with CUDA
    events (>= 20 reps after 3 warm-ups, q/k/v/out of one call = 0.5 GB > L2), against the BURST bf16 peak.
    Algorithmic FLOPs: fwd 3 B H T^2 d, bwd 8 B H T^2 d.  `gpu_reference_ms`: the reference's own PyTorch
    path (SDPA + do_shift, bf16 autocast; baseline/torch_reference.py) at the same shape on the same GPU."""
    import torch
    from mimic_b200 import ops
    dev, H, d, out = "cuda", 32, 128, []
    for T in (1024, 4096, 8192):
        B = 16384 // T
        g = torch.Generator(device=dev).manual_seed(3407 + T)
        q, k, v = (torch.randn(B, H, T, d, device=dev, generator=g).bfloat16() for _ in range(3))
        do = torch.randn(B, T, H, d, device=dev, generator=g).bfloat16()
        sh = torch.randn(H, d, device=dev, generator=g)            # trained-like parameters: mu is not ~0
        w = torch.randn(H, d, device=dev, generator=g) * 0.1
        b = torch.randn(H, device=dev, generator=g)
        kw = dict(causal=True, active=True, scale=d ** -0.5, impl="tcgen05")
        o, lse_c, lse_f, mu, saved = ops.shifted_attention_forward(q, k, v, None, sh, w, b, **kw)

        def timeit(fn):
            for _ in range(warm):
                fn()
            torch.cuda.synchronize()
            ms = []
            for _ in range(reps):
                a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                fn()
                e.record()
                torch.cuda.synchronize()
                ms.append(a.elapsed_time(e))
            return sum(ms) / len(ms), min(ms)
        f_ms, f_min = timeit(lambda: ops.shifted_attention_forward(q, k, v, None, sh, w, b, **kw))
        b_ms, b_min = timeit(lambda: ops.shifted_attention_backward(saved, o, do, lse_c, lse_f, mu, **kw))
        unit = B * H * T * T * d
        row = {"T": T, "B": B, "H": H, "d": d, "active": 1, "reps": reps,
               "fwd_ms": f_ms, "fwd_ms_min": f_min, "fwd_tflops": 3 * unit / f_ms / 1e9,
               "fwd_frac_of_burst": 3 * unit / f_ms / 1e9 / peaks["tf_burst"],
               "bwd_ms": b_ms, "bwd_ms_min": b_min, "bwd_tflops": 8 * unit / b_ms / 1e9,
               "bwd_frac_of_burst": 8 * unit / b_ms / 1e9 / peaks["tf_burst"]}
        del q, k, v, do, o, saved
        if with_reference:
            from baseline import torch_reference as TR
            row["gpu_reference_fwd_ms"] = TR.attn_fwd_bwd_ms(T, bwd=False, reps=3)
            rb = TR.attn_fwd_bwd_ms(T, bwd=True, reps=3)
            row["gpu_reference_ms"] = rb                              # forward + backward
            if rb:
                row["speedup_fwd_bwd_vs_gpu_reference"] = rb / (f_ms + b_ms)
        out.append(row)
    return out


def query_pass_in_graph(n_calls=200):
    """The query-pass attention of config 2 (B = 2, H = 32, Hkv = 8, T = S = 128, active = 1) as the step runs it: inside
    a CUDA graph.  n_calls calls captured back to back, one pair of events around a replay -> us per call (all launches of
    the op: forward = 1, backward = prep + main + dq conversion).  `own_kernels` above times the same ops eagerly, events
    around each call, launch gaps included."""
    import torch
    from mimic_b200 import ops
    dev, d, B, H, Hkv, T = "cuda", 128, 2, 32, 8, 128
    g = torch.Generator(device=dev).manual_seed(1)
    q = torch.randn(B, H, T, d, device=dev, generator=g).bfloat16()
    k, v = (torch.randn(B, Hkv, T, d, device=dev, generator=g).bfloat16() for _ in range(2))
    do = torch.randn(B, T, H, d, device=dev, generator=g).bfloat16()
    sh, w = torch.randn(H, d, device=dev, generator=g), torch.randn(H, d, device=dev, generator=g) * 0.1
    b = torch.randn(H, device=dev, generator=g)
    kw = dict(causal=True, active=True, scale=d ** -0.5, impl="tcgen05")
    o, lse_c, lse_f, mu, saved = ops.shifted_attention_forward(q, k, v, None, sh, w, b, **kw)
    res = {"shape": "B=2 H=32 Hkv=8 T=S=128 d=128 active=1 bf16", "calls_per_graph": n_calls}
    for name, fn in (("attn_bwd_active_T128_S128_us", lambda: ops.shifted_attention_backward(saved, o, do, lse_c, lse_f, mu, **kw)),
                     ("attn_fwd_active_T128_S128_us", lambda: ops.shifted_attention_forward(q, k, v, None, sh, w, b, **kw))):
        gr, st = torch.cuda.CUDAGraph(), torch.cuda.Stream()
        with torch.cuda.stream(st):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            with torch.cuda.graph(gr, stream=st):
                for _ in range(n_calls):
                    fn()
        torch.cuda.synchronize()
        for _ in range(2):
            gr.replay()
        best = None
        for _ in range(5):
            a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            gr.replay()
            e.record()
            torch.cuda.synchronize()
            us = a.elapsed_time(e) / n_calls * 1e3
            best = us if best is None else min(best, us)
        res[name] = best
        del gr
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="mimic", choices=["mimic", "reference"])
    ap.add_argument("--workload", default="idefics2-8b", choices=list(WORKLOADS))
    ap.add_argument("--layers", type=int, default=None, help="debug: override the number of decoder layers")
    ap.add_argument("--attn-impl", default="auto", choices=["auto", "simt", "tcgen05"])
    ap.add_argument("--cpu-layers", type=int, default=1, help="decoder layers in the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sweep", action="store_true", help="skip the config-5 kernel sweep (N=1 only by default)")
    ap.add_argument("--no-gpu-reference", action="store_true", help="skip the reference-PyTorch-path legs on the GPU")
    ap.add_argument("--no-graph", action="store_true", help="time the eager step instead of the CUDA-graph replay")
    ap.add_argument("--profile-steps", type=int, default=0,
                    help="run this many eager steps and exit (for ncu launch lists; no JSON line)")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        return run_reference(args, wl)
    args.warmup = max(args.warmup, 3)

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly ONE line (the JSON).  NCCL writes its NCCL_DEBUG log to stdout: keep a private handle on
    # the real stdout for the JSON line and point fd 1 at stderr, so the NCCL log stays visible (stderr) and countable.
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if world > 1 and os.environ.get("NCCL_DEBUG", "").upper() not in ("INFO", "TRACE"):
        os.environ["NCCL_DEBUG"] = "INFO"             # communicator / rank lines (on stderr, see above)
        os.environ.setdefault("NCCL_DEBUG_SUBSYS", "INIT")

    import torch
    import torch.distributed as dist
    import mimic_b200
    from mimic_b200 import AttnApproximator, HookedModel, ShiftModel, Strategy, ops
    from mimic_b200.backbone import build_synthetic, synthetic_batch, synthetic_processor
    from mimic_b200.trainer import MimicTrainer
    import types

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world} (launch with torchrun for N>1)"

    model, model_name = build_synthetic(wl["preset"], device=dev, num_layers=args.layers)
    cfgm = model.config
    if wl.get("lora"):
        from mimic_b200.lora import apply_lora
        torch.manual_seed(3407)
        apply_lora(model, wl["lora"], r=1, lora_alpha=2.0, lora_dropout=0.05)
    lmm = HookedModel(model, model_name, processor=synthetic_processor())
    enc = AttnApproximator(lmm, "ShiftStrategy.VECTOR_SHIFT | ShiftStrategy.LEARNABLE_SHIFT_SCALE | ShiftStrategy.MULTI_HEAD",
                           "ShiftStrategy.RECORD_HIDDEN_STATES", attn_impl=args.attn_impl).to(dev)
    cfg = types.SimpleNamespace(ce_loss_weight=0.5, align_loss_weight=1.0, data=types.SimpleNamespace(num_image_in_query=1))
    sm = ShiftModel(cfg, enc, Strategy.LM_LOSS | Strategy.LAYER_WISE_MSE)
    trainer = MimicTrainer(sm, lr=5e-3, weight_decay=1e-3, warmup_step=0.1, total_steps=1000, grad_clip_val=1.0,
                           accumulate_grad_batches=1)
    torch.cuda.set_stream(trainer.stream)     # timing events and copies share the trainer's stream
    B, S, T = wl["B"], wl["S_ctx"], wl["T_q"]
    full_h, qa_h = synthetic_batch(B, S, T, cfgm.vocab_size, seed=3407 + rank)
    full_pin, qa_pin = full_h.pin_memory(), qa_h.pin_memory()
    full_d, qa_d = full_h.to(dev), qa_h.to(dev)
    extra_ctx, extra_q = {}, {}
    if wl.get("n_img_ctx"):                 # synthetic perceiver-resampler outputs (the vision tower is out of scope)
        gi = torch.Generator(device=dev).manual_seed(3407 + rank)
        img = torch.randn(B, wl["n_img_ctx"] * 64, cfgm.vision_embed_dim, device=dev, generator=gi).to(torch.bfloat16) * 0.02
        extra_ctx = {"image_hidden_states": img}
        extra_q = {"image_hidden_states": img[:, -wl["n_img_q"] * 64:].contiguous()}
    pin_ctx = {k: v.cpu().pin_memory() for k, v in extra_ctx.items()}
    pin_q = {k: v.cpu().pin_memory() for k, v in extra_q.items()}

    def finish():
        """Leave through the interpreter's NORMAL exit path (atexit handlers and library destructors run).  A captured
        graph that holds NCCL kernels can stall process-group teardown, so it is dropped first and the group is
        destroyed explicitly while every rank is still alive."""
        import gc
        trainer._graph = None
        trainer._static = None
        ops.graph_keepalive.clear()
        gc.collect()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        json_out.flush()
        sys.stderr.flush()
        if os.environ.get("MIMIC_BENCH_EXIT") == "hard":      # escape hatch: run the exit handlers, then skip destructors
            import atexit
            atexit._run_exitfuncs()
            os._exit(0)

    use_graph = not args.no_graph

    def step_eager():
        trainer.micro_step({"input_ids": full_d, **extra_ctx}, {"input_ids": qa_d, **extra_q})

    def step_resident():
        if use_graph:
            trainer.graph_step([({"input_ids": full_d, **extra_ctx}, {"input_ids": qa_d, **extra_q})])
        else:
            step_eager()

    def step_e2e():
        if use_graph:      # pinned host ids -> static device buffers (H2D inside the timed region) -> replay
            trainer.graph_step([({"input_ids": full_pin, **pin_ctx}, {"input_ids": qa_pin, **pin_q})])
        else:
            f = full_pin.to(dev, non_blocking=True)
            q = qa_pin.to(dev, non_blocking=True)
            trainer.micro_step({"input_ids": f, **{k: v.to(dev, non_blocking=True) for k, v in pin_ctx.items()}},
                               {"input_ids": q, **{k: v.to(dev, non_blocking=True) for k, v in pin_q.items()}})
        return trainer.logged_scalars()          # D2H read of the step's result

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """(max over ranks of the device time of `steps` calls, per-rank times)."""
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            fn()
        b.record()
        barrier()
        ms = torch.tensor([a.elapsed_time(b)], device=dev)
        per_rank = [float(ms)]
        if world > 1:
            allms = [torch.zeros_like(ms) for _ in range(world)]
            dist.all_gather(allms, ms)
            per_rank = [float(x) for x in allms]
        return max(per_rank), per_rank

    if args.profile_steps:
        for _ in range(args.profile_steps):
            step_eager()
        torch.cuda.synchronize()
        print(f"profile: {args.profile_steps} eager steps, {ops.launch_count} launches of our kernels", file=sys.stderr)
        return finish()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()                     # started before warm-up: process start-up stays out of the timed region
    for _ in range(3):
        step_eager()                        # hooks installed, forward replaced, buffers allocated
    n0 = ops.launch_count
    step_eager()
    launches_per_step = ops.launch_count - n0
    if use_graph:
        trainer.capture({"input_ids": full_d, **extra_ctx}, {"input_ids": qa_d, **extra_q})
    for _ in range(args.warmup):
        step_resident()
    if rank == 0:
        sampler.rows.clear()                # keep only samples taken from here on (timed region)
    ms, ms_ranks = timed(step_resident, args.steps)
    launches = launches_per_step * args.steps
    clocks = sampler.stop() if rank == 0 else None
    for _ in range(2):
        step_e2e()
    ms_e2e, ms_e2e_ranks = timed(step_e2e, args.steps)
    scal = trainer.logged_scalars()
    # per-kernel device time of OUR kernels: CUDA events around every C-ABI call, on the launching stream,
    # over a second timed region of eager steps (events cannot bracket nodes inside a graph replay)
    ops.timer.enabled = True
    ops.timer.reset()
    n_eager = max(2, min(args.steps, 4))
    ms_eager, _ = timed(step_eager, n_eager)
    ops.timer.enabled = False
    ksum = ops.timer.summary()
    # who waits for whom: the device time of the all-reduce on each rank (launch to completion, eager steps) — the rank that
    # arrives last waits least; a uniform, small number means the ~1 MB collective itself is what the step pays
    ar_ranks = None
    if world > 1:
        mine = torch.tensor([ksum.get("nccl_allreduce_grads", (0, 0.0, 0.0))[1]], device=dev)
        allar = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allar, mine)
        ar_ranks = [float(x) for x in allar]

    tokens_per_step = B * (S + T) * world
    value = tokens_per_step / (ms / args.steps / 1e3)
    e2e = tokens_per_step / (ms_e2e / args.steps / 1e3)
    if rank != 0:
        return finish()

    peaks = load_peaks()
    H, d, L = cfgm.num_attention_heads, cfgm.head_dim, cfgm.num_hidden_layers
    # dominant own kernel of the step: the context-pass attention forward (inactive handle, causal), 2*B*H*S^2*d FLOPs
    key_ctx = f"attn_fwd_inactive_T{S}_S{S}"
    n_l, mean_ms, tot_ms = ksum.get(key_ctx, (0, float("nan"), 0.0))
    flops = 2.0 * B * H * S * S * d
    achieved = flops / (mean_ms * 1e-3) / 1e12 if n_l else None
    # dram__bytes_read.sum + dram__bytes_write.sum per launch of this kernel at this shape, from the committed
    # `ncu --set full` capture (profiles/, see TRAFFIC_SOURCE); the algorithmic bytes are q + k + v + out = 118 MB
    traffic = TRAFFIC_BYTES if (args.workload == "idefics2-8b" and args.layers is None) else None
    roofline = {"kernel": f"shifted_attn_fwd[{mimic_b200._lib.last_attn_impl()}] context pass (active=0, causal, S={S})",
                "bound": "tensor", "achieved": achieved, "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
                "frac": (achieved / peaks["tf_sustained"]) if achieved else None, "traffic": traffic,
                "traffic_source": TRAFFIC_SOURCE if traffic else None,
                "peak_source": peaks["src"] + " sustained bf16 (kernel timed inside a long step)",
                "frac_of_burst": (achieved / peaks["tf_burst"]) if achieved else None,
                "algorithmic_flops_per_launch": flops, "launches": n_l, "mean_ms": mean_ms,
                "share_of_step": (tot_ms / n_eager) / (ms / args.steps) if n_l else None,
                "timing": f"CUDA events on the launch stream around each of the {n_l} launches in {n_eager} eager steps of this run"}
    own = {k: {"launches": v[0], "mean_ms": round(v[1], 4), "total_ms": round(v[2], 3)} for k, v in sorted(ksum.items())}
    # HBM-bound kernels, algorithmic bytes: MSE 3*N*2B, CE 2*R*V*2B
    extra = {}
    if "mse_fwd_bwd" in ksum:
        n_el = L * B * T * cfgm.hidden_size
        extra["mse_fwd_bwd"] = {"bound": "hbm", "achieved_gbs": 3 * n_el * 2 / (ksum["mse_fwd_bwd"][1] * 1e-3) / 1e9,
                                "peak": peaks["hbm"]}
    if "ce_fwd_bwd" in ksum:
        extra["ce_fwd_bwd"] = {"bound": "hbm", "achieved_gbs": 2 * B * T * cfgm.vocab_size * 2 / (ksum["ce_fwd_bwd"][1] * 1e-3) / 1e9,
                               "peak": peaks["hbm"]}
    line = {
        "metric": "MimIC train tokens/s", "value": value, "unit": "tokens/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": workload_name(args.workload), "B_per_gpu": B, "S_ctx": S, "T_q": T,
                   "layers": L, "attn_impl": mimic_b200._lib.last_attn_impl(), "parallelism": f"dp{world}",
                   "step_launch": "whole-step CUDA graph replay" if use_graph else "eager",
                   "eager_ms_per_step": ms_eager / n_eager,
                   "l2": "inputs larger than L2: every step streams the 14.5 GB frozen backbone"},
        "e2e": {"value": e2e, "unit": "tokens/s", "h2d_bytes_per_step": (full_pin.numel() + qa_pin.numel()) * 8 +
                sum(v.numel() * v.element_size() for v in list(pin_ctx.values()) + list(pin_q.values())),
                "d2h_bytes_per_step": 4 * 4, "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "own_kernels": own, "hbm_kernels": extra,
        "tokens_per_s_per_gpu": value / world, "losses": scal,
        "per_rank_ms_per_step": [m / args.steps for m in ms_ranks],
        "per_rank_e2e_ms_per_step": [m / args.steps for m in ms_e2e_ranks],
        "per_rank_allreduce_ms_eager": ar_ranks,
        "native_library": os.path.relpath(mimic_b200._lib.LIB_PATH, ROOT),
    }
    if world == 1:
        if not args.no_gpu_reference:
            # the same step "the reference way" on this GPU: eager SDPA + do_shift, HF-style rotary, stack/select + mse_loss,
            # F.cross_entropy, torch clip + AdamW (baseline/torch_reference.py) on the same backbone, hooks and batch
            from baseline import torch_reference as TR
            with TR.reference_ops(ops):
                for _ in range(2):
                    step_eager()
                ms_ref, _ = timed(step_eager, 3)
            line["gpu_reference_step"] = {"ms_per_step": ms_ref / 3, "tokens_per_s": tokens_per_step / (ms_ref / 3 / 1e3),
                                          "steps": 3, "what": "reference PyTorch path (SDPA + do_shift + masked_select/mse_loss "
                                          "+ F.cross_entropy + torch AdamW), eager, same backbone / batch / GPU; keeps our "
                                          "row plan for the record path (favourable to the reference)",
                                          "speedup_value_vs_gpu_reference": (ms_ref / 3) / (ms / args.steps)}
        if not args.no_sweep:
            del trainer._graph
            trainer._graph = None
            torch.cuda.empty_cache()
            line["sweep"] = attention_sweep(peaks, with_reference=not args.no_gpu_reference)
            line["query_pass_in_graph"] = query_pass_in_graph()
            best = max(line["sweep"], key=lambda r: r["T"] == 4096)
            line["roofline_shifted"] = {"kernel": "shifted_attn_fwd tcgen05, active=1, causal, T=S=4096, B=4, H=32, d=128 (timed alone)",
                                        "bound": "tensor", "achieved": best["fwd_tflops"], "peak": peaks["tf_burst"],
                                        "unit": "TFLOP/s", "frac": best["fwd_frac_of_burst"],
                                        "peak_source": peaks["src"] + " burst bf16 (kernel timed alone)"}
        if not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            v, stp = cpu_baseline(wl, args.cpu_layers, threads)
            line["cpu_baseline"] = {"value": v, "unit": "tokens/s", "cores": threads, "kind": "port",
                                    "sample": f"{stp['n_layers']} of {stp['model_layers']} decoder layers at the full token "
                                              f"shape ({stp['wall_s']:.1f} s of CPU work), per-layer time x depth + LM head/CE "
                                              f"once (oracle/cpu_step.py)"}
    json_out.write(json.dumps(line) + "\n")
    return finish()


if __name__ == "__main__":
    main()
