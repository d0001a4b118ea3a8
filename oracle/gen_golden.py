"""Generate tests/golden/*.npz by RUNNING THE REFERENCE ITSELF (imported by path, CPU fp32).

Run in the authoring container only (needs /root/reference):

    python -m oracle.gen_golden

The reference has no golden vectors of its own; these fixtures are what pins the
oracle (tests/test_oracle_golden.py) and, through it and directly, the CUDA kernels
(tests/test_gpu_*.py).  TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import os
import types

import numpy as np
import torch
import torch.nn.functional as F

from . import ref_loader

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# name, B, H, Hkv, T, S, d, causal, left_pad (per sample), right_pad, active, multi_head
ATTN_CASES = [
    ("causal_square",      2, 4, 4, 48, 48, 32, True,  None,   None,   True,  True),
    ("causal_square_d128", 1, 2, 2, 160, 160, 128, True, None,  None,   True,  True),
    ("left_pad",           2, 4, 4, 40, 40, 32, True,  [0, 7], None,   True,  True),
    ("right_pad",          2, 4, 4, 40, 40, 32, True,  None,   [3, 0], True,  True),
    ("key_holes",          2, 4, 4, 40, 40, 32, True,  "holes", None,  True,  True),
    ("gqa_group4",         2, 8, 2, 36, 36, 32, True,  None,   None,   True,  True),
    ("kv_cache_T_lt_S",    2, 4, 4, 5,  37, 32, True,  None,   None,   True,  True),
    ("decode_T1",          2, 4, 4, 1,  33, 32, False, None,   None,   True,  True),
    ("inactive",           2, 4, 4, 48, 48, 32, True,  None,   None,   False, True),
    ("single_head_mu",     2, 4, 4, 24, 24, 32, True,  None,   None,   True,  False),
    ("d64_ragged",         1, 3, 3, 77, 77, 64, True,  None,   None,   True,  True),
]


def key_mask_for(case, B, S):
    name, *_rest = case
    left, right = case[8], case[9]
    km = torch.ones(B, S, dtype=torch.bool)
    if isinstance(left, str) and left == "holes":
        # PAD separators in the middle of the row, like ShiftModel's "ICE [PAD] query [PAD] answer"
        km[0, 11] = False
        km[0, 25] = False
        km[1, :4] = False
        km[1, 17] = False
    elif left is not None:
        for b, n in enumerate(left):
            km[b, :n] = False
    if right is not None:
        for b, n in enumerate(right):
            if n:
                km[b, S - n:] = False
    return km


def run_attn_case(case, seed, strategy=None):
    se = ref_loader.load_shift_encoder()
    name, B, H, Hkv, T, S, d, causal, _l, _r, active, multi_head = case
    g = torch.Generator().manual_seed(seed)
    q = torch.randn(B, H, T, d, generator=g)
    k = torch.randn(B, Hkv, S, d, generator=g)
    v = torch.randn(B, Hkv, S, d, generator=g)
    d_out = torch.randn(B, T, H, d, generator=g)
    if strategy is None:
        strategy = ref_loader.MIMIC_ATTN if multi_head else \
            "ShiftStrategy.VECTOR_SHIFT | ShiftStrategy.LEARNABLE_SHIFT_SCALE"
    enc = ref_loader.make_ref_encoder(H * d, 2, H, attn_strategy=strategy)
    has_lin = hasattr(enc, "log_Z1_lin")        # absent without LEARNABLE_SHIFT_SCALE (out = attn + shift, 717-729)
    layer = 1
    with torch.no_grad():                       # "trained-like" values so mu is not ~0
        enc.attn_shift.copy_(torch.randn(enc.attn_shift.shape, generator=g) * 0.5)
        for lin in (enc.log_Z1_lin if has_lin else []):
            lin.weight.copy_(torch.randn(lin.weight.shape, generator=g) * 0.2)
            lin.bias.copy_(torch.randn(lin.bias.shape, generator=g))
    enc.attn_shift_handles[layer].active = active
    km = key_mask_for(case, B, S)

    q_, k_, v_ = (t.clone().requires_grad_(True) for t in (q, k, v))
    G = H // Hkv
    kk = k_[:, :, None].expand(B, Hkv, G, S, d).reshape(B, H, S, d)      # HF repeat_kv
    vv = v_[:, :, None].expand(B, Hkv, G, S, d).reshape(B, H, S, d)
    plain_causal = causal and bool(km.all()) and T == S
    if plain_causal:                            # reference: mask None -> is_causal (348-350 / 447)
        attn = F.scaled_dot_product_attention(q_, kk, vv, attn_mask=None, is_causal=T > 1)
        row_valid = torch.ones(B, T, dtype=torch.bool)
    else:                                       # reference: explicit 4-D mask from HF, is_causal False
        s_idx = torch.arange(S)[None, None, None, :]
        t_idx = torch.arange(T)[None, None, :, None]
        mask = km[:, None, None, :].expand(B, 1, T, S).clone()
        if causal:
            mask &= s_idx <= t_idx + (S - T)
        row_valid = mask.any(-1)[:, 0]          # rows with no visible key: reference undefined
        add = torch.zeros(B, 1, T, S).masked_fill(~mask, torch.finfo(torch.float32).min)
        attn = F.scaled_dot_product_attention(q_, kk, vv, attn_mask=add, is_causal=False)
    attn_t = attn.transpose(1, 2)
    out = enc.do_shift(layer, q_, kk, attn_t)   # THE reference call
    out = out.reshape(B, T, H, d)
    d_out = d_out * row_valid[:, :, None, None]  # fully-masked rows never reach a loss
    params = [enc.attn_shift] + ([p for p in enc.log_Z1_lin[layer].parameters()] if has_lin else [])
    grads = torch.autograd.grad(out, [q_, k_, v_] + params, d_out, allow_unused=True)
    dq, dk, dv, dshift_all = grads[:4]
    rec = dict(
        q=q, k=k, v=v, d_out=d_out, key_mask=km, row_valid=row_valid,
        shift=enc.attn_shift[layer].detach().reshape(-1), out=out.detach(), dq=dq, dk=dk, dv=dv,
        meta=np.array([B, H, Hkv, T, S, d, int(causal), int(active), int(multi_head)], dtype=np.int64),
    )
    if has_lin:
        rec.update(w=enc.log_Z1_lin[layer].weight.detach(), b=enc.log_Z1_lin[layer].bias.detach())
    if active:
        rec.update(dshift=dshift_all[layer].reshape(-1))
        if has_lin:
            rec.update(dw=grads[4], db=grads[5])
    return {k_: (v_.numpy() if isinstance(v_, torch.Tensor) else v_) for k_, v_ in rec.items()}


def gen_attention():
    for i, case in enumerate(ATTN_CASES):
        rec = run_attn_case(case, 3407 + i)
        np.savez_compressed(os.path.join(OUT, f"attn_{case[0]}.npz"), **rec)
        print("attn", case[0], "out-norm", float(np.linalg.norm(rec["out"])))


# ---------------------------------------------------------------- other encoder variants
_V, _L, _M = "ShiftStrategy.VECTOR_SHIFT", "ShiftStrategy.LEARNABLE_SHIFT_SCALE", "ShiftStrategy.MULTI_HEAD"
VARIANT_CASES = [
    # attention-side variants of AttnApproximator.do_shift: key prefix, case tuple (ATTN_CASES layout), strategy string
    ("sh_pad",    ("", 2, 4, 2, 40, 40, 32, True, [0, 7], None, True, False), f"{_V} | {_L}"),   # head-averaged mu, pads, GQA
    ("sh_cache",  ("", 2, 4, 4, 6, 37, 32, True, None, None, True, False), f"{_V} | {_L}"),      # ... with T < S
    ("vec_multi", ("", 2, 4, 4, 24, 24, 32, True, None, None, True, True), f"{_V} | {_M}"),      # out = attn + shift[h]
    ("vec_flat",  ("", 2, 4, 4, 24, 24, 32, True, None, None, True, False), _V),                 # out = attn + shift
]


def gen_variants():
    """tests/golden/variants.npz: the AttnApproximator strategies other than mimic.yaml's, the norm-preserving shift
    hook of AttnFFNShift (src/shift_encoder.py:222-247) and of AttnApproximator's FFN hook (653-672), all produced by
    calling the reference's own methods."""
    se = ref_loader.load_shift_encoder()
    rec = {}
    for i, (key, case, strategy) in enumerate(VARIANT_CASES):
        r = run_attn_case(case, 4407 + i, strategy=strategy)
        rec.update({f"{key}.{k_}": v_ for k_, v_ in r.items()})
        print("variant", key, "out-norm", float(np.linalg.norm(r["out"])))

    g = torch.Generator().manual_seed(5407)
    D, L = 256, 3
    enc = se.AttnFFNShift(ref_loader.make_fake_lmm(D, L, 8), attn_strategy=f"{_V} | {_L}", ffn_strategy=_V,
                          shift_scale_init_value=0.1)
    with torch.no_grad():                                   # trained-like: the shift is visible next to h ~ N(0,1)
        enc.attn_shift.copy_(torch.randn(L, D, generator=g) * 0.8)
        enc.ffn_shift.copy_(torch.randn(L, D, generator=g) * 0.3)
        enc.attn_shift_scale.copy_(torch.tensor([0.1, 0.7, -0.4]))
    for prefix, layer in (("attn", 1), ("ffn", 2)):
        h = torch.randn(2, 9, D, generator=g).requires_grad_(True)
        dy = torch.randn(2, 9, D, generator=g)
        hook = enc._shift_hook(prefix)
        y, rest = hook(None, None, (h, "rest"), module_name=f"model.text_model.layers.{layer}.{prefix}")
        assert rest == "rest"
        shift, scale = getattr(enc, f"{prefix}_shift"), getattr(enc, f"{prefix}_shift_scale")
        leaves = [h, shift] + ([scale] if scale.requires_grad else [])
        grads = torch.autograd.grad(y, leaves, dy)
        rec.update({f"ns_{prefix}.h": h.detach(), f"ns_{prefix}.dy": dy, f"ns_{prefix}.shift": shift[layer].detach(),
                    f"ns_{prefix}.scale": scale[layer].detach(), f"ns_{prefix}.y": y.detach(),
                    f"ns_{prefix}.dh": grads[0], f"ns_{prefix}.dshift": grads[1][layer]})
        if scale.requires_grad:
            rec[f"ns_{prefix}.dscale"] = grads[2][layer]
    # AttnApproximator's FFN hook: same arithmetic, scale fixed at 1 (src/shift_encoder.py:661-667)
    enc2 = ref_loader.make_ref_encoder(D, L, 8, ffn_strategy=f"{_V} | ShiftStrategy.RECORD_HIDDEN_STATES")
    with torch.no_grad():
        enc2.ffn_shift.copy_(torch.randn(L, D, generator=g) * 0.5)
    captured = {}
    enc2.lmm.register_forward_hook = lambda target, fn: captured.setdefault("hook", fn)
    enc2.lmm.replace_module_method = lambda *a, **k: None
    enc2.register_shift_hooks()
    h = torch.randn(3, 5, D, generator=g).requires_grad_(True)
    dy = torch.randn(3, 5, D, generator=g)
    y = captured["hook"](None, None, h, module_name="model.text_model.layers.0.mlp")
    grads = torch.autograd.grad(y, [h, enc2.ffn_shift], dy)
    rec.update({"ns_approx.h": h.detach(), "ns_approx.dy": dy, "ns_approx.shift": enc2.ffn_shift[0].detach(),
                "ns_approx.y": y.detach(), "ns_approx.dh": grads[0], "ns_approx.dshift": grads[1][0]})
    rec = {k_: (v_.numpy() if isinstance(v_, torch.Tensor) else v_) for k_, v_ in rec.items()}
    np.savez_compressed(os.path.join(OUT, "variants.npz"), **rec)


# ---------------------------------------------------------------- shift_model pieces

class _Tok:
    pad_token_id, bos_token_id, eos_token_id, padding_side = 0, 1, 2, "left"
    pad_token, eos_token = "<pad>", "</s>"


def make_ref_shift_model(enc, strategy_str="Strategy.LM_LOSS | Strategy.LAYER_WISE_MSE"):
    sm = ref_loader.load_shift_model()
    enc.lmm.processor = types.SimpleNamespace(tokenizer=_Tok())
    cfg = types.SimpleNamespace(ce_loss_weight=0.5, align_loss_weight=1.0,
                                data=types.SimpleNamespace(num_image_in_query=1))
    return sm.ShiftModel(cfg, enc, eval(strategy_str, {"Strategy": sm.Strategy})), sm


def gen_loss_pieces():
    enc = ref_loader.make_ref_encoder(64, 3, 4)
    model, sm = make_ref_shift_model(enc)
    rec = {}
    # 1. the survey's hand example and two more rows
    ids = torch.tensor([[0, 0, 1, 5, 6, 0, 7, 8, 0, 9, 2, 3],
                        [1, 4, 4, 4, 0, 6, 6, 6, 0, 7, 7, 2]])
    rec["ids"] = ids.numpy()
    for n in (1, 2):
        rec[f"label_mask_sep{n}"] = model.generate_label_mask({"input_ids": ids}, n).numpy()
    rec["label_mask_sep1_keep_bos"] = model.generate_label_mask({"input_ids": ids}, 1, keep_bos=True).numpy()
    # 2. get_hidden_states + LAYER_WISE_MSE on ragged Nq_i
    g = torch.Generator().manual_seed(3407)
    L, B, S, D = 3, 2, 12, 64
    ctx = [torch.randn(B, S, D, generator=g) for _ in range(L)]
    mask_ctx = model.generate_label_mask({"input_ids": ids}, 1)          # 5 and 6 rows
    enc.ffn_hidden_states = ctx
    hp = model.get_hidden_states(mask_ctx)
    T = 8
    qry = [torch.randn(B, T, D, generator=g) for _ in range(L)]
    mask_q = torch.zeros(B, T, dtype=torch.bool)
    mask_q[0, T - int(mask_ctx[0].sum()):] = True
    mask_q[1, T - int(mask_ctx[1].sum()):] = True
    enc.ffn_hidden_states = [t.clone().requires_grad_(True) for t in qry]
    hq = model.get_hidden_states(mask_q)
    loss = model.calculate_layer_wise_loss(hq, hp)
    assert list(loss.keys()) == ["ffn_mse_loss"]
    grads = torch.autograd.grad(loss["ffn_mse_loss"], enc.ffn_hidden_states, retain_graph=True)
    rec.update(ctx=torch.stack(ctx).numpy(), qry=torch.stack(qry).numpy(), mask_ctx=mask_ctx.numpy(),
               mask_q=mask_q.numpy(), ffn_mse_loss=loss["ffn_mse_loss"].detach().numpy(),
               dqry=torch.stack(grads).numpy(),
               hp0=hp["ffn_hidden_states"][0].numpy(), hp1=hp["ffn_hidden_states"][1].numpy())
    # 3. COS_SIM variant
    model2, _ = make_ref_shift_model(enc, "Strategy.LAYER_WISE_COS_SIM")
    cos = model2.calculate_layer_wise_loss(hq, hp)["ffn_cos_sim"]
    rec["ffn_cos_sim"] = cos.detach().numpy()
    rec["dqry_cos"] = torch.stack(torch.autograd.grad(cos, enc.ffn_hidden_states)).numpy()
    # 4. logits KL
    a = torch.randn(7, 50, generator=g)
    b = torch.randn(7, 50, generator=g)
    m = torch.ones(7, dtype=torch.bool)
    a.requires_grad_(True)
    kl = model.calculate_logits_kl_loss(a, b, m, m)["logits_kl_loss"]
    rec.update(kl_a=a.detach().numpy(), kl_b=b.numpy(), kl=kl.detach().numpy(),
               dkl_a=torch.autograd.grad(kl, a)[0].numpy())
    # 4b. ragged selection: 5 of 9 query rows paired with 5 of 11 context rows (boolean-mask order), odd vocabulary
    a2 = torch.randn(9, 37, generator=g).requires_grad_(True)
    b2 = torch.randn(11, 37, generator=g)
    ma = torch.tensor([0, 1, 1, 0, 1, 0, 1, 1, 0], dtype=torch.bool)
    mb = torch.tensor([1, 0, 0, 1, 1, 0, 0, 1, 0, 0, 1], dtype=torch.bool)
    kl2 = model.calculate_logits_kl_loss(a2, b2, ma, mb)["logits_kl_loss"]
    rec.update(kl2_a=a2.detach().numpy(), kl2_b=b2.numpy(), kl2_ma=ma.numpy(), kl2_mb=mb.numpy(), kl2=kl2.detach().numpy(),
               dkl2_a=torch.autograd.grad(kl2, a2)[0].numpy())
    np.savez_compressed(os.path.join(OUT, "loss_pieces.npz"), **rec)
    print("loss pieces: mse", float(rec["ffn_mse_loss"]), "cos", float(rec["ffn_cos_sim"]), "kl", float(rec["kl"]))


# ---------------------------------------------------------------- tiny end-to-end step (BASELINE config 1)

def build_tiny_ids(seed=3407, B=2, n_ctx=96, n_query=(20, 14), n_answer=(3, 5), vocab=512):
    """[bos] ICE [PAD] query [PAD] answer [EOS], left-padded; and the query-only rows."""
    g = torch.Generator().manual_seed(seed)
    tok = _Tok()
    full, qa = [], []
    for i in range(B):
        ice = torch.randint(3, vocab, (n_ctx - 3 * i,), generator=g).tolist()
        query = torch.randint(3, vocab, (n_query[i],), generator=g).tolist()
        ans = torch.randint(3, vocab, (n_answer[i],), generator=g).tolist()
        full.append([tok.bos_token_id] + ice + [tok.pad_token_id] + query + [tok.pad_token_id] + ans + [tok.eos_token_id])
        qa.append([tok.bos_token_id] + query + [tok.pad_token_id] + ans + [tok.eos_token_id])

    def left_pad(rows):
        n = max(len(r) for r in rows)
        return torch.tensor([[tok.pad_token_id] * (n - len(r)) + r for r in rows])

    return left_pad(full), left_pad(qa)


def tiny_qwen2(seed=3407):
    from transformers import Qwen2Config, Qwen2ForCausalLM
    torch.manual_seed(seed)
    cfg = Qwen2Config(vocab_size=512, hidden_size=256, intermediate_size=512, num_hidden_layers=4,
                      num_attention_heads=8, num_key_value_heads=8, attn_implementation="sdpa",
                      max_position_embeddings=1024)
    cfg.text_config = cfg
    cfg.use_sliding_window = False
    lm = Qwen2ForCausalLM(cfg).float()
    lm.requires_grad_(False)

    class Wrap(torch.nn.Module):                       # names match language_model\.model\.layers\.\d+\.self_attn$
        def __init__(self):
            super().__init__()
            self.language_model = lm
            self.config = cfg

        def forward(self, **kw):
            return self.language_model(**kw)

    return Wrap()


def gen_tiny_step():
    mb = ref_loader.load_model_base()
    from torch import nn

    class TinyLMM(mb.ModelBase):
        def __init__(self, model):
            nn.Module.__init__(self)
            self.model = model
            self.model_name = "llava-interleave-qwen-7b-hf"
            self.config = model.config
            self._trackers_dict = {}

    lmm = TinyLMM(tiny_qwen2())
    enc = ref_loader.make_ref_encoder(256, 4, 8, lmm=lmm)
    g = torch.Generator().manual_seed(1234)
    with torch.no_grad():                              # trained-like so the shift path matters
        enc.attn_shift.copy_(torch.randn(enc.attn_shift.shape, generator=g) * 0.05)
        for lin in enc.log_Z1_lin:
            lin.weight.copy_(torch.randn(lin.weight.shape, generator=g) * 0.1)
            lin.bias.copy_(torch.randn(lin.bias.shape, generator=g) * 0.5)
    init = {k: v.detach().clone() for k, v in enc.state_dict().items() if not k.startswith("lmm")}
    model, sm = make_ref_shift_model(enc)
    full_ids, qa_ids = build_tiny_ids()
    tok = _Tok()

    # ---- ShiftModel.forward (src/shift_model.py:193-285) restated at token-id level; every
    # ---- mask / hidden-state / loss call below is the imported reference method.
    hooks = enc.register_record_hooks()
    inputs = {"input_ids": full_ids, "attention_mask": full_ids != tok.pad_token_id}
    q_inputs = {"input_ids": qa_ids, "attention_mask": qa_ids != tok.pad_token_id}
    with torch.no_grad():
        prefix_logits = lmm.model(**inputs)["logits"]
    prefix_hs = model.get_hidden_states(model.generate_label_mask(inputs, 1))
    hooks.update(enc.register_shift_hooks())
    logits = lmm.model(**q_inputs)["logits"]
    # LM loss, transformers==4.45.2 rule (oracle/losses.py::lm_ce_loss; parity unpinned there)
    keep = q_inputs["attention_mask"][:, 1:]
    ce = F.cross_entropy(logits[:, :-1][keep].float(), qa_ids[:, 1:][keep])
    shift_hs = model.get_hidden_states(q_inputs["attention_mask"] & (qa_ids != tok.bos_token_id))
    model.remove_hooks(hooks)
    layer_loss = model.calculate_layer_wise_loss(shift_hs, prefix_hs)
    loss = 0.5 * ce + 1.0 * sum(layer_loss.values())
    params = {n: p for n, p in enc.named_parameters() if p.requires_grad and not n.startswith("lmm")}
    # same step under Strategy.LM_LOSS | LAYER_WISE_COS_SIM | LOGITS_KL_DIV (src/shift_model.py:152-160, 181-191, 272-283)
    model_ck, _ = make_ref_shift_model(enc, "Strategy.LM_LOSS | Strategy.LAYER_WISE_COS_SIM | Strategy.LOGITS_KL_DIV")
    cos_loss = model_ck.calculate_layer_wise_loss(shift_hs, prefix_hs)
    kl_loss = model_ck.calculate_logits_kl_loss(logits, prefix_logits, model_ck.generate_label_mask(q_inputs, 1),
                                                model_ck.generate_label_mask(inputs, 2))
    loss_ck = 0.5 * ce + 1.0 * sum(cos_loss.values()) + 1.0 * sum(kl_loss.values())
    grads_ck = torch.autograd.grad(loss_ck, list(params.values()), retain_graph=True)
    grads = torch.autograd.grad(loss, list(params.values()))
    rec = {"cosk.loss": loss_ck.detach().numpy(), "cosk.ffn_cos_sim": cos_loss["ffn_cos_sim"].detach().numpy(),
           "cosk.logits_kl_loss": kl_loss["logits_kl_loss"].detach().numpy()}
    for (n, p), gr in zip(params.items(), grads_ck):
        rec["cosk.grad." + n] = gr.numpy()
    rec.update({"full_ids": full_ids.numpy(), "qa_ids": qa_ids.numpy(), "loss": loss.detach().numpy(),
           "ce_loss": ce.detach().numpy(), "ffn_mse_loss": layer_loss["ffn_mse_loss"].detach().numpy(),
           "logits_checksum": logits.detach().double().abs().sum().numpy(),
           "weights_checksum": np.array(sum(float(p.double().abs().sum()) for p in lmm.model.parameters()))})
    for (n, p), gr in zip(params.items(), grads):
        rec["init." + n] = init[n].numpy()
        rec["grad." + n] = gr.numpy()
    # one AdamW step with the reference's two weight-decay groups (src/shift_model.py:306-372)
    for (n, p), gr in zip(params.items(), grads):
        p.grad = gr
    torch.nn.utils.clip_grad_norm_(list(params.values()), 1.0)            # train.yaml grad_clip_val
    opt = torch.optim.AdamW([
        {"params": [p for n, p in params.items() if "bias" in n], "weight_decay": 0.0, "lr": 5e-3},
        {"params": [p for n, p in params.items() if "bias" not in n], "weight_decay": 1e-3, "lr": 5e-3}],
        weight_decay=1e-3)
    opt.step()
    for n, p in params.items():
        rec["after." + n] = p.detach().numpy()
    np.savez_compressed(os.path.join(OUT, "tiny_step.npz"), **rec)
    print("tiny step: loss", float(loss), "ce", float(ce), "mse", float(layer_loss["ffn_mse_loss"]))


class _OracleLora(torch.nn.Module):
    """LoRA on a frozen Linear as peft applies it (y = W x + (alpha / r) B A x; dropout 0 here), written out for the
    oracle so the fixture does not depend on the product's mimic_b200.lora."""

    def __init__(self, base, r, alpha, gen):
        super().__init__()
        self.base, self.scaling, self.enabled = base, alpha / r, True
        self.A = torch.nn.Parameter(torch.randn(r, base.in_features, generator=gen) * 0.1)
        self.B = torch.nn.Parameter(torch.randn(base.out_features, r, generator=gen) * 0.1)

    def forward(self, x):
        y = self.base(x)
        return y + (x @ self.A.t() @ self.B.t()) * self.scaling if self.enabled else y


def gen_tiny_step_lora():
    """BASELINE config 4 in miniature: MimIC + LoRA (r = 1, alpha = 2) on every self_attn.o_proj, the context pass with the
    adapters disabled (src/shift_model.py:225), through the reference's own hooks and loss methods."""
    mb = ref_loader.load_model_base()
    from torch import nn

    class TinyLMM(mb.ModelBase):
        def __init__(self, model):
            nn.Module.__init__(self)
            self.model = model
            self.model_name = "llava-interleave-qwen-7b-hf"
            self.config = model.config
            self._trackers_dict = {}

    wrap = tiny_qwen2()
    g = torch.Generator().manual_seed(4321)
    loras = {}
    for li, layer in enumerate(wrap.language_model.model.layers):
        lo = _OracleLora(layer.self_attn.o_proj, 1, 2.0, g)
        layer.self_attn.o_proj = lo
        loras[f"language_model.model.layers.{li}.self_attn.o_proj"] = lo
    lmm = TinyLMM(wrap)
    enc = ref_loader.make_ref_encoder(256, 4, 8, lmm=lmm)
    g2 = torch.Generator().manual_seed(1234)
    with torch.no_grad():
        enc.attn_shift.copy_(torch.randn(enc.attn_shift.shape, generator=g2) * 0.05)
        for lin in enc.log_Z1_lin:
            lin.weight.copy_(torch.randn(lin.weight.shape, generator=g2) * 0.1)
            lin.bias.copy_(torch.randn(lin.bias.shape, generator=g2) * 0.5)
    init = {k: v.detach().clone() for k, v in enc.state_dict().items() if not k.startswith("lmm")}
    model, sm = make_ref_shift_model(enc)
    full_ids, qa_ids = build_tiny_ids()
    tok = _Tok()
    hooks = enc.register_record_hooks()
    inputs = {"input_ids": full_ids, "attention_mask": full_ids != tok.pad_token_id}
    q_inputs = {"input_ids": qa_ids, "attention_mask": qa_ids != tok.pad_token_id}
    for lo in loras.values():
        lo.enabled = False                                    # with self.lmm.model.disable_adapter()
    with torch.no_grad():
        lmm.model(**inputs)
    for lo in loras.values():
        lo.enabled = True
    prefix_hs = model.get_hidden_states(model.generate_label_mask(inputs, 1))
    hooks.update(enc.register_shift_hooks())
    logits = lmm.model(**q_inputs)["logits"]
    keep = q_inputs["attention_mask"][:, 1:]
    ce = F.cross_entropy(logits[:, :-1][keep].float(), qa_ids[:, 1:][keep])
    shift_hs = model.get_hidden_states(q_inputs["attention_mask"] & (qa_ids != tok.bos_token_id))
    model.remove_hooks(hooks)
    layer_loss = model.calculate_layer_wise_loss(shift_hs, prefix_hs)
    loss = 0.5 * ce + 1.0 * sum(layer_loss.values())
    params = {n: p for n, p in enc.named_parameters() if p.requires_grad and not n.startswith("lmm")}
    lora_params = {f"{n}.{ab}": getattr(lo, ab) for n, lo in loras.items() for ab in ("A", "B")}
    grads = torch.autograd.grad(loss, list(params.values()) + list(lora_params.values()))
    rec = {"full_ids": full_ids.numpy(), "qa_ids": qa_ids.numpy(), "loss": loss.detach().numpy(),
           "ce_loss": ce.detach().numpy(), "ffn_mse_loss": layer_loss["ffn_mse_loss"].detach().numpy()}
    for (n, p), gr in zip(list(params.items()) + list(lora_params.items()), grads):
        rec["grad." + n] = gr.numpy()
    for n, v in init.items():
        rec["init." + n] = v.numpy()
    for n, p in lora_params.items():
        rec["lora." + n] = p.detach().numpy()
    np.savez_compressed(os.path.join(OUT, "tiny_step_lora.npz"), **rec)
    print("tiny step + LoRA: loss", float(loss.detach()), "ce", float(ce.detach()), "mse", float(layer_loss["ffn_mse_loss"].detach()))


def gen_tiny_step_licv():
    """The LIVE baseline's step in miniature (config/encoder/licv.yaml: AttnFFNShift with a learnable, norm-preserving
    FFN shift, scale init 0.1; Strategy.LM_LOSS | LOGITS_KL_DIV) through the reference's own hooks and loss methods."""
    mb = ref_loader.load_model_base()
    se = ref_loader.load_shift_encoder()
    from torch import nn

    class TinyLMM(mb.ModelBase):
        def __init__(self, model):
            nn.Module.__init__(self)
            self.model = model
            self.model_name = "llava-interleave-qwen-7b-hf"
            self.config = model.config
            self._trackers_dict = {}

    lmm = TinyLMM(tiny_qwen2())
    enc = se.AttnFFNShift(lmm, attn_strategy=None, ffn_strategy=f"{_V} | {_L}", shift_scale_init_value=0.1)
    g = torch.Generator().manual_seed(4321)
    with torch.no_grad():                              # trained-like: shift comparable to the MLP output norm
        enc.ffn_shift.copy_(torch.randn(enc.ffn_shift.shape, generator=g) * 0.05)
        enc.ffn_shift_scale.copy_(torch.tensor([0.1, 0.6, 1.3, -0.5]))
    init = {k: v.detach().clone() for k, v in enc.state_dict().items() if not k.startswith("lmm")}
    model, sm = make_ref_shift_model(enc, "Strategy.LM_LOSS | Strategy.LOGITS_KL_DIV")
    full_ids, qa_ids = build_tiny_ids()
    tok = _Tok()
    hooks = enc.register_record_hooks()                # none: nothing is recorded under licv.yaml
    assert not hooks
    inputs = {"input_ids": full_ids, "attention_mask": full_ids != tok.pad_token_id}
    q_inputs = {"input_ids": qa_ids, "attention_mask": qa_ids != tok.pad_token_id}
    with torch.no_grad():
        prefix_logits = lmm.model(**inputs)["logits"]
    hooks.update(enc.register_shift_hooks())
    assert set(hooks) == {"ffn_hook"} and len(hooks["ffn_hook"]) == 4
    logits = lmm.model(**q_inputs)["logits"]
    keep = q_inputs["attention_mask"][:, 1:]
    ce = F.cross_entropy(logits[:, :-1][keep].float(), qa_ids[:, 1:][keep])
    kl = model.calculate_logits_kl_loss(logits, prefix_logits, model.generate_label_mask(q_inputs, 1),
                                        model.generate_label_mask(inputs, 2))
    model.remove_hooks(hooks)
    loss = 0.5 * ce + 1.0 * sum(kl.values())
    params = {n: p for n, p in enc.named_parameters() if p.requires_grad and not n.startswith("lmm")}
    assert set(params) == {"ffn_shift", "ffn_shift_scale"}
    grads = torch.autograd.grad(loss, list(params.values()))
    rec = {"full_ids": full_ids.numpy(), "qa_ids": qa_ids.numpy(), "loss": loss.detach().numpy(),
           "ce_loss": ce.detach().numpy(), "logits_kl_loss": kl["logits_kl_loss"].detach().numpy()}
    for (n, p), gr in zip(params.items(), grads):
        rec["init." + n] = init[n].numpy()
        rec["grad." + n] = gr.numpy()
    np.savez_compressed(os.path.join(OUT, "tiny_step_licv.npz"), **rec)
    print("tiny step (licv): loss", float(loss), "ce", float(ce), "kl", float(kl["logits_kl_loss"]))


def main():
    assert ref_loader.available(), "reference tree not found; golden vectors can only be made where it is"
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    gen_attention()
    gen_loss_pieces()
    gen_tiny_step_lora()
    gen_tiny_step()
    gen_variants()
    gen_tiny_step_licv()


if __name__ == "__main__":
    main()
