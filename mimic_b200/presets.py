"""Workload presets and the synthetic token batch (torch only — no CUDA library needed).

Kept apart from backbone.py so that bench.py's reference arm (`--impl reference`, the CPU path of the reference) can
load this one file by path and share shapes and token batches with the product arm WITHOUT importing the package,
which would map libmimic_b200.so into the reference process."""
from __future__ import annotations

import types

import torch

PRESETS = {
    # name: (model_name for the encoder's substring match, config kwargs)
    "idefics2-8b": ("idefics2-8b-base", dict(hidden_size=4096, intermediate_size=14336, num_hidden_layers=32,
                                             num_attention_heads=32, num_key_value_heads=8, vocab_size=32003)),
    "idefics-9b": ("idefics-9b", dict(hidden_size=4096, intermediate_size=11008, num_hidden_layers=32,
                                      num_attention_heads=32, num_key_value_heads=32, vocab_size=32002,
                                      rms_norm_eps=1e-6, cross_layer_interval=4, vision_embed_dim=1280)),
    "llava-interleave-7b": ("llava-interleave-qwen-7b-hf", dict(hidden_size=4096, intermediate_size=11008,
                                                                num_hidden_layers=32, num_attention_heads=32,
                                                                num_key_value_heads=32, vocab_size=152000,
                                                                rms_norm_eps=1e-6, attention_bias=True,
                                                                rope_theta=1e6)),
    "tiny": ("idefics2-8b-tiny", dict(hidden_size=256, intermediate_size=512, num_hidden_layers=4,
                                      num_attention_heads=8, num_key_value_heads=8, vocab_size=512)),
}


class SyntheticTokenizer:
    pad_token_id, bos_token_id, eos_token_id, padding_side = 0, 1, 2, "left"
    pad_token, eos_token = "<pad>", "</s>"


def synthetic_processor():
    return types.SimpleNamespace(tokenizer=SyntheticTokenizer())


def synthetic_batch(B: int, S_ctx: int, T_q: int, vocab: int, *, seed: int = 3407, max_left_pad: int = 16,
                    n_answer: int = 4, device="cpu"):
    """VQAv2-like token batch: full = [pad..][bos] ICE [PAD] query [PAD] answer [eos] (length S_ctx),
    query-only = [pad..][bos] query [PAD] answer [eos] (length T_q); per-sample 0..max_left_pad left pads
    (exercises the padding path); ids are random in [3, vocab).  Returns (full_ids, query_ids) int64."""
    g = torch.Generator().manual_seed(seed)
    tok = SyntheticTokenizer
    full = torch.full((B, S_ctx), tok.pad_token_id, dtype=torch.int64)
    qa = torch.full((B, T_q), tok.pad_token_id, dtype=torch.int64)
    for i in range(B):
        lp = int(torch.randint(0, max_left_pad + 1, (1,), generator=g)) if max_left_pad > 0 else 0
        if i == 0:
            lp = 0                                            # at least one unpadded row
        n_q = T_q - lp - 1 - 1 - n_answer - 1                 # bos, PAD, answer, eos
        n_ice = S_ctx - lp - 1 - 1 - n_q - 1 - n_answer - 1
        assert n_q > 0 and n_ice > 0
        ice = torch.randint(3, vocab, (n_ice,), generator=g)
        query = torch.randint(3, vocab, (n_q,), generator=g)
        ans = torch.randint(3, vocab, (n_answer,), generator=g)
        pad, bos, eos = (torch.tensor([x]) for x in (tok.pad_token_id, tok.bos_token_id, tok.eos_token_id))
        full[i, lp:] = torch.cat([bos, ice, pad, query, pad, ans, eos])
        qa[i, lp:] = torch.cat([bos, query, pad, ans, eos])
    return full.to(device), qa.to(device)
