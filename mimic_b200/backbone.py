"""Synthetic frozen backbones with the module layout of the three reference targets, for
measurement and tests (there is no network for checkpoints): random-init weights of the named
architecture, module names matched by the reference's regexes (src/shift_encoder.py:95-111).

  idefics2-8b  : model.text_model.layers.N.{self_attn,mlp}      Mistral-7B text tower (GQA 32:8)
  idefics-9b   : model.layers.N.{self_attn,mlp}                 LLaMA-7B  (MHA)
  llava-interleave-qwen-7b : language_model.model.layers.N.*    Qwen2-7B  (MHA here, q/k/v bias)

The backbone GEMMs are stock PyTorch (cuBLAS): out of scope for hand-written kernels (SURVEY.md
§8d).  The stock `Attention.forward` below is plain SDPA — it is what runs until
`AttnApproximator.register_shift_hooks()` replaces it, exactly like an HF attention module.
"""
from __future__ import annotations

import types
from dataclasses import dataclass
from typing import Optional

import torch
import torch.nn.functional as F
from torch import nn

from .presets import PRESETS, SyntheticTokenizer, synthetic_batch, synthetic_processor  # noqa: F401


@dataclass
class DecoderConfig:
    hidden_size: int = 4096
    intermediate_size: int = 14336
    num_hidden_layers: int = 32
    num_attention_heads: int = 32
    num_key_value_heads: int = 8
    vocab_size: int = 32003
    rms_norm_eps: float = 1e-5
    rope_theta: float = 10000.0
    attention_bias: bool = False
    attention_dropout: float = 0.0
    head_dim: Optional[int] = None
    initializer_range: float = 0.02
    _attn_implementation: str = "sdpa"
    use_sliding_window: bool = False
    sliding_window: Optional[int] = None
    # Idefics v1 only: a gated cross-attention block in front of every `cross_layer_interval`-th decoder layer, keys /
    # values from `image_hidden_states` [B, n_img * 64, vision_embed_dim] (HF IdeficsGatedCrossAttentionLayer)
    cross_layer_interval: int = 0
    vision_embed_dim: int = 1280

    def __post_init__(self):
        if self.head_dim is None:
            self.head_dim = self.hidden_size // self.num_attention_heads
        self.text_config = self


class RMSNorm(nn.Module):
    def __init__(self, dim, eps):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(dim))
        self.eps = eps

    def forward(self, x):
        return F.rms_norm(x, (x.shape[-1],), self.weight, self.eps)


def rope_cos_sin(position_ids: torch.Tensor, head_dim: int, theta: float, dtype):
    inv = 1.0 / (theta ** (torch.arange(0, head_dim, 2, device=position_ids.device, dtype=torch.float32) / head_dim))
    ang = position_ids[..., None].float() * inv                       # [B,T,d/2]
    emb = torch.cat((ang, ang), dim=-1)
    return emb.cos().to(dtype), emb.sin().to(dtype)


def _rotate_half(x):
    x1, x2 = x[..., : x.shape[-1] // 2], x[..., x.shape[-1] // 2:]
    return torch.cat((-x2, x1), dim=-1)


class Attention(nn.Module):
    def __init__(self, cfg: DecoderConfig, layer_idx: int):
        super().__init__()
        self.config, self.layer_idx = cfg, layer_idx
        self.head_dim = cfg.head_dim
        self.num_key_value_groups = cfg.num_attention_heads // cfg.num_key_value_heads
        self.scaling = self.head_dim ** -0.5
        self.attention_dropout = cfg.attention_dropout
        self.is_causal = True
        self.sliding_window = None
        D, H, KV, d = cfg.hidden_size, cfg.num_attention_heads, cfg.num_key_value_heads, cfg.head_dim
        self.q_proj = nn.Linear(D, H * d, bias=cfg.attention_bias)
        self.k_proj = nn.Linear(D, KV * d, bias=cfg.attention_bias)
        self.v_proj = nn.Linear(D, KV * d, bias=cfg.attention_bias)
        self.o_proj = nn.Linear(H * d, D, bias=False)

    def forward(self, hidden_states, position_embeddings=None, attention_mask=None, past_key_values=None, **kwargs):
        """Stock path (what the reference's un-replaced HF module does): SDPA with a causal+key mask."""
        B, T, _ = hidden_states.shape
        d = self.head_dim
        q = self.q_proj(hidden_states).view(B, T, -1, d).transpose(1, 2)
        k = self.k_proj(hidden_states).view(B, T, -1, d).transpose(1, 2)
        v = self.v_proj(hidden_states).view(B, T, -1, d).transpose(1, 2)
        cos, sin = position_embeddings
        cos, sin = cos.unsqueeze(1), sin.unsqueeze(1)
        q, k = q * cos + _rotate_half(q) * sin, k * cos + _rotate_half(k) * sin
        g = self.num_key_value_groups
        if g > 1:
            k = k[:, :, None].expand(B, k.shape[1], g, T, d).reshape(B, -1, T, d)
            v = v[:, :, None].expand(B, v.shape[1], g, T, d).reshape(B, -1, T, d)
        if attention_mask is None:
            out = F.scaled_dot_product_attention(q, k, v, is_causal=T > 1)
        else:
            m = attention_mask.bool()[:, None, None, :] & torch.ones(T, T, dtype=torch.bool, device=q.device).tril()[None, None]
            out = F.scaled_dot_product_attention(q, k, v, attn_mask=m)
        return self.o_proj(out.transpose(1, 2).reshape(B, T, -1)), None


class MLP(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.gate_proj = nn.Linear(cfg.hidden_size, cfg.intermediate_size, bias=False)
        self.up_proj = nn.Linear(cfg.hidden_size, cfg.intermediate_size, bias=False)
        self.down_proj = nn.Linear(cfg.intermediate_size, cfg.hidden_size, bias=False)

    def forward(self, x):
        return self.down_proj(F.silu(self.gate_proj(x)) * self.up_proj(x))


class DecoderLayer(nn.Module):
    def __init__(self, cfg, layer_idx):
        super().__init__()
        self.self_attn = Attention(cfg, layer_idx)
        self.mlp = MLP(cfg)
        self.input_layernorm = RMSNorm(cfg.hidden_size, cfg.rms_norm_eps)
        self.post_attention_layernorm = RMSNorm(cfg.hidden_size, cfg.rms_norm_eps)

    def forward(self, x, position_embeddings, attention_mask):
        a, *_ = self.self_attn(hidden_states=self.input_layernorm(x), position_embeddings=position_embeddings,
                               attention_mask=attention_mask)
        x = x + a
        return x + self.mlp(self.post_attention_layernorm(x))


class CrossAttention(nn.Module):
    """Idefics v1 cross-attention (same IdeficsAttention class as self-attention in HF, `is_cross_attention=True`):
    queries from the text stream, keys / values from the image features, no RoPE, no causal mask.  It lives at
    `model.gated_cross_attn_layers.N.cross_attn`, which the reference's self-attention regex
    (`model\\.layers\\.\\d+\\.self_attn$`, src/shift_encoder.py:107) does not match: MimIC leaves it untouched."""

    def __init__(self, cfg: DecoderConfig):
        super().__init__()
        self.is_cross_attention = True
        self.head_dim = cfg.head_dim
        D, H, d = cfg.hidden_size, cfg.num_attention_heads, cfg.head_dim
        self.q_proj = nn.Linear(D, H * d, bias=False)
        self.k_proj = nn.Linear(cfg.vision_embed_dim, H * d, bias=False)
        self.v_proj = nn.Linear(cfg.vision_embed_dim, H * d, bias=False)
        self.o_proj = nn.Linear(H * d, D, bias=False)

    def forward(self, hidden_states, key_value_states=None, attention_mask=None, **kwargs):
        B, T, _ = hidden_states.shape
        d = self.head_dim
        q = self.q_proj(hidden_states).view(B, T, -1, d).transpose(1, 2)
        k = self.k_proj(key_value_states).view(B, key_value_states.shape[1], -1, d).transpose(1, 2)
        v = self.v_proj(key_value_states).view(B, key_value_states.shape[1], -1, d).transpose(1, 2)
        out = F.scaled_dot_product_attention(q, k, v, attn_mask=attention_mask)
        return self.o_proj(out.transpose(1, 2).reshape(B, T, -1)), None


class GatedCrossAttentionLayer(nn.Module):
    """x + tanh(alpha_cross_attn) * cross_attn(LN(x), image) ; then + tanh(alpha_dense) * mlp(LN(x))."""

    def __init__(self, cfg: DecoderConfig):
        super().__init__()
        self.cross_attn = CrossAttention(cfg)
        self.mlp = MLP(cfg)
        self.input_layernorm = RMSNorm(cfg.hidden_size, cfg.rms_norm_eps)
        self.post_attention_layernorm = RMSNorm(cfg.hidden_size, cfg.rms_norm_eps)
        self.alpha_cross_attn = nn.Parameter(torch.zeros(1))
        self.alpha_dense = nn.Parameter(torch.zeros(1))

    def forward(self, x, image_hidden_states, image_attention_mask=None):
        a, _ = self.cross_attn(self.input_layernorm(x), key_value_states=image_hidden_states,
                               attention_mask=image_attention_mask)
        x = x + torch.tanh(self.alpha_cross_attn).to(x.dtype) * a
        return x + torch.tanh(self.alpha_dense).to(x.dtype) * self.mlp(self.post_attention_layernorm(x))


class TextModel(nn.Module):
    def __init__(self, cfg: DecoderConfig):
        super().__init__()
        self.config = cfg
        self.embed_tokens = nn.Embedding(cfg.vocab_size, cfg.hidden_size)
        self.layers = nn.ModuleList(DecoderLayer(cfg, i) for i in range(cfg.num_hidden_layers))
        if cfg.cross_layer_interval > 0:
            n_cross = (cfg.num_hidden_layers + cfg.cross_layer_interval - 1) // cfg.cross_layer_interval
            self.gated_cross_attn_layers = nn.ModuleList(GatedCrossAttentionLayer(cfg) for _ in range(n_cross))
        self.norm = RMSNorm(cfg.hidden_size, cfg.rms_norm_eps)

    def forward(self, input_ids=None, inputs_embeds=None, attention_mask=None, position_ids=None,
                image_hidden_states=None, image_attention_mask=None):
        x = self.embed_tokens(input_ids) if inputs_embeds is None else inputs_embeds
        B, T, _ = x.shape
        if position_ids is None:
            position_ids = torch.arange(T, device=x.device)[None].expand(B, T)
        pe = rope_cos_sin(position_ids, self.config.head_dim, self.config.rope_theta, x.dtype)
        interval = self.config.cross_layer_interval
        for i, layer in enumerate(self.layers):
            if interval > 0 and i % interval == 0 and image_hidden_states is not None:
                x = self.gated_cross_attn_layers[i // interval](x, image_hidden_states, image_attention_mask)
            x = layer(x, pe, attention_mask)
        return self.norm(x)


class _Holder(nn.Module):
    pass


class SyntheticLM(nn.Module):
    """Top-level module whose `named_modules()` match the reference regexes of `family`."""

    def __init__(self, family: str, cfg: DecoderConfig):
        super().__init__()
        self.family, self.config = family, cfg
        text = TextModel(cfg)
        if family == "idefics2-8b":
            self.model = _Holder()
            self.model.text_model = text                     # model.text_model.layers.N.*
        elif family == "idefics-9b":
            self.model = text                                # model.layers.N.*
        elif family.startswith("llava-interleave"):
            self.language_model = _Holder()
            self.language_model.model = text                 # language_model.model.layers.N.*
        else:
            raise ValueError(family)
        self._text = [text]                                  # not a registered child under a second name
        self.lm_head = nn.Linear(cfg.hidden_size, cfg.vocab_size, bias=False)

    @property
    def text(self) -> TextModel:
        return self._text[0]

    def forward(self, input_ids=None, inputs_embeds=None, attention_mask=None, position_ids=None,
                logits_to_keep: int = 0, image_hidden_states=None, image_attention_mask=None, **kwargs):
        h = self.text(input_ids=input_ids, inputs_embeds=inputs_embeds, attention_mask=attention_mask,
                      position_ids=position_ids, image_hidden_states=image_hidden_states,
                      image_attention_mask=image_attention_mask)
        hk = h[:, -logits_to_keep:] if logits_to_keep else h          # HF semantics: 0 = all positions
        return {"logits": self.lm_head(hk), "last_hidden_state": h}


def build_synthetic(preset: str, *, dtype=torch.bfloat16, device="cuda", seed: int = 3407, num_layers: Optional[int] = None,
                    **overrides):
    """Random-init (N(0, 0.02), seed 3407 like the reference's pl.seed_everything, train.py:85) frozen
    backbone + the `model_name` the encoder keys on.  Returns (SyntheticLM, model_name)."""
    model_name, kw = PRESETS[preset]
    kw = dict(kw, **overrides)
    if num_layers is not None:
        kw["num_hidden_layers"] = num_layers
    cfg = DecoderConfig(**kw)
    family = "idefics2-8b" if preset == "tiny" else preset
    with torch.device(device):
        model = SyntheticLM(family if not family.startswith("llava") else "llava-interleave", cfg).to(dtype)
    gen = torch.Generator(device=device).manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if p.dim() >= 2:
                p.normal_(0.0, cfg.initializer_range, generator=gen)
            elif "norm" in name:
                p.fill_(1.0)
            elif "alpha_" in name:
                p.fill_(0.5)                                  # trained-like gate (HF initialises it at 0 = block switched off)
            else:
                p.zero_()
    model.requires_grad_(False)
    return model, model_name
