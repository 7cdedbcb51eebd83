"""CPU oracle of the reference TEXT ENCODERS -- TEST INFRASTRUCTURE ONLY (imported by tests/ and bench/smoke checkers, never by the
product package).

Functional torch-CPU fp32 restatement of

  CLIPTextModel.forward   /root/reference/models/clip/openclip.py:124-138 (embeddings :53-71, TransformerLayer :86-105,
                          MLP with exact-erf GELU :73-84) -- what models/diffusion.py:194-199 calls as ``clip.text_model``
  TextEncoder.forward     /root/reference/models/clip/clip.py:8-34 (TextEmbedding :36-57, TransformerEncoder :59-94, QuickGELU
                          activation_fn.py:4-9)
  MultiheadSelfAttention  /root/reference/models/clip/attention.py:12-88 with lookahead_mask=True: q/k/v/out Linear WITH bias,
                          causal SDPA, scale = head_dim ** -0.5

Pinned against the unmodified reference: tests/golden/make_golden_clip.py loads ``make_state_dict`` weights into the reference
modules with strict=True and stores their outputs in tests/golden/clip_golden.npz (checked in tests/test_oracle_golden.py).
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import torch
import torch.nn.functional as Fn

OPENCLIP_H = dict(kind="openclip", vocab=49408, hidden=1024, heads=16, layers=23, inter=4096, max_len=77, eps=1e-5)
CLIP_L = dict(kind="clip", vocab=49408, hidden=768, heads=12, layers=12, inter=3072, max_len=77, eps=1e-5)
SMALL_OPENCLIP = dict(kind="openclip", vocab=1000, hidden=256, heads=4, layers=3, inter=1024, max_len=77, eps=1e-5)
SMALL_CLIP = dict(kind="clip", vocab=1000, hidden=768, heads=12, layers=2, inter=3072, max_len=77, eps=1e-5)


def param_spec(kind, vocab, hidden, heads, layers, inter, max_len, **_) -> List[Tuple[str, Tuple[int, ...]]]:
    """(name, shape) in the reference's registration order."""
    c = hidden
    attn = lambda p: [(f"{p}.{n}_proj.{l}", (c, c) if l == "weight" else (c,)) for n in ("q", "k", "v", "out") for l in ("weight", "bias")]
    ln = lambda p: [(f"{p}.weight", (c,)), (f"{p}.bias", (c,))]
    if kind == "openclip":
        s = [("embeddings.token_embedding.weight", (vocab, c)), ("embeddings.position_embedding.weight", (max_len, c))]
        for i in range(layers):
            p = f"encoder.layers.{i}"
            s += ln(f"{p}.layer_norm1") + ln(f"{p}.layer_norm2")
            s += [(f"{p}.mlp.fc1.weight", (inter, c)), (f"{p}.mlp.fc1.bias", (inter,)), (f"{p}.mlp.fc2.weight", (c, inter)), (f"{p}.mlp.fc2.bias", (c,))]
            s += attn(f"{p}.self_attn")
        return s + ln("final_layer_norm")
    s = [("text_embedding.embedding.weight", (vocab, c)), ("text_embedding.position_embedding.weight", (max_len, c))]
    for i in range(layers):
        p = f"encoder_layers.{i}"
        s += attn(f"{p}.self_attn") + ln(f"{p}.layernorm_1")
        s += [(f"{p}.ffn.0.weight", (inter, c)), (f"{p}.ffn.0.bias", (inter,)), (f"{p}.ffn.2.weight", (c, inter)), (f"{p}.ffn.2.bias", (c,))]
        s += ln(f"{p}.layernorm_2")
    return s + ln("final_layer_norm")


def make_state_dict(seed: int, **cfg) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape in param_spec(**cfg):
        if "norm" in name.rsplit(".", 2)[-2]:
            sd[name] = (1.0 if name.endswith("weight") else 0.0) + 0.1 * torch.randn(shape, generator=g)
        elif "embedding" in name:
            sd[name] = torch.randn(shape, generator=g) * (0.02 if "position" not in name else 0.01) * 10   # O(0.1) token features
        elif name.endswith("weight"):
            sd[name] = (torch.rand(shape, generator=g) * 2 - 1) / math.sqrt(shape[1])
        else:
            sd[name] = (torch.rand(shape, generator=g) * 2 - 1) * 0.05
    return sd


def _attention(sd, p, x, heads):
    b, s, c = x.shape
    d = c // heads
    split = lambda t: t.view(b, s, heads, d).permute(0, 2, 1, 3)
    q = split(Fn.linear(x, sd[f"{p}.q_proj.weight"], sd[f"{p}.q_proj.bias"]))
    k = split(Fn.linear(x, sd[f"{p}.k_proj.weight"], sd[f"{p}.k_proj.bias"]))
    v = split(Fn.linear(x, sd[f"{p}.v_proj.weight"], sd[f"{p}.v_proj.bias"]))
    o = Fn.scaled_dot_product_attention(q, k, v, scale=d ** -0.5, is_causal=True).transpose(1, 2).reshape(b, s, c)
    return Fn.linear(o, sd[f"{p}.out_proj.weight"], sd[f"{p}.out_proj.bias"])


def text_forward(sd: Dict[str, torch.Tensor], ids: torch.Tensor, kind, hidden, heads, layers, eps, **_) -> torch.Tensor:
    """ids (B, S) int64 -> (B, S, hidden) fp32."""
    c = hidden
    lnf = lambda z, p: Fn.layer_norm(z, (c,), sd[f"{p}.weight"], sd[f"{p}.bias"], eps)
    s_len = ids.shape[-1]
    if kind == "openclip":
        x = Fn.embedding(ids, sd["embeddings.token_embedding.weight"]) + sd["embeddings.position_embedding.weight"][:s_len][None]
        for i in range(layers):
            p = f"encoder.layers.{i}"
            x = x + _attention(sd, f"{p}.self_attn", lnf(x, f"{p}.layer_norm1"), heads)
            h = Fn.gelu(Fn.linear(lnf(x, f"{p}.layer_norm2"), sd[f"{p}.mlp.fc1.weight"], sd[f"{p}.mlp.fc1.bias"]))
            x = x + Fn.linear(h, sd[f"{p}.mlp.fc2.weight"], sd[f"{p}.mlp.fc2.bias"])
        return lnf(x, "final_layer_norm")
    x = Fn.embedding(ids, sd["text_embedding.embedding.weight"]) + sd["text_embedding.position_embedding.weight"][:s_len][None]
    for i in range(layers):
        p = f"encoder_layers.{i}"
        x = x + _attention(sd, f"{p}.self_attn", lnf(x, f"{p}.layernorm_1"), heads)
        h = Fn.linear(lnf(x, f"{p}.layernorm_2"), sd[f"{p}.ffn.0.weight"], sd[f"{p}.ffn.0.bias"])
        h = h * torch.sigmoid(h * 1.702)
        x = x + Fn.linear(h, sd[f"{p}.ffn.2.weight"], sd[f"{p}.ffn.2.bias"])
    return lnf(x, "final_layer_norm")
