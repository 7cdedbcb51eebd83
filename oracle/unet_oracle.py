"""ORACLE (test infrastructure, not product code) — CPU fp32 restatement of the reference UNet
forward as plain functions over a state dict, plus the denoising loop body.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product path never does.

The arithmetic of the reference lives in PyTorch ATen (conv2d, linear, group_norm, layer_norm,
silu, gelu(erf), softmax(QK^T/sqrt(D))V, nearest interpolate — SURVEY.md §8(c) "third-party
arithmetic"), so the restatement uses the same documented torch.nn.functional primitives in
fp32 on CPU and re-derives only the reference's own structure: block order, skip wiring, eps
values, head split, time embedding, CFG order.  Each function cites the reference lines it follows
(paths relative to /root/reference).

Pinning: tests/golden/unet_*.npz were produced by the UNMODIFIED reference
(models/unet/unet.py) with weights from :func:`make_state_dict`, by tests/golden/make_golden.py.
tests/test_oracle_golden.py checks this restatement against them.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence

import torch
import torch.nn.functional as Fn

SD15 = dict(attention_head_dim=[8, 8, 8, 8], cross_attention_dim=768)
SD21 = dict(attention_head_dim=[5, 10, 20, 20], cross_attention_dim=1024)
BLOCK_OUT = [320, 640, 1280, 1280]


# --------------------------------------------------------------------------------------
# parameter inventory (the state-dict contract; reference unet.py:353-401)
# --------------------------------------------------------------------------------------
def _res_spec(p, cin, cout, temb=1280):
    s = [(f"{p}.groupnorm_1.weight", (cin,)), (f"{p}.groupnorm_1.bias", (cin,)),
         (f"{p}.conv_1.weight", (cout, cin, 3, 3)), (f"{p}.conv_1.bias", (cout,)),
         (f"{p}.groupnorm_2.weight", (cout,)), (f"{p}.groupnorm_2.bias", (cout,)),
         (f"{p}.conv_2.weight", (cout, cout, 3, 3)), (f"{p}.conv_2.bias", (cout,)),
         (f"{p}.t_embed.weight", (cout, temb)), (f"{p}.t_embed.bias", (cout,))]
    if cin != cout:
        s += [(f"{p}.proj_input.weight", (cout, cin, 1, 1)), (f"{p}.proj_input.bias", (cout,))]
    return s


def _tr_spec(p, c, dctx):
    t = f"{p}.transformer_block"
    return [(f"{p}.groupnorm.weight", (c,)), (f"{p}.groupnorm.bias", (c,)),
            (f"{p}.conv_input.weight", (c, c, 1, 1)), (f"{p}.conv_input.bias", (c,)),
            (f"{t}.layernorm_1.weight", (c,)), (f"{t}.layernorm_1.bias", (c,)),
            (f"{t}.attn1.q_proj.weight", (c, c)), (f"{t}.attn1.k_proj.weight", (c, c)),
            (f"{t}.attn1.v_proj.weight", (c, c)), (f"{t}.attn1.out_proj.weight", (c, c)),
            (f"{t}.attn1.out_proj.bias", (c,)),
            (f"{t}.layernorm_2.weight", (c,)), (f"{t}.layernorm_2.bias", (c,)),
            (f"{t}.attn2.q_proj.weight", (c, c)), (f"{t}.attn2.k_proj.weight", (c, dctx)),
            (f"{t}.attn2.v_proj.weight", (c, dctx)), (f"{t}.attn2.out_proj.weight", (c, c)),
            (f"{t}.attn2.out_proj.bias", (c,)),
            (f"{t}.layernorm_3.weight", (c,)), (f"{t}.layernorm_3.bias", (c,)),
            (f"{t}.ffn.0.proj.weight", (8 * c, c)), (f"{t}.ffn.0.proj.bias", (8 * c,)),
            (f"{t}.ffn.1.weight", (c, 4 * c)), (f"{t}.ffn.1.bias", (c,)),
            (f"{p}.conv_output.weight", (c, c, 1, 1)), (f"{p}.conv_output.bias", (c,))]


def param_spec(cross_attention_dim=768, in_channels=4, out_channels=4, t_embed_dim=320, **_):
    """(name, shape) for every UNet parameter, in the reference's registration order."""
    ch = BLOCK_OUT
    dctx = cross_attention_dim if isinstance(cross_attention_dim, int) else cross_attention_dim[0]
    temb = 4 * t_embed_dim
    s = [("time_embedding.ffn.0.weight", (temb, t_embed_dim)), ("time_embedding.ffn.0.bias", (temb,)),
         ("time_embedding.ffn.2.weight", (temb, temb)), ("time_embedding.ffn.2.bias", (temb,)),
         ("encoder.conv_in.weight", (ch[0], in_channels, 3, 3)), ("encoder.conv_in.bias", (ch[0],))]
    cin_l = [ch[0]] + ch
    for i in range(4):
        for j in range(2):
            cin = cin_l[i] if j == 0 else ch[i]
            s += _res_spec(f"encoder.down.{i}.block.{j}.0", cin, ch[i], temb)
            if i != 3:
                s += _tr_spec(f"encoder.down.{i}.block.{j}.1", ch[i], dctx)
        if i != 3:
            s += [(f"encoder.down.{i}.downsample.conv.weight", (ch[i], ch[i], 3, 3)),
                  (f"encoder.down.{i}.downsample.conv.bias", (ch[i],))]
    s += _res_spec("bottleneck.0", 1280, 1280, temb) + _tr_spec("bottleneck.1", 1280, dctx) \
        + _res_spec("bottleneck.2", 1280, 1280, temb)
    bin_ = ch + [ch[-1]]
    for j, i in enumerate(reversed(range(4))):
        in_ch, out_ch = bin_[i + 1], ch[i]
        mid_ch = bin_[i - 1] if i > 0 else 320
        for k, cin in enumerate((in_ch + out_ch, out_ch + out_ch, out_ch + mid_ch)):   # unet.py:313-322
            s += _res_spec(f"decoder.up.{j}.block.{k}.0", cin, out_ch, temb)
            if i != 3:
                s += _tr_spec(f"decoder.up.{j}.block.{k}.1", out_ch, dctx)
        if i != 0:
            s += [(f"decoder.up.{j}.upsample.conv.weight", (out_ch, out_ch, 3, 3)),
                  (f"decoder.up.{j}.upsample.conv.bias", (out_ch,))]
    s += [("output.0.weight", (320,)), ("output.0.bias", (320,)),
          ("output.2.weight", (out_channels, 320, 3, 3)), ("output.2.bias", (out_channels,))]
    return s


def make_state_dict(seed: int = 0, **cfg) -> Dict[str, torch.Tensor]:
    """Deterministic synthetic weights (no checkpoint is reachable offline).

    Same scale as PyTorch's default init (U(-1/sqrt(fan_in), 1/sqrt(fan_in))) so activations
    stay O(1) like a freshly constructed reference ``UNet()``; norm gains/offsets are perturbed
    so that they are actually exercised."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape in param_spec(**cfg):
        leaf = name.rsplit(".", 2)[-2]
        is_norm = "norm" in leaf or name.startswith("output.0")
        if is_norm:
            base = 1.0 if name.endswith("weight") else 0.0
            sd[name] = base + 0.1 * torch.randn(shape, generator=g)
        else:
            fan_in = 1
            if name.endswith("weight"):
                for d in shape[1:]:
                    fan_in *= d
                bound = 1.0 / math.sqrt(fan_in)
                sd[name] = (torch.rand(shape, generator=g) * 2 - 1) * bound
            else:
                sd[name] = (torch.rand(shape, generator=g) * 2 - 1) * 0.05
    return sd


# --------------------------------------------------------------------------------------
# forward
# --------------------------------------------------------------------------------------
def time_embedding(sd, timestep, t_embed_dim=320):
    """unet.py:209-220: [cos(t f), sin(t f)], f_i = exp(-ln(1e4) i/half); Linear-SiLU-Linear."""
    half = t_embed_dim // 2
    freqs = torch.exp(-math.log(10000) * torch.arange(0, half, dtype=torch.float32, device=timestep.device) / half)
    x = timestep[:, None].float() * freqs[None, :]
    e = torch.cat([torch.cos(x), torch.sin(x)], dim=-1)
    h = Fn.linear(e, sd["time_embedding.ffn.0.weight"], sd["time_embedding.ffn.0.bias"])
    return Fn.linear(Fn.silu(h), sd["time_embedding.ffn.2.weight"], sd["time_embedding.ffn.2.bias"])


def resblock(sd, p, x, t_embed, eps=1e-5):
    """unet.py:174-195."""
    h = Fn.silu(Fn.group_norm(x, 32, sd[f"{p}.groupnorm_1.weight"], sd[f"{p}.groupnorm_1.bias"], eps))
    h = Fn.conv2d(h, sd[f"{p}.conv_1.weight"], sd[f"{p}.conv_1.bias"], padding=1)
    t = Fn.linear(Fn.silu(t_embed), sd[f"{p}.t_embed.weight"], sd[f"{p}.t_embed.bias"])
    h = h + t[:, :, None, None]
    h = Fn.silu(Fn.group_norm(h, 32, sd[f"{p}.groupnorm_2.weight"], sd[f"{p}.groupnorm_2.bias"], eps))
    h = Fn.conv2d(h, sd[f"{p}.conv_2.weight"], sd[f"{p}.conv_2.bias"], padding=1)
    if f"{p}.proj_input.weight" in sd:
        x = Fn.conv2d(x, sd[f"{p}.proj_input.weight"], sd[f"{p}.proj_input.bias"])
    return h + x


USE_SDPA = False     # bench.py's "stock PyTorch eager on the same GPU" comparator sets this (the reference calls SDPA, attention.py:37-43)


def attention(sd, p, x, cond, heads):
    """unet/attention.py:29-50,70-87: no q/k/v bias, softmax(q k^T / sqrt(D)) v per head, out bias."""
    ctx = x if cond is None else cond
    q = Fn.linear(x, sd[f"{p}.q_proj.weight"])
    k = Fn.linear(ctx, sd[f"{p}.k_proj.weight"])
    v = Fn.linear(ctx, sd[f"{p}.v_proj.weight"])
    b, s, c = q.shape
    d = c // heads

    def split(t):
        return t.view(t.shape[0], t.shape[1], heads, d).permute(0, 2, 1, 3)

    q, k, v = split(q), split(k), split(v)
    if USE_SDPA:
        o = Fn.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(b, s, c)
    else:
        w = torch.softmax((q @ k.transpose(-1, -2)) * (d ** -0.5), dim=-1)
        o = (w @ v).transpose(1, 2).reshape(b, s, c)
    return Fn.linear(o, sd[f"{p}.out_proj.weight"], sd[f"{p}.out_proj.bias"])


def transformer(sd, p, x, cond, heads):
    """unet.py:73-91 (GroupNorm eps hard-coded 1e-6 at :66) and :127-150 (LayerNorm eps 1e-5)."""
    b, c, hh, ww = x.shape
    x_in = x
    x = Fn.group_norm(x, 32, sd[f"{p}.groupnorm.weight"], sd[f"{p}.groupnorm.bias"], 1e-6)
    x = Fn.conv2d(x, sd[f"{p}.conv_input.weight"], sd[f"{p}.conv_input.bias"])
    x = x.view(b, c, -1).transpose(-1, -2)
    t = f"{p}.transformer_block"
    ln = lambda z, i: Fn.layer_norm(z, (c,), sd[f"{t}.layernorm_{i}.weight"], sd[f"{t}.layernorm_{i}.bias"], 1e-5)
    x = x + attention(sd, f"{t}.attn1", ln(x, 1), None, heads)
    x = x + attention(sd, f"{t}.attn2", ln(x, 2), cond, heads)
    h = Fn.linear(ln(x, 3), sd[f"{t}.ffn.0.proj.weight"], sd[f"{t}.ffn.0.proj.bias"])
    a, gate = h.chunk(2, dim=-1)                                   # activation_fn.py:17-20
    h = a * Fn.gelu(gate)                                          # exact erf GELU
    x = x + Fn.linear(h, sd[f"{t}.ffn.1.weight"], sd[f"{t}.ffn.1.bias"])
    x = x.transpose(-1, -2).reshape(b, c, hh, ww)
    x = Fn.conv2d(x, sd[f"{p}.conv_output.weight"], sd[f"{p}.conv_output.bias"])
    return x + x_in


def unet_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, timestep: torch.Tensor, cond: torch.Tensor,
                 attention_head_dim: Sequence[int] = (8, 8, 8, 8), eps: float = 1e-5, **_) -> torch.Tensor:
    """unet.py:431-443 with encoder :284-295, bottleneck :383-391, decoder :337-351, head :398-401.

    NB (unet.py:372-373): ``attention_head_dim`` is used as the NUMBER of heads per level."""
    heads = list(attention_head_dim) if not isinstance(attention_head_dim, int) else [attention_head_dim] * 4
    if cond.shape[0] != x.shape[0]:
        cond = cond.expand(x.shape[0], -1, -1)                     # SDPA batch broadcast
    te = time_embedding(sd, timestep.reshape(-1))
    x = Fn.conv2d(x, sd["encoder.conv_in.weight"], sd["encoder.conv_in.bias"], padding=1)
    skips: List[torch.Tensor] = [x]
    for i in range(4):
        for j in range(2):
            x = resblock(sd, f"encoder.down.{i}.block.{j}.0", x, te, 1e-5 if i != 3 else eps)
            if i != 3:
                x = transformer(sd, f"encoder.down.{i}.block.{j}.1", x, cond, heads[i])
            skips.append(x)
        if i != 3:
            x = Fn.conv2d(x, sd[f"encoder.down.{i}.downsample.conv.weight"],
                          sd[f"encoder.down.{i}.downsample.conv.bias"], stride=2, padding=1)
            skips.append(x)
    x = resblock(sd, "bottleneck.0", x, te)
    x = transformer(sd, "bottleneck.1", x, cond, heads[-1])
    x = resblock(sd, "bottleneck.2", x, te)
    for j, i in enumerate(reversed(range(4))):
        prev_hw = skips[-1].shape[-1]
        for k in range(3):
            x = torch.cat([x, skips.pop()], dim=1)                 # x first (:342-343)
            x = resblock(sd, f"decoder.up.{j}.block.{k}.0", x, te, eps)
            if i != 3:
                x = transformer(sd, f"decoder.up.{j}.block.{k}.1", x, cond, heads[i])
        if i != 0:
            if not (skips and skips[-1].shape[-1] == prev_hw):     # :346-349 (degenerate 1x1 case)
                x = Fn.interpolate(x, scale_factor=2, mode="nearest")
            x = Fn.conv2d(x, sd[f"decoder.up.{j}.upsample.conv.weight"],
                          sd[f"decoder.up.{j}.upsample.conv.bias"], padding=1)
    x = Fn.silu(Fn.group_norm(x, 32, sd["output.0.weight"], sd["output.0.bias"], eps))
    return Fn.conv2d(x, sd["output.2.weight"], sd["output.2.bias"], padding=1)


def synthetic_inputs(batch, h, w, dctx=768, seed=1234, cfg=True):
    """SURVEY.md §8(d) 'synthetic inputs': latent N(0,1), context rows [uncond ; cond]."""
    g = torch.Generator().manual_seed(seed)
    latent = torch.randn((batch, 4, h, w), generator=g)
    ctx = torch.randn(((2 if cfg else 1) * batch, 77, dctx), generator=g)
    return latent, ctx
