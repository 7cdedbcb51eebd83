"""CPU oracle of the reference VAE DECODER (``VAE.decode``) -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import this module; the product package never does.

Functional torch-CPU fp32 restatement of /root/reference/models/vae/vae.py:

  VAE.decode                 vae.py:270-274   z / 0.18215 -> post_quant_conv (1x1) -> decoder
  VAE_Decoder.forward        vae.py:229-241   conv_in, mid (res, attention, res), 4 up blocks (3 res + nearest-2x upsample conv), GN, SiLU, conv_out
  ResidualBlock.forward      resnet.py:27-41  GN(32, eps 1e-6) -> SiLU -> conv3x3 -> GN -> SiLU -> conv3x3 (+ 1x1 shortcut when cin != cout)
  AttentionBlock.forward     vae.py:121-134   GN -> q/k/v Linear WITH bias -> single-head SDPA (head_dim = C = 512) -> proj_attn -> + x
  UpSample.forward           vae.py:37-40     nn.Upsample(scale_factor=2) (nearest) -> conv3x3

The arithmetic lives in PyTorch (ATen) exactly as for the UNet oracle.  Pinned against the UNMODIFIED reference executed in the
build container: tests/golden/make_golden_vae.py loads ``make_state_dict`` weights into the reference ``VAE`` with strict=True and
stores its ``decode`` outputs in tests/golden/vae_golden.npz; tests/test_oracle_golden.py checks this restatement against them.
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import torch
import torch.nn.functional as Fn

CH, CH_MULT, SCALE = 128, (1, 2, 4, 4), 0.18215


def _res(p, cin, cout):
    s = [(f"{p}.norm1.weight", (cin,)), (f"{p}.norm1.bias", (cin,)), (f"{p}.conv1.weight", (cout, cin, 3, 3)), (f"{p}.conv1.bias", (cout,)),
         (f"{p}.norm2.weight", (cout,)), (f"{p}.norm2.bias", (cout,)), (f"{p}.conv2.weight", (cout, cout, 3, 3)), (f"{p}.conv2.bias", (cout,))]
    if cin != cout:
        s += [(f"{p}.conv_shortcut.weight", (cout, cin, 1, 1)), (f"{p}.conv_shortcut.bias", (cout,))]
    return s


def _attn(p, c):
    s = [(f"{p}.group_norm.weight", (c,)), (f"{p}.group_norm.bias", (c,))]
    for n in ("query", "key", "value", "proj_attn"):
        s += [(f"{p}.{n}.weight", (c, c)), (f"{p}.{n}.bias", (c,))]
    return s


def decoder_blocks() -> List[Tuple[int, int, int, bool]]:
    """(up block index, cin of its first ResidualBlock, cout, has upsampler) in forward order (vae.py:208-224)."""
    out, block_in = [], CH * CH_MULT[-1]
    for j, i in enumerate(reversed(range(len(CH_MULT)))):
        block_out = CH * CH_MULT[i]
        out.append((j, block_in, block_out, i != 0))
        block_in = block_out
    return out


def param_spec(in_channels=3, z_channels=4, out_channels=3) -> List[Tuple[str, Tuple[int, ...]]]:
    """(name, shape) of every parameter of the reference ``VAE`` (encoder included: ``load_state_dict(strict=True)`` contract)."""
    s = [("encoder.conv_in.weight", (CH, in_channels, 3, 3)), ("encoder.conv_in.bias", (CH,))]
    cur = CH
    for i, m in enumerate(CH_MULT):
        out = CH * m
        for j in range(2):
            s += _res(f"encoder.down_blocks.{i}.resnets.{j}", cur if j == 0 else out, out)
        if i != len(CH_MULT) - 1:
            s += [(f"encoder.down_blocks.{i}.downsamplers.0.conv.weight", (out, out, 3, 3)), (f"encoder.down_blocks.{i}.downsamplers.0.conv.bias", (out,))]
        cur = out
    s += _res("encoder.mid_block.resnets.0", cur, cur) + _res("encoder.mid_block.resnets.1", cur, cur) + _attn("encoder.mid_block.attentions.0", cur)
    s += [("encoder.conv_norm_out.weight", (cur,)), ("encoder.conv_norm_out.bias", (cur,)),
          ("encoder.conv_out.weight", (2 * z_channels, cur, 3, 3)), ("encoder.conv_out.bias", (2 * z_channels,))]
    top = CH * CH_MULT[-1]
    s += [("decoder.conv_in.weight", (top, z_channels, 3, 3)), ("decoder.conv_in.bias", (top,))]
    s += _attn("decoder.mid_block.attentions.0", top) + _res("decoder.mid_block.resnets.0", top, top) + _res("decoder.mid_block.resnets.1", top, top)
    for j, cin, cout, up in decoder_blocks():
        for k in range(3):
            s += _res(f"decoder.up_blocks.{j}.resnets.{k}", cin if k == 0 else cout, cout)
        if up:
            s += [(f"decoder.up_blocks.{j}.upsamplers.0.conv.weight", (cout, cout, 3, 3)), (f"decoder.up_blocks.{j}.upsamplers.0.conv.bias", (cout,))]
    s += [("decoder.conv_norm_out.weight", (CH,)), ("decoder.conv_norm_out.bias", (CH,)),
          ("decoder.conv_out.weight", (out_channels, CH, 3, 3)), ("decoder.conv_out.bias", (out_channels,))]
    s += [("quant_conv.weight", (2 * z_channels, 2 * z_channels, 1, 1)), ("quant_conv.bias", (2 * z_channels,)),
          ("post_quant_conv.weight", (z_channels, z_channels, 1, 1)), ("post_quant_conv.bias", (z_channels,))]
    return s


def make_state_dict(seed: int = 0) -> Dict[str, torch.Tensor]:
    """Deterministic synthetic weights at PyTorch's default-init scale; norm gains / offsets perturbed so that they are exercised."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape in param_spec():
        if "norm" in name.rsplit(".", 2)[-2]:
            sd[name] = (1.0 if name.endswith("weight") else 0.0) + 0.1 * torch.randn(shape, generator=g)
        elif name.endswith("weight"):
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            sd[name] = (torch.rand(shape, generator=g) * 2 - 1) / math.sqrt(fan_in)
        else:
            sd[name] = (torch.rand(shape, generator=g) * 2 - 1) * 0.05
    return sd


def resblock(sd, p, x):
    """resnet.py:27-41 (dropout 0)."""
    h = Fn.silu(Fn.group_norm(x, 32, sd[f"{p}.norm1.weight"], sd[f"{p}.norm1.bias"], 1e-6))
    h = Fn.conv2d(h, sd[f"{p}.conv1.weight"], sd[f"{p}.conv1.bias"], padding=1)
    h = Fn.silu(Fn.group_norm(h, 32, sd[f"{p}.norm2.weight"], sd[f"{p}.norm2.bias"], 1e-6))
    h = Fn.conv2d(h, sd[f"{p}.conv2.weight"], sd[f"{p}.conv2.bias"], padding=1)
    if f"{p}.conv_shortcut.weight" in sd:
        x = Fn.conv2d(x, sd[f"{p}.conv_shortcut.weight"], sd[f"{p}.conv_shortcut.bias"])
    return h + x


def attention(sd, p, x):
    """vae.py:121-134 with :99-118 and :55-80: one head of dimension C, non-causal SDPA (default scale 1/sqrt(C))."""
    b, c, hh, ww = x.shape
    xn = Fn.group_norm(x, 32, sd[f"{p}.group_norm.weight"], sd[f"{p}.group_norm.bias"], 1e-6).view(b, c, -1).transpose(1, 2)
    q = Fn.linear(xn, sd[f"{p}.query.weight"], sd[f"{p}.query.bias"])
    k = Fn.linear(xn, sd[f"{p}.key.weight"], sd[f"{p}.key.bias"])
    v = Fn.linear(xn, sd[f"{p}.value.weight"], sd[f"{p}.value.bias"])
    o = Fn.scaled_dot_product_attention(q.unsqueeze(1), k.unsqueeze(1), v.unsqueeze(1)).squeeze(1)
    o = Fn.linear(o, sd[f"{p}.proj_attn.weight"], sd[f"{p}.proj_attn.bias"])
    return o.transpose(1, 2).reshape(b, c, hh, ww) + x


def decode(sd: Dict[str, torch.Tensor], z: torch.Tensor) -> torch.Tensor:
    """VAE.decode (vae.py:270-274): latent (B,4,h,w) -> image (B,3,8h,8w), fp32."""
    x = z / SCALE
    x = Fn.conv2d(x, sd["post_quant_conv.weight"], sd["post_quant_conv.bias"])
    x = Fn.conv2d(x, sd["decoder.conv_in.weight"], sd["decoder.conv_in.bias"], padding=1)
    x = resblock(sd, "decoder.mid_block.resnets.0", x)
    x = attention(sd, "decoder.mid_block.attentions.0", x)
    x = resblock(sd, "decoder.mid_block.resnets.1", x)
    for j, _, _, up in decoder_blocks():
        for k in range(3):
            x = resblock(sd, f"decoder.up_blocks.{j}.resnets.{k}", x)
        if up:
            x = Fn.interpolate(x, scale_factor=2.0, mode="nearest")
            x = Fn.conv2d(x, sd[f"decoder.up_blocks.{j}.upsamplers.0.conv.weight"], sd[f"decoder.up_blocks.{j}.upsamplers.0.conv.bias"], padding=1)
    x = Fn.silu(Fn.group_norm(x, 32, sd["decoder.conv_norm_out.weight"], sd["decoder.conv_norm_out.bias"], 1e-6))
    return Fn.conv2d(x, sd["decoder.conv_out.weight"], sd["decoder.conv_out.bias"], padding=1)
