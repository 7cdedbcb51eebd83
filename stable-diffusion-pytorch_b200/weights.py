"""diffusers-layout UNet checkpoint -> reference state-dict keys.

The reference does this with two hand-unrolled 740-line tables (utils/model_converter.py:49-791 for SD-1.5 and
:793-1535 for SD-2.1).  The mapping is regular, so here it is a handful of rules driven by the architecture
description (arch.py); `tests/test_weights.py` checks that the rules reproduce the reference tables' key set
(golden key lists in tests/golden/converter_keys_*.json, generated from the unmodified reference).

Only the UNet (the hot path) is mapped; CLIP / VAE loading stays with the reference.
"""
from __future__ import annotations

from typing import Dict, Iterable, Tuple

import torch

from .arch import ResBlock, Transformer, UNetArch

_RES = (("groupnorm_1", "norm1"), ("conv_1", "conv1"), ("t_embed", "time_emb_proj"), ("groupnorm_2", "norm2"),
        ("conv_2", "conv2"), ("proj_input", "conv_shortcut"))
_TR = (("groupnorm", "norm"), ("conv_input", "proj_in"), ("conv_output", "proj_out"),
       ("transformer_block.layernorm_1", "transformer_blocks.0.norm1"),
       ("transformer_block.layernorm_2", "transformer_blocks.0.norm2"),
       ("transformer_block.layernorm_3", "transformer_blocks.0.norm3"),
       ("transformer_block.attn1.q_proj", "transformer_blocks.0.attn1.to_q"),
       ("transformer_block.attn1.k_proj", "transformer_blocks.0.attn1.to_k"),
       ("transformer_block.attn1.v_proj", "transformer_blocks.0.attn1.to_v"),
       ("transformer_block.attn1.out_proj", "transformer_blocks.0.attn1.to_out.0"),
       ("transformer_block.attn2.q_proj", "transformer_blocks.0.attn2.to_q"),
       ("transformer_block.attn2.k_proj", "transformer_blocks.0.attn2.to_k"),
       ("transformer_block.attn2.v_proj", "transformer_blocks.0.attn2.to_v"),
       ("transformer_block.attn2.out_proj", "transformer_blocks.0.attn2.to_out.0"),
       ("transformer_block.ffn.0.proj", "transformer_blocks.0.ff.net.0.proj"),
       ("transformer_block.ffn.1", "transformer_blocks.0.ff.net.2"))


def _module_pairs(a: UNetArch) -> Iterable[Tuple[str, str]]:
    """(reference module prefix, diffusers module prefix) for every parameter-holding leaf module."""
    yield "time_embedding.ffn.0", "time_embedding.linear_1"
    yield "time_embedding.ffn.2", "time_embedding.linear_2"
    yield "encoder.conv_in", "conv_in"

    def res(ref: str, dif: str):
        for r, d in _RES:
            yield f"{ref}.{r}", f"{dif}.{d}"

    def tr(ref: str, dif: str):
        for r, d in _TR:
            yield f"{ref}.{r}", f"{dif}.{d}"

    for i, st in enumerate(a.down):
        for j, (rb, tb) in enumerate(st.blocks):
            yield from res(rb.prefix, f"down_blocks.{i}.resnets.{j}")
            if tb is not None:
                yield from tr(tb.prefix, f"down_blocks.{i}.attentions.{j}")
        if st.resample is not None:
            yield st.resample.prefix, f"down_blocks.{i}.downsamplers.0.conv"
    yield from res(a.mid[0].prefix, "mid_block.resnets.0")
    yield from tr(a.mid[1].prefix, "mid_block.attentions.0")
    yield from res(a.mid[2].prefix, "mid_block.resnets.1")
    for j, st in enumerate(a.up):
        for k, (rb, tb) in enumerate(st.blocks):
            yield from res(rb.prefix, f"up_blocks.{j}.resnets.{k}")
            if tb is not None:
                yield from tr(tb.prefix, f"up_blocks.{j}.attentions.{k}")
        if st.resample is not None:
            yield st.resample.prefix, f"up_blocks.{j}.upsamplers.0.conv"
    yield "output.0", "conv_norm_out"
    yield "output.2", "conv_out"


def key_map(a: UNetArch, names: Iterable[str]) -> Dict[str, str]:
    """reference parameter name -> diffusers parameter name, for the parameter names in ``names``."""
    mods = dict(_module_pairs(a))
    out = {}
    for n in names:
        mod, leaf = n.rsplit(".", 1)
        if mod not in mods:
            raise KeyError(f"no diffusers counterpart for {n}")
        out[n] = f"{mods[mod]}.{leaf}"
    return out


def convert_state_dict(a: UNetArch, diffusers_sd: Dict[str, torch.Tensor], names_shapes) -> Dict[str, torch.Tensor]:
    """diffusers tensors -> reference-keyed tensors.  SD-2.1 stores proj_in/proj_out as Linear [C, C]; the reference
    keeps them as 1x1 conv weights [C, C, 1, 1] (model_converter.py:822,844), so 2-D tensors are unsqueezed."""
    km = key_map(a, [n for n, _ in names_shapes])
    out = {}
    for n, shape in names_shapes:
        t = diffusers_sd[km[n]]
        if t.dim() == 2 and len(shape) == 4:
            t = t[:, :, None, None]
        if tuple(t.shape) != tuple(shape):
            raise ValueError(f"{km[n]}: shape {tuple(t.shape)} does not match {n} {tuple(shape)}")
        out[n] = t
    return out


def load_unet_state_dict(path: str, a: UNetArch, device: str = "cpu") -> Dict[str, torch.Tensor]:
    """Read ``diffusion_pytorch_model.safetensors`` (or a torch ``.ckpt`` holding the same keys) and return the
    reference-keyed state dict (what load_unet_weights_v1_5 / v2_1 return under ['unet'])."""
    from .arch import param_spec
    if path.endswith(".safetensors"):
        from safetensors.torch import load_file
        sd = load_file(path, device=device)
    else:
        sd = torch.load(path, map_location=device)
    return convert_state_dict(a, sd, param_spec(a))
