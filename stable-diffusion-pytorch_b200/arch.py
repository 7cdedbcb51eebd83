"""Static description of the SD-1.5 / SD-2.1 conditional UNet as a flat block list.

This is the product-side restatement of the reference's module tree
(models/unet/unet.py:253-401): it yields, in the reference's registration order, every
parameter name + shape (the state-dict contract) and the block sequence the
step program is generated from.  No torch ops here.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Sequence, Tuple, Union

BLOCK_OUT = (320, 640, 1280, 1280)      # hard-wired in the reference too (unet.py:300,384,390,399)


@dataclass
class ResBlock:
    prefix: str
    cin: Tuple[int, ...]        # channel counts of the (concatenated) input sources, x first
    cout: int
    eps: float
    tb_offset: int = 0          # column offset of this block's Linear(SiLU(t_emb)) in the packed table

    @property
    def cin_total(self):
        return sum(self.cin)

    @property
    def has_proj(self):
        return self.cin_total != self.cout


@dataclass
class Transformer:
    prefix: str
    c: int
    heads: int
    dctx: int
    index: int = 0              # running index (cross-attention K/V cache slot)


@dataclass
class Conv:
    prefix: str
    cin: int
    cout: int
    stride: int = 1
    upsample: bool = False


@dataclass
class Stage:
    blocks: list = field(default_factory=list)     # [(ResBlock, Transformer|None), ...]
    resample: Union[Conv, None] = None


@dataclass
class UNetArch:
    in_channels: int
    out_channels: int
    t_embed_dim: int
    dctx: int
    heads: List[int]
    eps: float
    down: List[Stage]
    mid: list
    up: List[Stage]
    res_blocks: List[ResBlock]
    transformers: List[Transformer]
    tb_total: int

    @property
    def temb(self):
        return 4 * self.t_embed_dim


def _as_list(v, n):
    return [v] * n if isinstance(v, int) else list(v)


def build_arch(attention_head_dim=8, cross_attention_dim=768, in_channels=4, out_channels=4,
               block_out_channels: Sequence[int] = BLOCK_OUT, down_block_types=None, t_embed_dim=320,
               num_attention_heads=None, eps=1e-5) -> UNetArch:
    ch = tuple(block_out_channels)
    if ch != BLOCK_OUT:
        # the reference hard-codes 1280 for the bottleneck and 320 for the decoder/head (SURVEY §2 notes)
        raise ValueError(f"block_out_channels must be {list(BLOCK_OUT)} (the reference only works with these), got {list(ch)}")
    n = len(ch)
    heads = _as_list(attention_head_dim if num_attention_heads is None else num_attention_heads, n)
    dctx_l = _as_list(cross_attention_dim, n)
    if len(set(dctx_l)) != 1:
        raise ValueError("per-level cross_attention_dim must be uniform")
    dctx = dctx_l[0]
    for c, h in zip(ch, heads):
        if c % h:
            raise ValueError('Number of heads must be divisible by Embedding Dimension')      # unet.py:97-98
    res, trs = [], []

    def mk_res(prefix, cin, cout, e):
        r = ResBlock(prefix, tuple(cin), cout, e)
        res.append(r)
        return r

    def mk_tr(prefix, c, h):
        t = Transformer(prefix, c, h, dctx, index=len(trs))
        trs.append(t)
        return t

    down = []
    cin_l = (ch[0],) + ch
    for i in range(n):
        st = Stage()
        for j in range(2):
            cin = cin_l[i] if j == 0 else ch[i]
            # unet.py:272-273: levels 0..2 build their ResBlocks without forwarding eps (default 1e-5)
            r = mk_res(f"encoder.down.{i}.block.{j}.0", (cin,), ch[i], 1e-5 if i != n - 1 else eps)
            t = mk_tr(f"encoder.down.{i}.block.{j}.1", ch[i], heads[i]) if i != n - 1 else None
            st.blocks.append((r, t))
        if i != n - 1:
            st.resample = Conv(f"encoder.down.{i}.downsample.conv", ch[i], ch[i], stride=2)
        down.append(st)

    mid = [mk_res("bottleneck.0", (1280,), 1280, 1e-5), mk_tr("bottleneck.1", ch[-1], heads[-1]),
           mk_res("bottleneck.2", (1280,), 1280, 1e-5)]

    # skip channel stack, in push order (unet.py:286-294)
    skips = [ch[0]]
    for i in range(n):
        skips += [ch[i], ch[i]]
        if i != n - 1:
            skips.append(ch[i])
    up = []
    x_c = ch[-1]
    for j, i in enumerate(reversed(range(n))):
        st = Stage()
        for k in range(3):
            sk = skips.pop()
            r = mk_res(f"decoder.up.{j}.block.{k}.0", (x_c, sk), ch[i], eps)
            t = mk_tr(f"decoder.up.{j}.block.{k}.1", ch[i], heads[i]) if i != n - 1 else None
            st.blocks.append((r, t))
            x_c = ch[i]
        if i != 0:
            st.resample = Conv(f"decoder.up.{j}.upsample.conv", ch[i], ch[i], upsample=True)
        up.append(st)
    off = 0
    for r in res:
        r.tb_offset = off
        off += r.cout
    return UNetArch(in_channels, out_channels, t_embed_dim, dctx, heads, eps, down, mid, up, res, trs, off)


def _res_params(r: ResBlock, temb):
    p, ci, co = r.prefix, r.cin_total, r.cout
    s = [(f"{p}.groupnorm_1.weight", (ci,)), (f"{p}.groupnorm_1.bias", (ci,)),
         (f"{p}.conv_1.weight", (co, ci, 3, 3)), (f"{p}.conv_1.bias", (co,)),
         (f"{p}.groupnorm_2.weight", (co,)), (f"{p}.groupnorm_2.bias", (co,)),
         (f"{p}.conv_2.weight", (co, co, 3, 3)), (f"{p}.conv_2.bias", (co,)),
         (f"{p}.t_embed.weight", (co, temb)), (f"{p}.t_embed.bias", (co,))]
    if r.has_proj:
        s += [(f"{p}.proj_input.weight", (co, ci, 1, 1)), (f"{p}.proj_input.bias", (co,))]
    return s


def _tr_params(t: Transformer):
    p, c, d = t.prefix, t.c, t.dctx
    b = f"{p}.transformer_block"
    return [(f"{p}.groupnorm.weight", (c,)), (f"{p}.groupnorm.bias", (c,)),
            (f"{p}.conv_input.weight", (c, c, 1, 1)), (f"{p}.conv_input.bias", (c,)),
            (f"{b}.layernorm_1.weight", (c,)), (f"{b}.layernorm_1.bias", (c,)),
            (f"{b}.attn1.q_proj.weight", (c, c)), (f"{b}.attn1.k_proj.weight", (c, c)),
            (f"{b}.attn1.v_proj.weight", (c, c)), (f"{b}.attn1.out_proj.weight", (c, c)),
            (f"{b}.attn1.out_proj.bias", (c,)),
            (f"{b}.layernorm_2.weight", (c,)), (f"{b}.layernorm_2.bias", (c,)),
            (f"{b}.attn2.q_proj.weight", (c, c)), (f"{b}.attn2.k_proj.weight", (c, d)),
            (f"{b}.attn2.v_proj.weight", (c, d)), (f"{b}.attn2.out_proj.weight", (c, c)),
            (f"{b}.attn2.out_proj.bias", (c,)),
            (f"{b}.layernorm_3.weight", (c,)), (f"{b}.layernorm_3.bias", (c,)),
            (f"{b}.ffn.0.proj.weight", (8 * c, c)), (f"{b}.ffn.0.proj.bias", (8 * c,)),
            (f"{b}.ffn.1.weight", (c, 4 * c)), (f"{b}.ffn.1.bias", (c,)),
            (f"{p}.conv_output.weight", (c, c, 1, 1)), (f"{p}.conv_output.bias", (c,))]


def param_spec(a: UNetArch):
    """(name, shape) of every parameter, in the reference's state-dict order."""
    te = a.temb
    s = [("time_embedding.ffn.0.weight", (te, a.t_embed_dim)), ("time_embedding.ffn.0.bias", (te,)),
         ("time_embedding.ffn.2.weight", (te, te)), ("time_embedding.ffn.2.bias", (te,)),
         ("encoder.conv_in.weight", (BLOCK_OUT[0], a.in_channels, 3, 3)), ("encoder.conv_in.bias", (BLOCK_OUT[0],))]

    def stage(st: Stage):
        out = []
        for r, t in st.blocks:
            out += _res_params(r, te)
            if t is not None:
                out += _tr_params(t)
        if st.resample is not None:
            c = st.resample
            out += [(f"{c.prefix}.weight", (c.cout, c.cin, 3, 3)), (f"{c.prefix}.bias", (c.cout,))]
        return out

    for st in a.down:
        s += stage(st)
    s += _res_params(a.mid[0], te) + _tr_params(a.mid[1]) + _res_params(a.mid[2], te)
    for st in a.up:
        s += stage(st)
    s += [("output.0.weight", (320,)), ("output.0.bias", (320,)),
          ("output.2.weight", (a.out_channels, 320, 3, 3)), ("output.2.bias", (a.out_channels,))]
    return s


# ======================================================================================
# VAE (reference models/vae/vae.py): parameter contract and decoder structure
# ======================================================================================
VAE_CH, VAE_CH_MULT = 128, (1, 2, 4, 4)


def _vae_res(p, cin, cout):
    s = [(f"{p}.norm1.weight", (cin,)), (f"{p}.norm1.bias", (cin,)), (f"{p}.conv1.weight", (cout, cin, 3, 3)), (f"{p}.conv1.bias", (cout,)),
         (f"{p}.norm2.weight", (cout,)), (f"{p}.norm2.bias", (cout,)), (f"{p}.conv2.weight", (cout, cout, 3, 3)), (f"{p}.conv2.bias", (cout,))]
    if cin != cout:
        s += [(f"{p}.conv_shortcut.weight", (cout, cin, 1, 1)), (f"{p}.conv_shortcut.bias", (cout,))]
    return s


def _vae_attn(p, c):
    s = [(f"{p}.group_norm.weight", (c,)), (f"{p}.group_norm.bias", (c,))]
    for n in ("query", "key", "value", "proj_attn"):
        s += [(f"{p}.{n}.weight", (c, c)), (f"{p}.{n}.bias", (c,))]
    return s


def vae_decoder_blocks():
    """(up-block index, cin of its first ResidualBlock, cout, has upsampler) in forward order (vae.py:208-224)."""
    out, block_in = [], VAE_CH * VAE_CH_MULT[-1]
    for j, i in enumerate(reversed(range(len(VAE_CH_MULT)))):
        block_out = VAE_CH * VAE_CH_MULT[i]
        out.append((j, block_in, block_out, i != 0))
        block_in = block_out
    return out


def vae_param_spec(in_channels=3, z_channels=4, out_channels=3):
    """(name, shape) of every parameter of the reference ``VAE`` in its registration order (vae.py:136-262); the encoder is part
    of the ``load_state_dict(strict=True)`` contract although only the decoder runs here."""
    ch, mult = VAE_CH, VAE_CH_MULT
    s = [("encoder.conv_in.weight", (ch, in_channels, 3, 3)), ("encoder.conv_in.bias", (ch,))]
    cur = ch
    for i, m in enumerate(mult):
        out = ch * m
        for j in range(2):
            s += _vae_res(f"encoder.down_blocks.{i}.resnets.{j}", cur if j == 0 else out, out)
        if i != len(mult) - 1:
            s += [(f"encoder.down_blocks.{i}.downsamplers.0.conv.weight", (out, out, 3, 3)), (f"encoder.down_blocks.{i}.downsamplers.0.conv.bias", (out,))]
        cur = out
    s += _vae_res("encoder.mid_block.resnets.0", cur, cur) + _vae_res("encoder.mid_block.resnets.1", cur, cur) + _vae_attn("encoder.mid_block.attentions.0", cur)
    s += [("encoder.conv_norm_out.weight", (cur,)), ("encoder.conv_norm_out.bias", (cur,)),
          ("encoder.conv_out.weight", (2 * z_channels, cur, 3, 3)), ("encoder.conv_out.bias", (2 * z_channels,))]
    top = ch * mult[-1]
    s += [("decoder.conv_in.weight", (top, z_channels, 3, 3)), ("decoder.conv_in.bias", (top,))]
    s += _vae_attn("decoder.mid_block.attentions.0", top) + _vae_res("decoder.mid_block.resnets.0", top, top) + _vae_res("decoder.mid_block.resnets.1", top, top)
    for j, cin, cout, up in vae_decoder_blocks():
        for k in range(3):
            s += _vae_res(f"decoder.up_blocks.{j}.resnets.{k}", cin if k == 0 else cout, cout)
        if up:
            s += [(f"decoder.up_blocks.{j}.upsamplers.0.conv.weight", (cout, cout, 3, 3)), (f"decoder.up_blocks.{j}.upsamplers.0.conv.bias", (cout,))]
    s += [("decoder.conv_norm_out.weight", (ch,)), ("decoder.conv_norm_out.bias", (ch,)),
          ("decoder.conv_out.weight", (out_channels, ch, 3, 3)), ("decoder.conv_out.bias", (out_channels,))]
    s += [("quant_conv.weight", (2 * z_channels, 2 * z_channels, 1, 1)), ("quant_conv.bias", (2 * z_channels,)),
          ("post_quant_conv.weight", (z_channels, z_channels, 1, 1)), ("post_quant_conv.bias", (z_channels,))]
    return s
