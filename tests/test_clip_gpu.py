"""GPU parity of the text encoders against golden outputs of the UNMODIFIED reference
(tests/golden/make_golden_clip.py -> clip_golden.npz) and, at full OpenCLIP-H / CLIP-L size, against the CPU oracle.
Gates: fp32 rel-L2 <= 1e-4, bf16 <= 1e-2."""
import os

import numpy as np
import pytest
import torch

from oracle import clip_oracle as CO
from stable_diffusion_pytorch_b200 import CLIPTextConfig, CLIPTextModel, OpenCLIP, TextEncoder

pytestmark = pytest.mark.gpu
FP32_TOL, BF16_TOL = 1e-4, 1e-2


def rel_l2(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def _openclip(cfg):
    return CLIPTextModel(CLIPTextConfig(hidden_size=cfg["hidden"], intermediate_size=cfg["inter"], num_attention_heads=cfg["heads"],
                                        num_hidden_layers=cfg["layers"], vocab_size=cfg["vocab"], max_position_embeddings=cfg["max_len"]))


@pytest.mark.parametrize("precision,tol", [("fp32", FP32_TOL), ("bf16", BF16_TOL)])
def test_text_encoders_match_reference_golden(golden_dir, dev, precision, tol):
    g = np.load(os.path.join(golden_dir, "clip_golden.npz"))
    for tag, cfg, seed, net in (("openclip", CO.SMALL_OPENCLIP, 5, _openclip(CO.SMALL_OPENCLIP)),
                                ("clip", CO.SMALL_CLIP, 6, TextEncoder(CO.SMALL_CLIP["vocab"], CO.SMALL_CLIP["hidden"], 77, CO.SMALL_CLIP["layers"]))):
        sd = CO.make_state_dict(seed, **cfg)
        net.load_state_dict(sd, strict=True)
        net = net.to(dev).eval().set_precision(precision)
        ids = torch.from_numpy(g[f"{tag}_ids"]).to(dev)
        with torch.no_grad():
            y = net(ids)
            e = rel_l2(y.cpu().numpy(), g[f"{tag}_out"])
            print(f"text encoder {tag} {precision}: {tuple(y.shape)} rel-L2 {e:.3e}")
            assert y.shape == g[f"{tag}_out"].shape and y.dtype == torch.float32 and e < tol
            assert torch.equal(net(ids), y) and torch.equal(net(ids), y)             # graph capture + replay
            # causality: changing a LATER token must not change earlier positions (lookahead_mask=True)
            ids2 = ids.clone()
            ids2[:, 40:] = (ids2[:, 40:] + 7) % cfg["vocab"]
            y2 = net(ids2)
            assert torch.equal(y2[:, :40], y[:, :40]) and not torch.equal(y2[:, 40:], y[:, 40:])
            # shorter sequences (S < 77) go through their own plan
            ys = net(ids[:1, :20])
            assert rel_l2(ys.cpu().numpy(), g[f"{tag}_out"][:1, :20]) < tol


def test_openclip_h_full_size_vs_oracle(dev):
    """The text tower models/diffusion.py:190-200 runs (OpenCLIP ViT-H: 1024 wide, 16 heads, 23 layers), cond + uncond prompts."""
    cfg = dict(CO.OPENCLIP_H, vocab=4096)                       # full architecture, reduced vocabulary table (embedding rows are a gather)
    sd = CO.make_state_dict(8, **cfg)
    wrap = OpenCLIP.__new__(OpenCLIP)
    torch.nn.Module.__init__(wrap)
    wrap.text_model = _openclip(cfg)
    wrap.load_state_dict({f"text_model.{k}": v for k, v in sd.items()}, strict=True)
    wrap = wrap.to(dev).eval()
    g = torch.Generator().manual_seed(4)
    ids = torch.randint(0, cfg["vocab"], (2, 77), generator=g)
    with torch.no_grad():
        want = CO.text_forward(sd, ids, **cfg).numpy()
        for precision, tol in (("bf16", BF16_TOL), ("fp32", FP32_TOL)):
            wrap.text_model.set_precision(precision)
            y = wrap.encode_text(ids.to(dev)).cpu().numpy()
            e = rel_l2(y, want)
            print(f"OpenCLIP-H text tower {precision}: rel-L2 {e:.3e}")
            assert e < tol


def test_text_encoder_rejects_cpu(dev):
    net = _openclip(CO.SMALL_OPENCLIP)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        net(torch.zeros((1, 77), dtype=torch.long))
    with pytest.raises(RuntimeError):
        net.to(dev)(torch.zeros((1, 78), dtype=torch.long, device=dev))
