"""Golden fixtures of the text encoders from the UNMODIFIED reference (build container only):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_clip.py

oracle.clip_oracle.make_state_dict weights are loaded (strict=True) into the reference ``CLIPTextModel`` (models/clip/openclip.py)
and ``TextEncoder`` (models/clip/clip.py) at reduced sizes; outputs -> tests/golden/clip_golden.npz."""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("SD_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from models.clip.openclip import CLIPTextConfig, CLIPTextModel          # noqa: E402  (reference)
from oracle import clip_oracle as CO                                       # noqa: E402

torch.set_grad_enabled(False)


def reference_text_encoder(cfg):
    """models/clip/clip.py imports ``from ..activation_fn import QuickGELU`` relative to a package above ``models``; load it
    under a synthetic parent package so that the UNMODIFIED file executes."""
    import importlib.util
    pkg = types.ModuleType("refpkg"); pkg.__path__ = [REF]
    sys.modules["refpkg"] = pkg
    for name, path in (("refpkg.activation_fn", os.path.join(REF, "models", "activation_fn.py")),):
        spec = importlib.util.spec_from_file_location(name, path)
        m = importlib.util.module_from_spec(spec); sys.modules[name] = m; spec.loader.exec_module(m)
    sub = types.ModuleType("refpkg.clip"); sub.__path__ = [os.path.join(REF, "models", "clip")]
    sys.modules["refpkg.clip"] = sub
    for name, fn in (("refpkg.clip.attention", "attention.py"), ("refpkg.clip.clip", "clip.py")):
        spec = importlib.util.spec_from_file_location(name, os.path.join(REF, "models", "clip", fn))
        m = importlib.util.module_from_spec(spec); sys.modules[name] = m; spec.loader.exec_module(m)
    TE = sys.modules["refpkg.clip.clip"].TextEncoder
    net = TE(n_vocab=cfg["vocab"], embed_dim=cfg["hidden"], max_len=cfg["max_len"])
    net.encoder_layers = torch.nn.ModuleList(list(net.encoder_layers)[:cfg["layers"]])      # reduced depth (the class hard-codes 12)
    return net.eval()


def main():
    out = {}
    g = torch.Generator().manual_seed(99)
    # --- OpenCLIP-style CLIPTextModel
    cfg = CO.SMALL_OPENCLIP
    rc = CLIPTextConfig(hidden_size=cfg["hidden"], intermediate_size=cfg["inter"], num_attention_heads=cfg["heads"],
                        num_hidden_layers=cfg["layers"], vocab_size=cfg["vocab"], max_position_embeddings=cfg["max_len"])
    ref = CLIPTextModel(rc).eval()
    sd = CO.make_state_dict(5, **cfg)
    assert list(ref.state_dict().keys()) == list(sd.keys())
    ref.load_state_dict(sd, strict=True)
    ids = torch.randint(0, cfg["vocab"], (3, 77), generator=g)
    y = ref(ids)
    out["openclip_ids"], out["openclip_out"] = ids.numpy(), y.numpy()
    e = float((CO.text_forward(sd, ids, **cfg) - y).norm() / y.norm())
    print(f"openclip small: out {tuple(y.shape)}; oracle rel-L2 {e:.2e}")
    assert e < 1e-5
    # --- CLIP-L-style TextEncoder (QuickGELU)
    cfg = CO.SMALL_CLIP
    ref2 = reference_text_encoder(cfg)
    sd2 = CO.make_state_dict(6, **cfg)
    assert list(ref2.state_dict().keys()) == list(sd2.keys()), (list(ref2.state_dict().keys())[:8], list(sd2.keys())[:8])
    ref2.load_state_dict(sd2, strict=True)
    ids2 = torch.randint(0, cfg["vocab"], (2, 77), generator=g)
    y2 = ref2(ids2)
    out["clip_ids"], out["clip_out"] = ids2.numpy(), y2.numpy()
    e = float((CO.text_forward(sd2, ids2, **cfg) - y2).norm() / y2.norm())
    print(f"clip small: out {tuple(y2.shape)}; oracle rel-L2 {e:.2e}")
    assert e < 1e-5
    np.savez_compressed(os.path.join(HERE, "clip_golden.npz"), **out)
    print("wrote clip_golden.npz")


if __name__ == "__main__":
    main()
