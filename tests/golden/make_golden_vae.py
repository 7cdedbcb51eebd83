"""Golden fixtures of the VAE decoder from the UNMODIFIED reference (run in the build container only):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_vae.py

Weights: oracle.vae_oracle.make_state_dict(seed), loaded into the reference ``VAE`` with strict=True (pins the full 248-key
state-dict contract, encoder included); outputs of ``VAE.decode`` (models/vae/vae.py:270-274) -> tests/golden/vae_golden.npz."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("SD_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from models.vae.vae import VAE                      # noqa: E402  (reference)
from oracle import vae_oracle as VO                 # noqa: E402

torch.set_grad_enabled(False)


def main():
    sd = VO.make_state_dict(3)
    ref = VAE().eval()
    assert list(ref.state_dict().keys()) == [n for n, _ in VO.param_spec()], "parameter names / registration order differ from the reference"
    ref.load_state_dict(sd, strict=True)
    out = {}
    g = torch.Generator().manual_seed(2024)
    for name, shape in (("z8", (2, 4, 8, 8)), ("z8x16", (1, 4, 8, 16)), ("z24", (1, 4, 24, 24))):
        z = torch.randn(shape, generator=g) * 0.18215 * 4.0          # latents at the scale the sampler hands to decode
        out[name] = z.numpy()
        y = ref.decode(z)
        out[f"img_{name}"] = y.numpy()
        e = float((VO.decode(sd, z) - y).norm() / y.norm())
        print(f"{name}: reference image {tuple(y.shape)}, |y| mean {float(y.abs().mean()):.3f}; oracle restatement rel-L2 {e:.2e}")
        assert e < 2e-5
    np.savez_compressed(os.path.join(HERE, "vae_golden.npz"), **out)
    print("wrote vae_golden.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
