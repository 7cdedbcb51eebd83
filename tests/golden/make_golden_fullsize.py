"""Full-size golden trajectories from the UNMODIFIED reference (slow: minutes of CPU).

    python tests/golden/make_golden_fullsize.py sd15   # BASELINE config 2: SD1.5-arch, 64x64, DDIM-50, CFG 7.5, B=1
    python tests/golden/make_golden_fullsize.py sd21   # BASELINE config 4 at B=1: SD2.1-arch, 96x96, DDIM-50, v-pred

Weights: oracle.unet_oracle.make_state_dict(seed); inputs: oracle.unet_oracle.synthetic_inputs.
Writes tests/golden/loop_<arch>_full.npz (initial latent, context, latents after steps 1/10/25/50).
"""
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.environ.get("SD_REFERENCE", "/root/reference"))

from models.scheduler import DDIMSampler            # noqa: E402  (reference)
from models.unet.unet import UNet                   # noqa: E402  (reference)
from oracle import unet_oracle as UO                # noqa: E402

torch.set_grad_enabled(False)
torch.set_num_threads(int(os.environ.get("NTHREADS", "4")))
which = sys.argv[1]
if which == "sd15":
    cfg, seed, hw, dctx, ptype, inseed = UO.SD15, 0, 64, 768, "epsilon", 1234
else:
    cfg, seed, hw, dctx, ptype, inseed = UO.SD21, 1, 96, 1024, "v_prediction", 4321
net = UNet(attention_head_dim=cfg["attention_head_dim"], cross_attention_dim=cfg["cross_attention_dim"]).eval()
net.load_state_dict(UO.make_state_dict(seed, **cfg), strict=True)
lat, ctx = UO.synthetic_inputs(1, hw, hw, dctx, seed=inseed)
smp = DDIMSampler(prediction_type=ptype)
smp._set_inference_steps(50)
out = {"lat": lat.numpy(), "ctx": ctx.numpy(), "timesteps": smp.timesteps.numpy()}
latent = lat.clone()
t0 = time.time()
for i, ts in enumerate(smp.timesteps):
    ts = ts.unsqueeze(0)
    o = net(latent.repeat(2, 1, 1, 1), ts, ctx)
    u, c = o.chunk(2)
    latent = smp.reverse_process(latent, ts, u + 7.5 * (c - u))
    if i == 0:
        out["unet_out_step1"] = o.numpy().copy()
    if i + 1 in (1, 10, 25, 50):
        out[f"latent_step{i + 1}"] = latent.numpy().copy()
    print(i, time.time() - t0, float(latent.abs().mean()), flush=True)
np.savez_compressed(os.path.join(HERE, f"loop_{which}_full.npz"), **out)
print("done", which)
