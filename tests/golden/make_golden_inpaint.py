"""Golden vectors for the inpainting loop body (models/diffusion.py:387-398), produced by the UNMODIFIED reference sampler.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_inpaint.py      (build container only: needs /root/reference)

The statements below are the reference's loop body with the UNet call replaced by a fixed synthetic prediction.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("SD_REFERENCE", "/root/reference")
sys.path.insert(0, REF)

from models.scheduler import DDIMSampler, DDPMSampler      # noqa: E402  (reference)

torch.set_grad_enabled(False)


def main():
    g = torch.Generator().manual_seed(11)
    B, C, h, w = 2, 4, 16, 24
    out = {}
    latent = torch.randn((B, C, h, w), generator=g)
    pred2 = torch.randn((2 * B, C, h, w), generator=g)
    encoded = torch.randn((1, C, h, w), generator=g)
    mask = torch.rand((1, 1, h, w), generator=g) > 0.6          # downsampled_mask: bool, True = repaint
    out.update(latent=latent.numpy(), pred2=pred2.numpy(), encoded=encoded.numpy(), mask=mask.numpy())
    cfg_scale = 7.5
    for ptype in ("epsilon", "v_prediction"):
        s = DDIMSampler(prediction_type=ptype)
        s._set_inference_steps(50)
        s.set_strength(0.8)
        for t in (int(s.timesteps[0]), int(s.timesteps[17]), int(s.timesteps[-1])):
            timestep = torch.tensor([t])
            # ---- models/diffusion.py:387-398, verbatim semantics
            cond_output, uncond_output = pred2.chunk(2)
            pred_noise = cfg_scale * (cond_output - uncond_output) + cond_output
            noised_orig_img, _ = s.forward_process(encoded, timestep, pred_noise)
            lf = torch.where(~mask.repeat(1, C, 1, 1), noised_orig_img, latent)
            lf = s.reverse_process(lf, timestep, pred_noise)
            out[f"{ptype}_{t}_cfg"] = lf.numpy()
            # without CFG: the prediction is the first half
            pn = pred2[:B]
            no, _ = s.forward_process(encoded, timestep, pn)
            lf = torch.where(~mask.repeat(1, C, 1, 1), no, latent)
            out[f"{ptype}_{t}_nocfg"] = s.reverse_process(lf, timestep, pn).numpy()
        out[f"{ptype}_ts"] = s.timesteps.numpy()
    # ---- the same loop body with sampler='ddpm' (models/diffusion.py:314-316): reverse_process draws its noise from the global
    # generator (ddpm.py:80), so the draw is reproduced by re-seeding
    s = DDPMSampler()
    s._set_inference_steps(50)
    s.set_strength(0.8)
    for t in (int(s.timesteps[0]), int(s.timesteps[17]), int(s.timesteps[-1])):
        timestep = torch.tensor([t])
        cond_output, uncond_output = pred2.chunk(2)
        pred_noise = cfg_scale * (cond_output - uncond_output) + cond_output
        noised_orig_img, _ = s.forward_process(encoded, timestep, pred_noise)
        lf = torch.where(~mask.repeat(1, C, 1, 1), noised_orig_img, latent)
        torch.manual_seed(1000 + t)
        out[f"ddpm_{t}_cfg"] = s.reverse_process(lf, t, pred_noise).numpy()
        torch.manual_seed(1000 + t)
        out[f"ddpm_{t}_noise"] = torch.randn(lf.shape, dtype=lf.dtype).numpy()
    out["ddpm_ts"] = s.timesteps.numpy()
    np.savez_compressed(os.path.join(HERE, "inpaint_golden.npz"), **out)
    print("wrote inpaint_golden.npz", {k: v.shape for k, v in out.items() if k.endswith("_ts")})


if __name__ == "__main__":
    main()
