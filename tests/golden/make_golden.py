"""Generate the golden fixtures in tests/golden/ by running the UNMODIFIED reference.

Run in the build container (the only place /root/reference exists):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Nothing under tests/ reads /root/reference at test time; the tests read the .npz files this
script wrote.  Weights come from oracle.unet_oracle.make_state_dict(seed) and are loaded into the
reference ``UNet`` with ``strict=True`` (which also pins the state-dict key contract).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("SD_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)          # the reference imports itself as top-level `models` / `utils`

from models.scheduler import DDIMSampler, DDPMSampler      # noqa: E402  (reference)
from models.unet.unet import UNet                          # noqa: E402  (reference)
from oracle import unet_oracle as UO                       # noqa: E402

torch.set_grad_enabled(False)


def sampler_golden():
    out = {}
    d = DDIMSampler()
    out["betas"], out["alphas"], out["alphas_hat"] = d.betas.numpy(), d.alphas.numpy(), d.alphas_hat.numpy()
    dc = DDIMSampler(use_cosine_schedule=True)
    out["cos_alphas"], out["cos_alphas_hat"] = dc.alphas.numpy(), dc.alphas_hat.numpy()
    for n in (1, 7, 10, 20, 50, 250, 999):
        d._set_inference_steps(n)
        out[f"ddim_ts_{n}"] = d.timesteps.numpy()
        out[f"ddim_prev_{n}"] = np.array([int(d._get_prev_timestep(t)) for t in d.timesteps], dtype=np.int64)
        p = DDPMSampler()
        p._set_inference_steps(n)
        out[f"ddpm_ts_{n}"] = p.timesteps.numpy()
    d._set_inference_steps(50)
    for s in (0.3, 0.6, 0.8, 1.0):
        d._set_inference_steps(50)
        d.set_strength(s)
        out[f"ddim_strength_{s}"] = d.timesteps.numpy()

    g = torch.Generator().manual_seed(7)
    x = torch.randn((2, 4, 16, 16), generator=g)
    pred2 = torch.randn((4, 4, 16, 16), generator=g)
    noise = torch.randn((2, 4, 16, 16), generator=g)
    out["x"], out["pred2"], out["noise"] = x.numpy(), pred2.numpy(), noise.numpy()
    u, c = pred2.chunk(2)
    eps = u + 7.5 * (c - u)                                   # models/diffusion.py:234-235
    out["cfg_eps"] = eps.numpy()
    for ptype in ("epsilon", "v_prediction"):
        for n in (10, 50):
            s = DDIMSampler(prediction_type=ptype)
            s._set_inference_steps(n)
            for t in s.timesteps[[0, 1, n // 2, n - 1]]:
                y = s.reverse_process(x, t.unsqueeze(0), eps)
                out[f"ddim_{ptype}_{n}_{int(t)}"] = y.numpy()
    # eta > 0: the reference draws randn_like(x_t) from the global generator
    s = DDIMSampler()
    s._set_inference_steps(10)
    for t in s.timesteps[[0, 5, 9]]:
        torch.manual_seed(99)
        y = s.reverse_process(x, t.unsqueeze(0), eps, eta=0.5)
        torch.manual_seed(99)
        out[f"ddim_eta_noise_{int(t)}"] = torch.randn_like(x).numpy()
        out[f"ddim_eta0.5_10_{int(t)}"] = y.numpy()
    sc = DDIMSampler(use_cosine_schedule=True)
    sc._set_inference_steps(10)
    for t in sc.timesteps[[0, 9]]:
        out[f"ddim_cos_10_{int(t)}"] = sc.reverse_process(x, t.unsqueeze(0), eps).numpy()
    p = DDPMSampler()
    p._set_inference_steps(10)
    for t in p.timesteps[[0, 4, 9]]:
        torch.manual_seed(123)
        y = p.reverse_process(x, t.unsqueeze(0), eps)
        torch.manual_seed(123)
        out[f"ddpm_noise_{int(t)}"] = torch.randn(x.shape).numpy()
        out[f"ddpm_10_{int(t)}"] = y.numpy()
    tt = torch.tensor([3, 977])
    xt, _ = d.forward_process(x, tt, noise)
    out["fwd_t"], out["fwd_xt"] = tt.numpy(), xt.numpy()
    aT, sT = 0.0047 ** 0.5, (1 - 0.0047) ** 0.5
    out["onestep_x0"] = ((x - sT * eps) / aT).numpy()         # models/diffusion.py:111-113
    np.savez_compressed(os.path.join(HERE, "sampler_golden.npz"), **out)
    print("sampler_golden.npz", len(out), "arrays")


def build_ref_unet(cfg, seed):
    sd = UO.make_state_dict(seed, **cfg)
    net = UNet(attention_head_dim=cfg["attention_head_dim"], cross_attention_dim=cfg["cross_attention_dim"]).eval()
    net.load_state_dict(sd, strict=True)
    keys = list(net.state_dict().keys())
    assert keys == [n for n, _ in UO.param_spec(**cfg)], "param_spec order/name drift vs reference"
    return net, sd


def unet_golden():
    # SD1.5-arch
    net, sd = build_ref_unet(UO.SD15, seed=0)
    out = {"n_params": np.int64(sum(p.numel() for p in net.parameters()))}
    lat, ctx = UO.synthetic_inputs(1, 16, 16, 768, seed=1234)
    out["lat16"], out["ctx"] = lat.numpy(), ctx.numpy()
    for t in (981, 1):
        y = net(lat.repeat(2, 1, 1, 1), torch.tensor([t]), ctx)
        out[f"out16_t{t}"] = y.numpy()
    # per-sample timesteps (training-style call) on 8x8
    lat8, ctx8 = UO.synthetic_inputs(1, 8, 8, 768, seed=5)
    out["lat8"], out["ctx8"] = lat8.numpy(), ctx8.numpy()
    out["out8_t500_20"] = net(lat8.repeat(2, 1, 1, 1), torch.tensor([500, 20]), ctx8).numpy()
    # broadcast context (one-step path): context batch 1 against latent batch 2
    out["out8_ctx1_t999"] = net(lat8.repeat(2, 1, 1, 1), torch.tensor([999]), ctx8[:1]).numpy()
    # non-square latent
    g = torch.Generator().manual_seed(11)
    latr = torch.randn((1, 4, 8, 16), generator=g)
    out["lat8x16"] = latr.numpy()
    out["out8x16_t301"] = net(latr, torch.tensor([301]), ctx8[1:]).numpy()

    # full loop, config 1 scaled down: 16x16 latent, DDIM-10, CFG 7.5, B=1 (diffusion.py:223-236)
    smp = DDIMSampler()
    smp._set_inference_steps(10)
    latent = lat.clone()
    traj = []
    for ts in smp.timesteps:
        ts = ts.unsqueeze(0)
        o = net(latent.repeat(2, 1, 1, 1), ts, ctx)
        u, c = o.chunk(2)
        eps = u + 7.5 * (c - u)
        latent = smp.reverse_process(latent, ts, eps)
        traj.append(latent.numpy().copy())
    out["loop16_ddim10_step1"] = traj[0]
    out["loop16_ddim10_final"] = traj[-1]
    np.savez_compressed(os.path.join(HERE, "unet_sd15_golden.npz"), **out)
    print("unet_sd15_golden.npz", {k: v.shape for k, v in out.items()})
    del net, sd

    # SD2.1-arch (heads 5/10/20/20 x 64, context 1024, v-prediction)
    net, sd = build_ref_unet(UO.SD21, seed=1)
    out = {"n_params": np.int64(sum(p.numel() for p in net.parameters()))}
    lat, ctx = UO.synthetic_inputs(1, 16, 16, 1024, seed=4321)
    out["lat16"], out["ctx"] = lat.numpy(), ctx.numpy()
    out["out16_t961"] = net(lat.repeat(2, 1, 1, 1), torch.tensor([961]), ctx).numpy()
    smp = DDIMSampler(prediction_type="v_prediction")
    smp._set_inference_steps(5)
    latent = lat.clone()
    for ts in smp.timesteps:
        ts = ts.unsqueeze(0)
        o = net(latent.repeat(2, 1, 1, 1), ts, ctx)
        u, c = o.chunk(2)
        latent = smp.reverse_process(latent, ts, u + 7.5 * (c - u))
    out["loop16_ddim5_v_final"] = latent.numpy()
    # SwiftBrush one-step: t = timesteps[0] of a fresh sampler (=999), context batch 1, no CFG
    fresh = DDIMSampler()
    t0 = fresh.timesteps[0].unsqueeze(0)
    out["onestep_t"] = t0.numpy()
    lat2 = torch.randn((2, 4, 8, 8), generator=torch.Generator().manual_seed(3))
    out["onestep_lat"] = lat2.numpy()
    pn = net(lat2, t0, ctx[:1])
    aT, sT = 0.0047 ** 0.5, (1 - 0.0047) ** 0.5
    out["onestep_pred"] = pn.numpy()
    out["onestep_x0"] = ((lat2 - sT * pn) / aT).numpy()
    np.savez_compressed(os.path.join(HERE, "unet_sd21_golden.npz"), **out)
    print("unet_sd21_golden.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    torch.manual_seed(0)
    sampler_golden()
    unet_golden()
