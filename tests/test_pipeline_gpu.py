"""GPU parity of the graph-replayed denoising loop (DenoiseLoop) against the golden trajectories of the
unmodified reference and against the step-by-step drop-in API."""
import os

import numpy as np
import pytest
import torch

from oracle import unet_oracle as UO
from stable_diffusion_pytorch_b200 import DDIMSampler, DDPMSampler, UNet
from oracle import sampler_oracle as SO
from stable_diffusion_pytorch_b200.pipeline import DenoiseLoop, denoise, img2img, inpaint, one_step

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def net15(dev):
    net = UNet()
    net.load_state_dict(UO.make_state_dict(0, **UO.SD15), strict=True)
    return net.to(dev).eval()


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("bf16", 1e-2)])
def test_loop_matches_golden_and_stepwise_api(net15, golden_dir, dev, precision, tol):
    g = np.load(os.path.join(golden_dir, "unet_sd15_golden.npz"))
    net15.set_precision(precision)
    lat, ctx = torch.from_numpy(g["lat16"]).to(dev), torch.from_numpy(g["ctx"]).to(dev)
    smp = DDIMSampler()
    smp._set_inference_steps(10)
    with torch.no_grad():
        final = denoise(net15, smp, lat, ctx, do_cfg=True, cfg_scale=7.5)
        e = rel_l2(final.cpu().numpy(), g["loop16_ddim10_final"])
        print(f"DenoiseLoop {precision}: final latent rel-L2 {e:.3e}")
        assert e < tol
        # the same loop through the drop-in step-by-step API must give the identical latent
        x = lat.clone()
        for ts in smp.timesteps.to(dev):
            ts = ts.unsqueeze(0)
            x = smp.reverse_process(x, ts, net15(x.repeat(2, 1, 1, 1), ts, ctx), cfg_scale=7.5)
        assert torch.equal(x, final)
        # re-running the captured loop is deterministic and reusable
        loop = DenoiseLoop(net15, smp, 1, 16, 16)
        a = loop.run(lat, ctx)
        b = loop.run(lat, ctx)
        assert torch.equal(a, final) and torch.equal(b, final)


def test_loop_ddpm_and_strength(net15, dev):
    net15.set_precision("fp32")
    lat, ctx = UO.synthetic_inputs(2, 8, 8, 768, seed=9)
    lat, ctx = lat.to(dev), ctx.to(dev)
    p = DDPMSampler()
    p._set_inference_steps(4)
    with torch.no_grad():
        torch.manual_seed(77)
        got = denoise(net15, p, lat, ctx)
        torch.manual_seed(77)
        x = lat.clone()
        for ts in p.timesteps.to(dev):
            ts = ts.unsqueeze(0)
            x = p.reverse_process(x, ts, net15(x.repeat(2, 1, 1, 1), ts, ctx), cfg_scale=7.5)
        assert torch.equal(got, x)
        d = DDIMSampler()
        d._set_inference_steps(10)
        d.set_strength(0.6)                                    # img2img: walk only the last 6 timesteps
        got = denoise(net15, d, lat, ctx)
        x = lat.clone()
        for ts in d.timesteps.to(dev):
            ts = ts.unsqueeze(0)
            x = d.reverse_process(x, ts, net15(x.repeat(2, 1, 1, 1), ts, ctx), cfg_scale=7.5)
        assert len(d.timesteps) == 6 and torch.equal(got, x)


def test_one_step_sd21(golden_dir, dev):
    g = np.load(os.path.join(golden_dir, "unet_sd21_golden.npz"))
    net = UNet(attention_head_dim=[5, 10, 20, 20], cross_attention_dim=1024)
    net.load_state_dict(UO.make_state_dict(1, **UO.SD21), strict=True)
    net = net.to(dev).eval().set_precision("fp32")
    with torch.no_grad():
        x0 = one_step(net, DDIMSampler(), torch.from_numpy(g["onestep_lat"]).to(dev), torch.from_numpy(g["ctx"])[:1].to(dev))
    assert rel_l2(x0.cpu().numpy(), g["onestep_x0"]) < 1e-4
    del net
    torch.cuda.empty_cache()


# ------------------------------------------------------------------------------------------------
# BASELINE configs at FULL size: golden trajectories from the unmodified reference
# (tests/golden/make_golden_fullsize.py; 50 DDIM steps, CFG 7.5).
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("bf16", 1e-2)])
def test_config2_sd15_512_ddim50_full(net15, golden_dir, dev, precision, tol):
    """BASELINE config 2: SD1.5-arch, 64x64 latent, DDIM-50, CFG 7.5, batch 1."""
    g = np.load(os.path.join(golden_dir, "loop_sd15_full.npz"))
    net15.set_precision(precision)
    smp = DDIMSampler()
    smp._set_inference_steps(50)
    assert np.array_equal(smp.timesteps.numpy(), g["timesteps"])            # bookkeeping: bit-exact
    lat, ctx = torch.from_numpy(g["lat"]).to(dev), torch.from_numpy(g["ctx"]).to(dev)
    loop = DenoiseLoop(net15, smp, 1, 64, 64)
    with torch.no_grad():
        loop.reset(lat, ctx)
        errs = {}
        for i in range(50):
            loop.step()
            if i + 1 in (1, 10, 25, 50):
                errs[i + 1] = rel_l2(loop.latent.cpu().numpy(), g[f"latent_step{i + 1}"])
    print(f"config 2 ({precision}) latent rel-L2 after steps 1/10/25/50: " + " ".join(f"{errs[k]:.3e}" for k in (1, 10, 25, 50)))
    assert errs[50] < tol
    net15.set_precision("bf16")


def test_config4_sd21_768_ddim50_vpred_full(golden_dir, dev):
    """BASELINE config 4 at batch 1: SD2.1-arch, 96x96 latent, DDIM-50, v-prediction, CFG 7.5."""
    g = np.load(os.path.join(golden_dir, "loop_sd21_full.npz"))
    net = UNet(attention_head_dim=[5, 10, 20, 20], cross_attention_dim=1024)
    net.load_state_dict(UO.make_state_dict(1, **UO.SD21), strict=True)
    net = net.to(dev).eval()
    lat, ctx = torch.from_numpy(g["lat"]).to(dev), torch.from_numpy(g["ctx"]).to(dev)
    res = {}
    for precision in ("bf16", "fp32"):
        net.set_precision(precision)
        smp = DDIMSampler(prediction_type="v_prediction")
        smp._set_inference_steps(50)
        loop = DenoiseLoop(net, smp, 1, 96, 96)
        with torch.no_grad():
            loop.reset(lat, ctx)
            for i in range(50):
                loop.step()
                if i + 1 in (1, 10, 25, 50):
                    res[(precision, i + 1)] = rel_l2(loop.latent.cpu().numpy(), g[f"latent_step{i + 1}"])
        print(f"config 4 ({precision}) latent rel-L2 after steps 1/10/25/50: " + " ".join(f"{res[(precision, k)]:.3e}" for k in (1, 10, 25, 50)))
        del loop
        net.invalidate()
        torch.cuda.empty_cache()
    assert res[("fp32", 50)] < 1e-4
    assert res[("bf16", 50)] < 1e-2
    del net
    torch.cuda.empty_cache()


def test_inpaint_and_img2img_loops(net15, dev):
    """Inpainting loop (models/diffusion.py:379-398) as graph replays == step-by-step drop-in API == CPU oracle loop; img2img
    (set_strength + forward_process + loop, :204-236) == its step-by-step form."""
    net15.set_precision("fp32")
    B, h, w = 1, 16, 16
    g = torch.Generator().manual_seed(21)
    lat = torch.randn((B, 4, h, w), generator=g)
    ctx = torch.randn((2 * B, 77, 768), generator=g)              # rows [cond ; uncond] in the inpaint loop
    enc = torch.randn((1, 4, h, w), generator=g)
    mask = torch.rand((1, 1, h, w), generator=g) > 0.5
    d = DDIMSampler()
    d._set_inference_steps(10)
    d.set_strength(0.4)                                           # 4 steps
    with torch.no_grad():
        got = inpaint(net15, d, lat.to(dev), ctx.to(dev), enc.to(dev), mask.to(dev), do_cfg=True, cfg_scale=7.5)
        x = lat.to(dev)
        for ts in d.timesteps.to(dev):
            ts = ts.unsqueeze(0)
            x = d.inpaint_step(x, ts, net15(x.repeat(2, 1, 1, 1), ts, ctx.to(dev)), enc.to(dev), mask.to(dev), cfg_scale=7.5)
        assert torch.equal(got, x)
    # CPU oracle: the reference's statements with the oracle UNet
    sd = UO.make_state_dict(0, **UO.SD15)
    _, alphas, a_hat = SO.schedule_fp32()
    xo = lat.clone()
    for t in d.timesteps.tolist():
        pred = UO.unet_forward(sd, xo.repeat(2, 1, 1, 1), torch.tensor([t]), ctx, **UO.SD15)
        xo = torch.from_numpy(SO.inpaint_step(xo.numpy(), t, pred.numpy(), enc.numpy(), mask[0, 0].numpy(), 7.5, alphas, a_hat, 1000, 10))
    e = rel_l2(got.cpu().numpy(), xo.numpy())
    print(f"inpaint loop fp32 vs oracle: rel-L2 {e:.3e}")
    assert e < 1e-4
    # the same loop with sampler='ddpm' (diffusion.py:314-316): graph replays == step-by-step API under the same global seed
    pm = DDPMSampler()
    pm._set_inference_steps(10)
    pm.set_strength(0.4)
    with torch.no_grad():
        torch.manual_seed(77)
        got = inpaint(net15, pm, lat.to(dev), ctx.to(dev), enc.to(dev), mask.to(dev), do_cfg=True, cfg_scale=7.5)
        torch.manual_seed(77)
        x = lat.to(dev)
        for ts in pm.timesteps.to(dev):
            ts = ts.unsqueeze(0)
            x = pm.inpaint_step(x, ts, net15(x.repeat(2, 1, 1, 1), ts, ctx.to(dev)), enc.to(dev), mask.to(dev), cfg_scale=7.5)
        assert torch.isfinite(got).all() and torch.equal(got, x)
    # img2img
    d2 = DDIMSampler()
    d2._set_inference_steps(10)
    noise = torch.randn((B, 4, h, w), generator=g).to(dev)
    ctx2 = torch.randn((2 * B, 77, 768), generator=g).to(dev)
    with torch.no_grad():
        got = img2img(net15, d2, enc.to(dev), noise, ctx2, 0.6)
        d3 = DDIMSampler()
        d3._set_inference_steps(10)
        d3.set_strength(0.6)
        x, _ = d3.forward_process(enc.to(dev), d3.timesteps[0].unsqueeze(0), noise)
        for ts in d3.timesteps.to(dev):
            ts = ts.unsqueeze(0)
            x = d3.reverse_process(x, ts, net15(x.repeat(2, 1, 1, 1), ts, ctx2), cfg_scale=7.5)
        assert torch.equal(got, x)
