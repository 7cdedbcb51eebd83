"""GPU parity: fused CFG + DDIM/DDPM update kernels (through the C-ABI) vs the oracle and the
golden vectors from the unmodified reference — bit-exact."""
import os

import numpy as np
import pytest
import torch

from oracle import sampler_oracle as SO
from stable_diffusion_pytorch_b200 import DDIMSampler, DDPMSampler, x0_from_eps

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sg(golden_dir):
    return np.load(os.path.join(golden_dir, "sampler_golden.npz"))


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def _use_golden_tables(s, sg, cosine=False):
    # pin the float tables to the fixture so bit-exactness does not depend on the host's linspace
    s.alphas = torch.from_numpy(sg["cos_alphas" if cosine else "alphas"]).clone()
    s.alphas_hat = torch.from_numpy(sg["cos_alphas_hat" if cosine else "alphas_hat"]).clone()
    s._coef_cache = {}


@pytest.mark.parametrize("ptype", ["epsilon", "v_prediction"])
@pytest.mark.parametrize("n", [10, 50])
@pytest.mark.parametrize("t_on_device", [True, False])
def test_ddim_golden_bit_exact(sg, dev, ptype, n, t_on_device):
    s = DDIMSampler(prediction_type=ptype)
    _use_golden_tables(s, sg)
    s._set_inference_steps(n)
    x, pred2, eps = (torch.from_numpy(sg[k]).to(dev) for k in ("x", "pred2", "cfg_eps"))
    for t in s.timesteps[[0, 1, n // 2, n - 1]]:
        tt = t.unsqueeze(0).to(dev) if t_on_device else t.unsqueeze(0)
        want = sg[f"ddim_{ptype}_{n}_{int(t)}"]
        got = s.reverse_process(x, tt, eps)                       # plain signature
        assert got.dtype == torch.float32 and np.array_equal(got.cpu().numpy(), want)
        got = s.reverse_process(x, tt, pred2, cfg_scale=7.5)       # fused CFG
        assert np.array_equal(got.cpu().numpy(), want)
        got = s.step(x, int(t), eps)                               # python-int timestep
        assert np.array_equal(got.cpu().numpy(), want)


def test_ddim_cosine_and_eta(sg, dev):
    x, eps = torch.from_numpy(sg["x"]).to(dev), torch.from_numpy(sg["cfg_eps"]).to(dev)
    s = DDIMSampler(use_cosine_schedule=True)
    _use_golden_tables(s, sg, cosine=True)
    s._set_inference_steps(10)
    for t in (901, 1):
        assert np.array_equal(s.reverse_process(x, torch.tensor([t], device=dev), eps).cpu().numpy(), sg[f"ddim_cos_10_{t}"])
    s = DDIMSampler()
    _use_golden_tables(s, sg)
    s._set_inference_steps(10)
    for t in (901, 401, 1):       # reference quirk: eta>0 gives NaN (negative variance), reproduced
        got = s.reverse_process(x, torch.tensor([t], device=dev), eps, eta=0.5).cpu().numpy()
        assert np.array_equal(got, sg[f"ddim_eta0.5_10_{t}"], equal_nan=True)


def test_ddpm_golden_bit_exact(sg, dev):
    p = DDPMSampler()
    _use_golden_tables(p, sg)
    p._set_inference_steps(10)
    x, pred2, eps = (torch.from_numpy(sg[k]).to(dev) for k in ("x", "pred2", "cfg_eps"))
    for t in (900, 500, 0):
        nz = torch.from_numpy(sg[f"ddpm_noise_{t}"]).to(dev)
        got = p.reverse_process(x, torch.tensor([t], device=dev), eps, noise=nz)
        assert np.array_equal(got.cpu().numpy(), sg[f"ddpm_10_{t}"]), t
        got = p.reverse_process(x, torch.tensor([t], device=dev), pred2, cfg_scale=7.5, noise=nz)
        assert np.array_equal(got.cpu().numpy(), sg[f"ddpm_10_{t}"]), t
    # global-RNG draw: same generator state on the same device => same sample as torch.randn
    torch.manual_seed(5)
    a = p.reverse_process(x, torch.tensor([500], device=dev), eps)
    torch.manual_seed(5)
    nz = torch.randn(x.shape, dtype=x.dtype, device=dev)
    b = p.reverse_process(x, torch.tensor([500], device=dev), eps, noise=nz)
    assert torch.equal(a, b)


def test_forward_process_and_onestep(sg, dev):
    d = DDIMSampler()
    _use_golden_tables(d, sg)
    x, nz, eps = (torch.from_numpy(sg[k]).to(dev) for k in ("x", "noise", "cfg_eps"))
    xt, n2 = d.forward_process(x, torch.from_numpy(sg["fwd_t"]).to(dev), nz)
    assert np.array_equal(xt.cpu().numpy(), sg["fwd_xt"]) and n2 is nz
    assert np.array_equal(x0_from_eps(x, eps).cpu().numpy(), sg["onestep_x0"])


@pytest.mark.parametrize("shape", [(1, 4, 64, 64), (3, 4, 24, 40), (64, 4, 64, 64), (1, 1, 1, 3), (0, 4, 8, 8)])
def test_ddim_vs_oracle_shapes(dev, shape):
    """seeded inputs, ragged / empty / full-size shapes (config 3: batch 64 at 64x64)."""
    g = torch.Generator().manual_seed(hash(shape) % 1000)
    x = torch.randn(shape, generator=g)
    pred2 = torch.randn((2 * shape[0],) + shape[1:], generator=g)
    s = DDIMSampler()
    s._set_inference_steps(50)
    for t in (981, 21, 1):
        got = s.reverse_process(x.to(dev), torch.tensor([t], device=dev), pred2.to(dev), cfg_scale=7.5).cpu().numpy()
        u, c = SO.cfg_blend(pred2.numpy())
        want = SO.ddim_reverse(x.numpy(), t, SO.cfg_combine(u, c, 7.5), s.alphas.numpy(), s.alphas_hat.numpy(), 1000, 50)
        assert got.shape == tuple(shape) and np.array_equal(got, want)


@pytest.mark.gpu
@pytest.mark.parametrize("ptype", ["epsilon", "v_prediction"])
def test_inpaint_step_golden_bit_exact(sg, golden_dir, dev, ptype):
    """Fused inpainting step (CFG in the inpaint form + re-noise + mask select + DDIM) == the reference's statements, bit for bit."""
    g = np.load(os.path.join(golden_dir, "inpaint_golden.npz"))
    s = DDIMSampler(prediction_type=ptype)
    _use_golden_tables(s, sg)
    s._set_inference_steps(50)
    s.set_strength(0.8)
    assert np.array_equal(s.timesteps.numpy(), g[f"{ptype}_ts"])
    lat, pred2, enc = (torch.from_numpy(g[k]).to(dev) for k in ("latent", "pred2", "encoded"))
    mask = torch.from_numpy(g["mask"]).to(dev)
    for t in (s.timesteps[0], s.timesteps[17], s.timesteps[-1]):
        tt = t.unsqueeze(0).to(dev)
        got = s.inpaint_step(lat, tt, pred2, enc, mask, cfg_scale=7.5)
        assert np.array_equal(got.cpu().numpy(), g[f"{ptype}_{int(t)}_cfg"])
        got = s.inpaint_step(lat, int(t), pred2[:2], enc.expand(2, -1, -1, -1), mask[0, 0])
        assert np.array_equal(got.cpu().numpy(), g[f"{ptype}_{int(t)}_nocfg"])


@pytest.mark.gpu
def test_inpaint_step_ddpm_golden_bit_exact(sg, golden_dir, dev):
    """sdk_ddpm_inpaint_step (inpaint CFG form + re-noise + mask select + ancestral DDPM update) == the reference's statements with
    sampler='ddpm' (models/diffusion.py:314-316,387-398), bit for bit, including t == 0 where the noise is multiplied by zero."""
    g = np.load(os.path.join(golden_dir, "inpaint_golden.npz"))
    s = DDPMSampler()
    _use_golden_tables(s, sg)
    s._set_inference_steps(50)
    s.set_strength(0.8)
    assert np.array_equal(s.timesteps.numpy(), g["ddpm_ts"])
    lat, pred2, enc = (torch.from_numpy(g[k]).to(dev) for k in ("latent", "pred2", "encoded"))
    mask = torch.from_numpy(g["mask"]).to(dev)
    for t in (s.timesteps[0], s.timesteps[17], s.timesteps[-1]):
        nz = torch.from_numpy(g[f"ddpm_{int(t)}_noise"]).to(dev)
        got = s.inpaint_step(lat, t.unsqueeze(0).to(dev), pred2, enc, mask, cfg_scale=7.5, noise=nz)
        assert np.array_equal(got.cpu().numpy(), g[f"ddpm_{int(t)}_cfg"]), int(t)
    # without `noise` the draw comes from torch's global generator on the tensor's device, like ddpm.py:80
    torch.manual_seed(3)
    a = s.inpaint_step(lat, int(s.timesteps[0]), pred2, enc, mask, cfg_scale=7.5)
    torch.manual_seed(3)
    nz = torch.randn(lat.shape, dtype=lat.dtype, device=dev)
    b = s.inpaint_step(lat, int(s.timesteps[0]), pred2, enc, mask, cfg_scale=7.5, noise=nz)
    assert torch.equal(a, b)


def test_out_of_range_timestep_poisons(dev):
    s = DDIMSampler()
    s._set_inference_steps(10)
    x = torch.ones(1, 4, 8, 8, device=dev)
    y = s.reverse_process(x, torch.tensor([1000], device=dev), x)     # reference: IndexError
    assert torch.isnan(y).all()


def test_graph_capture_no_sync(dev):
    """The step is capturable in a CUDA graph with the timestep read from device memory."""
    s = DDIMSampler()
    s._set_inference_steps(10)
    x = torch.randn(2, 4, 32, 32, device=dev)
    pred2 = torch.randn(4, 4, 32, 32, device=dev)
    t = torch.tensor([901], device=dev)
    s.reverse_process(x, t, pred2, cfg_scale=7.5)                      # warm the coefficient table
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        y = s.reverse_process(x, t, pred2, cfg_scale=7.5)
    for tv in (901, 401, 1):
        t.fill_(tv)
        g.replay()
        torch.cuda.synchronize()
        want = s.reverse_process(x, torch.tensor([tv], device=dev), pred2, cfg_scale=7.5)
        assert torch.equal(y, want)
