"""CPU suite, part 1: pin the ORACLE (oracle/) against golden vectors produced by the unmodified
reference (tests/golden/make_golden.py).  No GPU, no /root/reference needed."""
import os

import numpy as np
import pytest
import torch

from oracle import sampler_oracle as SO
from oracle import unet_oracle as UO


@pytest.fixture(scope="module")
def sg(golden_dir):
    return np.load(os.path.join(golden_dir, "sampler_golden.npz"))


def test_schedule_table_close(sg):
    b, a, ah = SO.schedule_fp32()
    np.testing.assert_allclose(b, sg["betas"], rtol=0, atol=1e-8)
    np.testing.assert_allclose(ah, sg["alphas_hat"], rtol=0, atol=1e-6)
    # SURVEY §8(a) S3 probe values
    assert float(sg["alphas_hat"][981]) == 0.005775495897978544
    assert float(sg["alphas_hat"][1]) == 0.9982960224151611
    assert float(sg["alphas_hat"][999]) == 0.00466009508818388


@pytest.mark.parametrize("n", [1, 7, 10, 20, 50, 250, 999])
def test_timestep_grids_bit_exact(sg, n):
    ts = SO.ddim_timesteps(1000, n)
    assert ts.dtype == np.int64 and np.array_equal(ts, sg[f"ddim_ts_{n}"])
    assert np.array_equal(SO.ddpm_timesteps(1000, n), sg[f"ddpm_ts_{n}"])
    prev = np.array([SO.prev_timestep(int(t), 1000, n) for t in ts], dtype=np.int64)
    assert np.array_equal(prev, sg[f"ddim_prev_{n}"])


def test_known_grids(sg):
    assert sg["ddim_ts_10"].tolist() == [901, 801, 701, 601, 501, 401, 301, 201, 101, 1]
    assert sg["ddim_prev_10"].tolist() == [801, 701, 601, 501, 401, 301, 201, 101, 1, -99]
    assert sg["ddim_ts_50"][0] == 981 and sg["ddim_ts_50"][-1] == 1
    assert sg["ddpm_ts_50"][0] == 980 and sg["ddpm_ts_50"][-1] == 0


@pytest.mark.parametrize("s", [0.3, 0.6, 0.8, 1.0])
def test_strength_slice(sg, s):
    got = SO.strength_slice(SO.ddim_timesteps(1000, 50), 50, s)
    assert np.array_equal(got, sg[f"ddim_strength_{s}"])
    if s == 0.6:
        assert len(got) == 30 and got[0] == 581


def test_cfg_combine_bit_exact(sg):
    u, c = SO.cfg_blend(sg["pred2"])
    assert np.array_equal(SO.cfg_combine(u, c, 7.5), sg["cfg_eps"])


@pytest.mark.parametrize("ptype", ["epsilon", "v_prediction"])
@pytest.mark.parametrize("n", [10, 50])
def test_ddim_reverse_bit_exact(sg, ptype, n):
    ts = SO.ddim_timesteps(1000, n)
    for t in ts[[0, 1, n // 2, n - 1]]:
        y = SO.ddim_reverse(sg["x"], int(t), sg["cfg_eps"], sg["alphas"], sg["alphas_hat"], 1000, n, ptype)
        assert np.array_equal(y, sg[f"ddim_{ptype}_{n}_{int(t)}"]), (ptype, n, int(t))


def test_ddim_eta_bit_exact(sg):
    for t in (901, 401, 1):
        y = SO.ddim_reverse(sg["x"], t, sg["cfg_eps"], sg["alphas"], sg["alphas_hat"], 1000, 10,
                            eta=0.5, noise=sg[f"ddim_eta_noise_{t}"])
        # NB: the reference's eta>0 variance uses alphas[t] instead of alphas_hat[t] (ddim.py:73) and is
        # negative for the SD schedule -> sqrt -> NaN latents.  The quirk is reproduced, not fixed.
        assert np.array_equal(y, sg[f"ddim_eta0.5_10_{t}"], equal_nan=True), t


def test_ddim_cosine_bit_exact(sg):
    for t in (901, 1):
        y = SO.ddim_reverse(sg["x"], t, sg["cfg_eps"], sg["cos_alphas"], sg["cos_alphas_hat"], 1000, 10)
        assert np.array_equal(y, sg[f"ddim_cos_10_{t}"]), t


def test_ddpm_reverse_bit_exact(sg):
    for t in (900, 500, 0):
        y = SO.ddpm_reverse(sg["x"], t, sg["cfg_eps"], sg["alphas_hat"], 1000, 10, sg[f"ddpm_noise_{t}"])
        assert np.array_equal(y, sg[f"ddpm_10_{t}"]), t


def test_forward_process_and_onestep_bit_exact(sg):
    assert np.array_equal(SO.forward_process(sg["x"], sg["fwd_t"], sg["noise"], sg["alphas_hat"]), sg["fwd_xt"])
    assert np.array_equal(SO.x0_from_eps(sg["x"], sg["cfg_eps"]), sg["onestep_x0"])


# ---------------------------------------------------------------------------------------------
def rel_l2(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


@pytest.mark.parametrize("ptype", ["epsilon", "v_prediction"])
def test_inpaint_step_bit_exact(sg, golden_dir, ptype):
    """Oracle restatement of the inpainting loop body vs the reference's own statements (tests/golden/make_golden_inpaint.py)."""
    g = np.load(os.path.join(golden_dir, "inpaint_golden.npz"))
    ts = g[f"{ptype}_ts"]
    assert np.array_equal(ts, SO.strength_slice(SO.ddim_timesteps(1000, 50), 50, 0.8))
    mask = g["mask"][0, 0]
    for t in (int(ts[0]), int(ts[17]), int(ts[-1])):
        y = SO.inpaint_step(g["latent"], t, g["pred2"], g["encoded"], mask, 7.5, sg["alphas"], sg["alphas_hat"], 1000, 50, ptype)
        assert np.array_equal(y, g[f"{ptype}_{t}_cfg"]), (ptype, t)
        y = SO.inpaint_step(g["latent"], t, g["pred2"][:2], g["encoded"], mask, None, sg["alphas"], sg["alphas_hat"], 1000, 50, ptype)
        assert np.array_equal(y, g[f"{ptype}_{t}_nocfg"]), (ptype, t)


def test_inpaint_step_ddpm_bit_exact(sg, golden_dir):
    """The inpainting loop body with sampler='ddpm' (models/diffusion.py:314-316): oracle vs the reference's statements, with the
    reference's own global-generator noise draw recorded in the fixture."""
    g = np.load(os.path.join(golden_dir, "inpaint_golden.npz"))
    ts = g["ddpm_ts"]
    assert np.array_equal(ts, SO.strength_slice(SO.ddpm_timesteps(1000, 50), 50, 0.8))
    mask = g["mask"][0, 0]
    for t in (int(ts[0]), int(ts[17]), int(ts[-1])):
        y = SO.inpaint_step(g["latent"], t, g["pred2"], g["encoded"], mask, 7.5, sg["alphas"], sg["alphas_hat"], 1000, 50,
                            ddpm_noise=g[f"ddpm_{t}_noise"])
        assert np.array_equal(y, g[f"ddpm_{t}_cfg"]), t


def test_param_inventory():
    n15 = sum(int(np.prod(s)) for _, s in UO.param_spec(**UO.SD15))
    n21 = sum(int(np.prod(s)) for _, s in UO.param_spec(**UO.SD21))
    assert n15 == 859_520_964 and n21 == 865_910_724          # SURVEY §8(c) probe
    assert len(UO.param_spec(**UO.SD15)) == 686


@pytest.fixture(scope="module")
def sd15():
    return UO.make_state_dict(0, **UO.SD15)


def test_unet_oracle_sd15_matches_reference(golden_dir, sd15):
    g = np.load(os.path.join(golden_dir, "unet_sd15_golden.npz"))
    assert int(g["n_params"]) == 859_520_964
    T = torch.from_numpy
    with torch.no_grad():
        lat, ctx = T(g["lat16"]), T(g["ctx"])
        y = UO.unet_forward(sd15, lat.repeat(2, 1, 1, 1), torch.tensor([981]), ctx, **UO.SD15)
        assert rel_l2(y.numpy(), g["out16_t981"]) < 2e-5            # fp32, different summation order
        y = UO.unet_forward(sd15, T(g["lat8"]).repeat(2, 1, 1, 1), torch.tensor([500, 20]), T(g["ctx8"]), **UO.SD15)
        assert rel_l2(y.numpy(), g["out8_t500_20"]) < 2e-5
        y = UO.unet_forward(sd15, T(g["lat8"]).repeat(2, 1, 1, 1), torch.tensor([999]), T(g["ctx8"])[:1], **UO.SD15)
        assert rel_l2(y.numpy(), g["out8_ctx1_t999"]) < 2e-5
        y = UO.unet_forward(sd15, T(g["lat8x16"]), torch.tensor([301]), T(g["ctx8"])[1:], **UO.SD15)
        assert rel_l2(y.numpy(), g["out8x16_t301"]) < 2e-5


def test_unet_oracle_sd21_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "unet_sd21_golden.npz"))
    assert int(g["n_params"]) == 865_910_724
    sd = UO.make_state_dict(1, **UO.SD21)
    T = torch.from_numpy
    with torch.no_grad():
        y = UO.unet_forward(sd, T(g["lat16"]).repeat(2, 1, 1, 1), torch.tensor([961]), T(g["ctx"]), **UO.SD21)
        assert rel_l2(y.numpy(), g["out16_t961"]) < 2e-5
        assert g["onestep_t"].tolist() == [999]
        pn = UO.unet_forward(sd, T(g["onestep_lat"]), torch.tensor([999]), T(g["ctx"])[:1], **UO.SD21)
        assert rel_l2(pn.numpy(), g["onestep_pred"]) < 2e-5
        assert rel_l2(SO.x0_from_eps(g["onestep_lat"], pn.numpy()), g["onestep_x0"]) < 2e-5


def test_vae_oracle_matches_reference_golden(golden_dir):
    """oracle/vae_oracle.decode restates VAE.decode (models/vae/vae.py:270-274): pinned on images produced by the unmodified
    reference (tests/golden/make_golden_vae.py), and the product's parameter contract equals the oracle's (= the reference's)."""
    import torch
    from oracle import vae_oracle as VO
    from stable_diffusion_pytorch_b200.arch import vae_param_spec
    assert vae_param_spec() == VO.param_spec() and len(VO.param_spec()) == 248
    g = np.load(os.path.join(golden_dir, "vae_golden.npz"))
    sd = VO.make_state_dict(3)
    with torch.no_grad():
        for name in ("z8", "z8x16"):
            y = VO.decode(sd, torch.from_numpy(g[name])).numpy()
            assert np.linalg.norm(y - g[f"img_{name}"]) / np.linalg.norm(g[f"img_{name}"]) < 2e-5


def test_clip_oracle_matches_reference_golden(golden_dir):
    """oracle/clip_oracle.text_forward restates CLIPTextModel.forward (openclip.py:133-137) and TextEncoder.forward (clip.py:28-34):
    pinned on outputs of the unmodified reference classes; the product's parameter contract equals the oracle's."""
    import torch
    from oracle import clip_oracle as CO
    from stable_diffusion_pytorch_b200.clip import text_param_spec
    g = np.load(os.path.join(golden_dir, "clip_golden.npz"))
    for tag, cfg, seed in (("openclip", CO.SMALL_OPENCLIP, 5), ("clip", CO.SMALL_CLIP, 6)):
        assert text_param_spec(cfg["kind"], cfg["vocab"], cfg["hidden"], cfg["heads"], cfg["layers"], cfg["inter"], cfg["max_len"]) == CO.param_spec(**cfg)
        sd = CO.make_state_dict(seed, **cfg)
        with torch.no_grad():
            y = CO.text_forward(sd, torch.from_numpy(g[f"{tag}_ids"]), **cfg).numpy()
        assert np.linalg.norm(y - g[f"{tag}_out"]) / np.linalg.norm(g[f"{tag}_out"]) < 1e-5
