"""CPU suite, part 2: host-side bookkeeping of the PRODUCT samplers (bit-exact vs golden) and
loud failure on CPU tensors (no CPU fallback)."""
import os

import numpy as np
import pytest
import torch

from oracle import sampler_oracle as SO
from stable_diffusion_pytorch_b200 import DDIMSampler, DDPMSampler


@pytest.fixture(scope="module")
def sg(golden_dir):
    return np.load(os.path.join(golden_dir, "sampler_golden.npz"))


def test_tables_match_reference(sg):
    d = DDIMSampler()
    # same torch ops as the reference; identical on the machine that wrote the fixture, and within
    # one ulp on hosts whose torch.linspace dispatches to a different SIMD width
    np.testing.assert_allclose(d.alphas_hat.numpy(), sg["alphas_hat"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(d.alphas.numpy(), sg["alphas"], rtol=0, atol=2e-7)
    c = DDIMSampler(use_cosine_schedule=True)
    np.testing.assert_allclose(c.alphas_hat.numpy(), sg["cos_alphas_hat"], rtol=0, atol=1e-6)
    p = DDPMSampler()
    np.testing.assert_allclose(p.alphas_hat.numpy(), sg["alphas_hat"], rtol=0, atol=1e-6)


@pytest.mark.parametrize("n", [1, 7, 10, 20, 50, 250, 999])
def test_timesteps_bit_exact(sg, n):
    d = DDIMSampler()
    assert d.inference_steps == 1000 and d.timesteps[0] == 999 and d.timesteps.dtype == torch.int64
    d._set_inference_steps(n)
    assert d.timesteps.dtype == torch.int64
    assert np.array_equal(d.timesteps.numpy(), sg[f"ddim_ts_{n}"])
    prev = np.array([int(d._get_prev_timestep(t)) for t in d.timesteps], dtype=np.int64)
    assert np.array_equal(prev, sg[f"ddim_prev_{n}"])
    p = DDPMSampler()
    p._set_inference_steps(n)
    assert np.array_equal(p.timesteps.numpy(), sg[f"ddpm_ts_{n}"])


@pytest.mark.parametrize("s", [0.3, 0.6, 0.8, 1.0])
def test_set_strength(sg, s):
    d = DDIMSampler()
    d._set_inference_steps(50)
    d.set_strength(s)
    assert np.array_equal(d.timesteps.numpy(), sg[f"ddim_strength_{s}"])


def test_coefficient_table_matches_oracle_scalars(sg):
    """The [T,8] table the kernel reads holds exactly the scalars the oracle derives per timestep."""
    F = np.float32
    for ptype in ("epsilon", "v_prediction"):
        d = DDIMSampler(prediction_type=ptype)
        d._set_inference_steps(10)
        tab = d.coefficient_table().numpy()
        a_hat = d.alphas_hat.numpy()
        for t in d.timesteps.tolist():
            a = float(a_hat[t])
            assert tab[t, 0] == F((1 - a) ** 0.5) and tab[t, 1] == F(a ** 0.5)
            prev = t - 100
            ap = a_hat[prev] if prev >= 0 else F(1.0)
            assert tab[t, 2] == F(np.sqrt(ap)) and tab[t, 3] == F(np.sqrt(F(F(1) - ap)))
            assert tab[t, 4] == 0.0


def test_ddpm_requires_inference_steps():
    p = DDPMSampler()
    assert not hasattr(p, "inference_steps")          # reference: ddpm.py:11-27 never sets it
    with pytest.raises(AttributeError):
        p.coefficient_table()


def test_cpu_tensors_are_rejected():
    d = DDIMSampler()
    d._set_inference_steps(10)
    x = torch.zeros(1, 4, 8, 8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        d.reverse_process(x, d.timesteps[:1], x)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        d.forward_process(x, d.timesteps[:1], x)
    p = DDPMSampler()
    p._set_inference_steps(10)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        p.reverse_process(x, p.timesteps[:1], x)


def test_from_config(tmp_path):
    import json
    (tmp_path / "scheduler_config.json").write_text(json.dumps(
        {"num_train_timesteps": 1000, "beta_start": 0.00085, "beta_end": 0.012, "prediction_type": "v_prediction"}))
    d = DDIMSampler.from_config(str(tmp_path))
    assert d.prediction_type == "v_prediction" and d.noise_step == 1000
    p = DDPMSampler.from_config(str(tmp_path))       # the reference crashes here (ddpm.py:88); we do not
    assert p.noise_step == 1000
    assert DDIMSampler.step is DDIMSampler.reverse_process
