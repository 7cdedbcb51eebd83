"""ORACLE (test infrastructure, not product code) — CPU restatement of the reference's
scheduler arithmetic and CFG blend, in numpy fp32 with every rounding step explicit.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.  The product path (stable-diffusion-pytorch_b200/) never does.

Pinning: tests/golden/sampler_*.npz were produced by running the UNMODIFIED reference
(/root/reference/models/scheduler/{ddim,ddpm}.py) in the build container with
tests/golden/make_golden.py; tests/test_oracle_golden.py checks this file against them
BIT-EXACTLY (timesteps, tables and updated latents).

Functions cite the reference lines they restate (paths relative to /root/reference).
"""
from __future__ import annotations

import numpy as np

F = np.float32


def schedule_fp32(noise_step=1000, beta_start=0.00085, beta_end=0.012):
    """betas/alphas/alphas_hat (models/scheduler/ddim.py:9-11, ddpm.py:12-14):
    betas = linspace(sqrt(b0), sqrt(b1), T, fp32)**2, alphas = 1 - betas, alphas_hat = cumprod.

    torch.linspace's CPU kernel evaluates start + i*step in vector-width-dependent pieces, so its
    last bit is not a property of the reference (it varies with the host's SIMD dispatch).  The
    restatement therefore evaluates the documented formula and is compared with the golden table
    to 1e-6 absolute; everything DOWNSTREAM of the table (the update rules below) is restated
    bit-exactly and is tested bit-exactly on the golden table itself."""
    start, end = F(beta_start ** 0.5), F(beta_end ** 0.5)
    n = noise_step
    lin = (np.float64(start) + (np.float64(end) - np.float64(start)) * np.arange(n) / (n - 1)).astype(F)
    betas = (lin * lin).astype(F)
    alphas = (F(1.0) - betas).astype(F)
    a_hat = np.empty(n, dtype=F)
    acc = F(1.0)
    for k in range(n):
        acc = F(acc * alphas[k])
        a_hat[k] = acc
    return betas, alphas, a_hat


def ddim_timesteps(noise_step, inference_steps):
    """ddim.py:28-31: (arange(N) * (T//N) + 1)[::-1], int64."""
    stride = noise_step // inference_steps
    return np.array([k * stride + 1 for k in range(inference_steps)][::-1], dtype=np.int64)


def ddpm_timesteps(noise_step, inference_steps):
    """ddpm.py:29-32: (arange(N) * (T//N))[::-1], int64 (no +1)."""
    stride = noise_step // inference_steps
    return np.array([k * stride for k in range(inference_steps)][::-1], dtype=np.int64)


def prev_timestep(t, noise_step, inference_steps):
    """ddim.py:37-39."""
    return t - noise_step // inference_steps


def strength_slice(timesteps, inference_steps, strength):
    """ddim.py:41-43."""
    return timesteps[inference_steps - int(inference_steps * strength):]


def cfg_blend(pred_2b):
    """models/diffusion.py:233-235 with the scale applied by the caller: returns (u, c)."""
    b = pred_2b.shape[0] // 2
    return pred_2b[:b], pred_2b[b:]


def cfg_combine(u, c, scale):
    """models/diffusion.py:235: u + s*(c - u), three separate fp32 roundings."""
    d = (c.astype(F) - u.astype(F)).astype(F)
    return (u.astype(F) + (F(scale) * d).astype(F)).astype(F)


def ddim_reverse(x_t, t, model_output, alphas, a_hat, noise_step, inference_steps,
                 prediction_type="epsilon", eta=0.0, noise=None):
    """ddim.py:58-87 for one scalar timestep, fp32 tensors, python-double scalars where the
    reference goes through .item()."""
    x_t = x_t.astype(F)
    mo = model_output.astype(F)
    a_t = float(a_hat[t])                          # .item(): fp32 -> python double      (:64)
    s1 = F((1 - a_t) ** 0.5)                       # python scalar * fp32 tensor -> fp32 (:66)
    s2 = F(a_t ** 0.5)
    if prediction_type == "epsilon":
        pred_x0 = (((x_t - (s1 * mo).astype(F)).astype(F)) / s2).astype(F)            # (:66)
        pred_eps = mo                                                                    # (:67)
    elif prediction_type == "v_prediction":
        pred_x0 = ((s2 * x_t).astype(F) - (s1 * mo).astype(F)).astype(F)                # (:69)
        pred_eps = ((s2 * mo).astype(F) + (s1 * x_t).astype(F)).astype(F)               # (:70)
    else:
        raise ValueError(prediction_type)
    prev_t = prev_timestep(t, noise_step, inference_steps)
    alpha_t = alphas[t]                            # NOTE: alphas, not alphas_hat        (:73)
    a_prev = a_hat[prev_t] if prev_t >= 0 else F(1.0)                                   # (:74)
    variance = F(F(F(F(1) - a_prev) / F(F(1) - alpha_t)) * F(F(1) - F(alpha_t / a_prev)))  # (:76)
    std = F(np.sqrt(F(F(eta) * variance)))                                              # (:77)
    cdir = F(np.sqrt(F(F(F(1) - a_prev) - F(std * std))))                               # (:79)
    direction = (cdir * pred_eps).astype(F)
    prev_xt = ((F(np.sqrt(a_prev)) * pred_x0).astype(F) + direction).astype(F)           # (:81)
    if eta > 0:
        prev_xt = (prev_xt + (noise.astype(F) * std).astype(F)).astype(F)                # (:83-85)
    return prev_xt


def ddpm_reverse(x_t, t, model_output, a_hat, noise_step, inference_steps, noise):
    """ddpm.py:62-82; `noise` is the randn draw of :80 supplied by the caller."""
    x_t = x_t.astype(F)
    mo = model_output.astype(F)
    prev_t = prev_timestep(t, noise_step, inference_steps)
    a_t = a_hat[t]
    a_prev = a_hat[prev_t] if prev_t >= 0 else F(1.0)
    cur_a = F(min(max(F(a_t / a_prev), F(0)), F(0.999)))                                 # (:68)
    cur_b = F(F(1) - cur_a)
    inv = F(F(1) / F(np.sqrt(cur_a)))
    ce = F(F(F(1) - cur_a) / F(np.sqrt(F(F(1) - a_t))))
    mu = (inv * (x_t - (ce * mo).astype(F)).astype(F)).astype(F)                         # (:72)
    if t > 0:
        var = F(F(F(F(1) - a_prev) / F(F(1) - a_t)) * cur_b)                             # (:76)
        var = F(max(var, F(1e-20)))
        std = F(np.sqrt(var))
        return (mu + (std * noise.astype(F)).astype(F)).astype(F)
    return (mu + (F(0) * noise.astype(F)).astype(F)).astype(F)                           # (:74,81)


def forward_process(x0, t, noise, a_hat):
    """ddim.py:46-55: per-sample t (n,), sqrt(a)*x0 + sqrt(1-a)*noise."""
    a = a_hat[np.asarray(t)].astype(F)[:, None, None, None]
    return ((np.sqrt(a).astype(F) * x0.astype(F)).astype(F)
            + (np.sqrt((F(1) - a).astype(F)).astype(F) * noise.astype(F)).astype(F)).astype(F)


def x0_from_eps(x, eps, alpha_T=0.0047 ** 0.5, sigma_T=(1 - 0.0047) ** 0.5):
    """models/diffusion.py:111-113 (SwiftBrush one-step)."""
    return (((x.astype(F) - (F(sigma_T) * eps.astype(F)).astype(F)).astype(F)) / F(alpha_T)).astype(F)


def inpaint_step(x_t, t, pred_2b, orig, mask, cfg_scale, alphas, a_hat, noise_step, inference_steps, prediction_type="epsilon",
                 ddpm_noise=None):
    """Loop body of the inpainting path after the UNet call (models/diffusion.py:387-398), fp32, op by op:
    cond, uncond = chunk(2); e = s*(cond - uncond) + cond; noised = forward_process(orig, t, e);
    x = where(~mask, noised, x_t); reverse_process(x, t, e).  ``pred_2b`` may also be the plain (B, ...) prediction with
    cfg_scale None.  mask: bool [h, w], True = repaint.  ``ddpm_noise`` given: the ancestral DDPM update (ddpm.py:62-82) with that
    randn draw instead of DDIM (the reference's inpaint takes sampler='ddpm' too, diffusion.py:314-316)."""
    f = np.float32
    if cfg_scale is None:
        e = pred_2b.astype(f)
    else:
        c, u = np.split(pred_2b.astype(f), 2, axis=0)
        e = (f(cfg_scale) * (c - u)).astype(f) + c
    noised = forward_process(np.broadcast_to(orig, e.shape).astype(f), np.array([t]), e, a_hat)
    x = np.where(mask[None, None].astype(bool), x_t.astype(f), noised)
    if ddpm_noise is not None:
        return ddpm_reverse(x, t, e, a_hat, noise_step, inference_steps, ddpm_noise)
    return ddim_reverse(x, t, e, alphas, a_hat, noise_step, inference_steps, prediction_type=prediction_type)
