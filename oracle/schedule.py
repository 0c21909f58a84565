"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

float64 numpy restatement of the noise schedules and the derived DDPM coefficient tables.

Follows: models/diffusion_models/sheduler.py:25-68 (make_beta_schedule) and
models/diffusion_models/diffusion.py:49-96 (GaussianDiffusion.set_new_noise_schedule).
"""
import math

import numpy as np


def betas_for(schedule, n_timestep, linear_start=1e-4, linear_end=2e-2, cosine_s=8e-3):
    """sheduler.py:25-68.  Returns a float64 numpy vector of length n_timestep."""
    T = int(n_timestep)
    if schedule == "linear":
        return np.linspace(linear_start, linear_end, T, dtype=np.float64)
    if schedule == "quad":
        return np.linspace(math.sqrt(linear_start), math.sqrt(linear_end), T, dtype=np.float64) ** 2
    if schedule in ("warmup10", "warmup50"):
        frac = 0.1 if schedule == "warmup10" else 0.5
        out = np.full(T, linear_end, dtype=np.float64)
        n_warm = int(T * frac)
        out[:n_warm] = np.linspace(linear_start, linear_end, n_warm, dtype=np.float64)
        return out
    if schedule == "const":
        return np.full(T, linear_end, dtype=np.float64)
    if schedule == "jsd":
        return 1.0 / np.linspace(T, 1, T, dtype=np.float64)
    if schedule == "cosine":
        steps = np.arange(T + 1, dtype=np.float64) / T + cosine_s
        abar = np.cos(steps / (1 + cosine_s) * math.pi / 2) ** 2
        abar = abar / abar[0]
        return np.minimum(1 - abar[1:] / abar[:-1], 0.999)
    raise NotImplementedError(schedule)


BUFFER_NAMES = (
    "betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_alphas_cumprod",
    "sqrt_one_minus_alphas_cumprod", "log_one_minus_alphas_cumprod", "sqrt_recip_alphas_cumprod",
    "sqrt_recipm1_alphas_cumprod", "posterior_variance", "posterior_log_variance_clipped",
    "posterior_mean_coef1", "posterior_mean_coef2",
)


def ddpm_tables(schedule_opt):
    """diffusion.py:49-96.  Returns (dict of 12 float32 arrays [T], sqrt_alphas_cumprod_prev float64 [T+1])."""
    b = betas_for(schedule_opt["schedule"], schedule_opt["n_timestep"],
                  schedule_opt["linear_start"], schedule_opt["linear_end"])
    a = 1.0 - b
    abar = np.cumprod(a)
    abar_prev = np.concatenate([[1.0], abar[:-1]])
    sqrt_abar_prev_ext = np.sqrt(np.concatenate([[1.0], abar]))   # diffusion.py:68-69 (host numpy, T+1)
    post_var = b * (1.0 - abar_prev) / (1.0 - abar)
    t64 = {
        "betas": b,
        "alphas_cumprod": abar,
        "alphas_cumprod_prev": abar_prev,
        "sqrt_alphas_cumprod": np.sqrt(abar),
        "sqrt_one_minus_alphas_cumprod": np.sqrt(1.0 - abar),
        "log_one_minus_alphas_cumprod": np.log(1.0 - abar),
        "sqrt_recip_alphas_cumprod": np.sqrt(1.0 / abar),
        "sqrt_recipm1_alphas_cumprod": np.sqrt(1.0 / abar - 1.0),
        "posterior_variance": post_var,
        "posterior_log_variance_clipped": np.log(np.maximum(post_var, 1e-20)),
        "posterior_mean_coef1": b * np.sqrt(abar_prev) / (1.0 - abar),
        "posterior_mean_coef2": (1.0 - abar_prev) * np.sqrt(a) / (1.0 - abar),
    }
    return {k: v.astype(np.float32) for k, v in t64.items()}, sqrt_abar_prev_ext
