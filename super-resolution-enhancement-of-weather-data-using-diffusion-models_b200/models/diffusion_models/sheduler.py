"""Noise schedules (host-side float64 tables) -- same names/arguments as the reference's
models/diffusion_models/sheduler.py:25-68 (the file name keeps the reference's spelling)."""
import math

import numpy as np


def _warmup(linear_start, linear_end, n_timestep, frac):
    out = np.full(n_timestep, linear_end, dtype=np.float64)
    n_warm = int(n_timestep * frac)
    out[:n_warm] = np.linspace(linear_start, linear_end, n_warm, dtype=np.float64)
    return out


def _cosine(n_timestep, cosine_s):
    grid = np.arange(n_timestep + 1, dtype=np.float64) / n_timestep + cosine_s
    abar = np.cos(grid / (1 + cosine_s) * math.pi / 2) ** 2
    abar = abar / abar[0]
    return np.minimum(1 - abar[1:] / abar[:-1], 0.999)


_SCHEDULES = {
    "quad": lambda s, e, n, c: np.linspace(s ** 0.5, e ** 0.5, n, dtype=np.float64) ** 2,
    "linear": lambda s, e, n, c: np.linspace(s, e, n, dtype=np.float64),
    "warmup10": lambda s, e, n, c: _warmup(s, e, n, 0.1),
    "warmup50": lambda s, e, n, c: _warmup(s, e, n, 0.5),
    "const": lambda s, e, n, c: np.full(n, e, dtype=np.float64),
    "jsd": lambda s, e, n, c: 1.0 / np.linspace(n, 1, n, dtype=np.float64),
    "cosine": lambda s, e, n, c: _cosine(n, c),
}


def make_beta_schedule(schedule, n_timestep, linear_start=1e-4, linear_end=2e-2, cosine_s=8e-3):
    try:
        fn = _SCHEDULES[schedule]
    except KeyError:
        raise NotImplementedError(schedule)
    return fn(linear_start, linear_end, int(n_timestep), cosine_s)
