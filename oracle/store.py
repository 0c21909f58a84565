"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Writes a small synthetic store in the reference's on-disk layout (data/conversions/netcdf_to_npy.py:134-246):

    <root>/{lr,hr}/<variable>/meta/metadata.json
    <root>/{lr,hr}/<variable>/samples/<year>/<YYYY-MM-DD-HH>.npy      (float32, (H, W), hourly)

The same function feeds ``oracle/make_golden.py store`` (which runs the REAL reference DataHandler on it) and the tests
(which run the package's DataHandler on an identical copy), so the fixtures only hold the reference's OUTPUTS."""
import json
import os
from datetime import datetime, timedelta

import numpy as np

VARIABLES = ("t2m", "z500")
START = "2000-01-01-00"
HOURS = 24 * 65                       # 2000-01-01-00 .. 2000-03-06-00
LR_SHAPE = (8, 16)
SCALE = 4
FMT = "%Y-%m-%d-%H"
SPEC = dict(months_subset=[1, 3], groups=[[1], [3]], train=("2000-01-01-00", "2000-03-04-00"), val=("2000-03-04-00", "2000-03-06-00"),
            val_batch_size=3, train_batch_size=4)


def field(var_index, hour, rng_seed=7):
    """(H, W) float32 high-resolution field of variable ``var_index`` at ``hour`` (deterministic)."""
    h, w = LR_SHAPE[0] * SCALE, LR_SHAPE[1] * SCALE
    rng = np.random.default_rng([rng_seed, var_index, hour])
    lat = np.linspace(-1.0, 1.0, h)[:, None]
    lon = np.linspace(0.0, 2 * np.pi, w, endpoint=False)[None, :]
    base = 275.0 + 40.0 * var_index - 25.0 * lat ** 2 + 6.0 * np.sin(lon + 2 * np.pi * hour / 24.0) + 0.01 * hour
    return (base + rng.normal(0.0, 2.5, size=(h, w))).astype(np.float32)


def write_store(root, variables=VARIABLES, hours=HOURS):
    t0 = datetime.strptime(START, FMT)
    h, w = LR_SHAPE[0] * SCALE, LR_SHAPE[1] * SCALE
    for vi, var in enumerate(variables):
        for kind, (gh, gw) in (("lr", LR_SHAPE), ("hr", (h, w))):
            base = os.path.join(root, kind, var)
            os.makedirs(os.path.join(base, "meta"), exist_ok=True)
            meta = {"name": var, "time_variate": True, "dims": ["lat", "lon"], "shape": [gh, gw],
                    "coords": [{"name": "lat", "values": np.linspace(-87.1875, 87.1875, gh).tolist(), "dims": ["lat"]},
                               {"name": "lon", "values": np.linspace(0.0, 360.0, gw, endpoint=False).tolist(), "dims": ["lon"]}],
                    "attrs": {"units": "K"}}
            with open(os.path.join(base, "meta", "metadata.json"), "w") as fh:
                json.dump(meta, fh)
        for hour in range(hours):
            t = t0 + timedelta(hours=hour)
            hr = field(vi, hour)
            lr = hr.reshape(LR_SHAPE[0], SCALE, LR_SHAPE[1], SCALE).mean(axis=(1, 3)).astype(np.float32)
            for kind, arr in (("lr", lr), ("hr", hr)):
                d = os.path.join(root, kind, var, "samples", str(t.year))
                os.makedirs(d, exist_ok=True)
                np.save(os.path.join(d, t.strftime(FMT) + ".npy"), arr)
    return root
