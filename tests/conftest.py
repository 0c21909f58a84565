import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    import torch
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: (torch.from_numpy(z[k]) if z[k].dtype.kind == "f" else z[k]) for k in z.files}


def manifest(tag, cfg=None):
    """(name, shape) list of the reference state_dict for model ``tag`` (tests/golden/manifest.npz, Cfg-A at 128x256);
    the only shape that depends on the image size is fd_spliter.noise_func = Linear(inner, image_width)."""
    m = load_golden("manifest")
    keys = [str(k) for k in m[tag + ".keys"]]
    shapes = [tuple(int(x) for x in str(s).split(",")) if str(s) else () for s in m[tag + ".shapes"]]
    out = []
    for k, s in zip(keys, shapes):
        if cfg is not None and k.startswith("fd_spliter.noise_func."):
            s = (cfg["image_width"],) + tuple(s[1:])
        if cfg is not None and cfg.get("image_channels", 1) != 1:
            # multi-variable configurations: the stem, the HF-CA query projections and the head scale with C_img
            c = cfg["image_channels"]
            if k == "downs.0.weight":
                s = (s[0], cfg["in_channel"]) + tuple(s[2:])
            elif k.startswith("hf_ca_list.") and k.endswith(".q.weight"):
                s = (s[0], s[1] * c) + tuple(s[2:])
            elif k == "final_conv.block.3.weight":
                s = (cfg["out_channel"],) + tuple(s[1:])
            elif k == "final_conv.block.3.bias":
                s = (cfg["out_channel"],)
        out.append((k, s))
    return out


def module_manifest(arch, cfg):
    """(name, shape) list taken from the package's own UNet built with ``cfg`` -- for configurations other than the canonical
    one of manifest.npz (its key / shape equality with the reference is tested on the canonical configuration)."""
    import wsr
    U = wsr.sub("models.diffusion_models.%s.unet" % arch).UNet
    net = U(in_channel=cfg["in_channel"], out_channel=cfg["out_channel"], norm_groups=cfg["norm_groups"],
            inner_channel=cfg["inner_channel"], channel_mults=cfg["channel_mults"], attn_res=cfg["attn_res"],
            res_blocks=cfg["res_blocks"], dropout=cfg["dropout"], image_height=cfg["image_height"],
            image_width=cfg["image_width"], image_channels=cfg["image_channels"])
    return [(k, tuple(v.shape)) for k, v in net.state_dict().items()]


def rel_l2(a, b):
    a = a.double().flatten()
    b = b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


@pytest.fixture(scope="session")
def golden():
    return load_golden
