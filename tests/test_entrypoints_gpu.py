"""End-to-end GPU test of the config-driven entry points (reference train.py / sample.py flags): a small configuration per
architecture is trained for two iterations (optimize_parameters on the hand-written backward pass), validated (generate_sr +
device-side metric accumulators) and sampled to a file."""
import json
import os

import numpy as np
import pytest
import torch

import wsr
from conftest import ROOT

pytestmark = pytest.mark.gpu


def _small_config(tmp_path, arch, in_channel):
    cfgmod = wsr.sub("configs.config")
    src = open(os.path.join(ROOT, "configs_examples", "resdiff_eval_b200.json")).read()
    opt = json.loads(cfgmod.strip_comments(src))
    opt["name"] = "t_" + arch
    opt["model"]["architecture"] = arch
    opt["model"]["unet"].update(in_channel=in_channel, attn_res=[4])
    opt["model"]["diffusion"].update(image_height=32, image_width=64)
    for ph in ("train", "val"):
        opt["model"]["beta_schedule"][ph]["n_timestep"] = 4
    opt["data"].update(batch_size=2, val_batch_size=2, height=32)
    opt["train"].update(n_iter=2, print_freq=1, save_checkpoint_freq=1000)
    for k in ("log", "tb_logger", "results", "checkpoint"):
        opt["path"][k] = str(tmp_path / k)
    p = tmp_path / (arch + ".json")
    p.write_text(json.dumps(opt))
    return str(p)


@pytest.mark.parametrize("arch,in_channel", [("resdiff", 5), ("phydiff", 5), ("sr3", 2)])
def test_train_val_sample_entry_points(tmp_path, arch, in_channel, caplog):
    cfg = _small_config(tmp_path, arch, in_channel)
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        train, sample = wsr.sub("train"), wsr.sub("sample")
        import logging
        with caplog.at_level(logging.INFO, logger="base"):
            train.main(["-c", cfg, "-p", "train", "-gpu", "0"])
            train.main(["-c", cfg, "-p", "val", "-gpu", "0"])
        text = caplog.text
        assert "l_pix" in text and "RMSE" in text and "MAE" in text
        out = tmp_path / "out"
        sample.main(["-c", cfg, "-o", str(out), "-gpu", "0"])
        sr = np.load(out / "sr.npy")
        assert sr.shape == (2, 1, 32, 64) and np.isfinite(sr).all()
    finally:
        os.chdir(cwd)
