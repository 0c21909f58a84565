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


def test_entry_points_on_the_on_disk_store(tmp_path, caplog):
    """``dataroot`` = a store in the reference's layout -> DataHandler (fitted per-month statistics, month
    subset), device loader, train / val (RMSE in physical units from the fitted statistics) / sample by date."""
    from oracle import store
    root = store.write_store(str(tmp_path / "store"), variables=("t2m",), hours=24 * 40)
    cfgmod = wsr.sub("configs.config")
    opt = json.loads(cfgmod.strip_comments(open(os.path.join(ROOT, "configs_examples", "resdiff_eval_b200.json")).read()))
    opt["name"] = "t_store"
    opt["model"]["architecture"] = "sr3"
    opt["model"]["unet"].update(in_channel=2, attn_res=[4])
    opt["model"]["diffusion"].update(image_height=32, image_width=64)
    for ph in ("train", "val"):
        opt["model"]["beta_schedule"][ph]["n_timestep"] = 4
    opt["data"].update(dataroot=root, batch_size=4, val_batch_size=4, height=32, variables=["t2m"], months_subset=[1, 2],
                       transform_groups={"winter": [1, 2]}, transformation="GlobalStandardScaling", num_workers=2, use_shuffle=True,
                       train_min_date="2000-01-01-00", train_max_date="2000-02-05-00", val_min_date="2000-02-05-00", val_max_date="2000-02-06-00")
    opt["train"].update(n_iter=3, print_freq=1, save_checkpoint_freq=1000)
    for k in ("log", "tb_logger", "results", "checkpoint"):
        opt["path"][k] = str(tmp_path / k)
    cfg = tmp_path / "store.json"
    cfg.write_text(json.dumps(opt))
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        train, sample = wsr.sub("train"), wsr.sub("sample")
        import logging
        with caplog.at_level(logging.INFO, logger="base"):
            train.main(["-c", str(cfg), "-p", "train", "-gpu", "0"])
            std_units, phys_units = train.main(["-c", str(cfg), "-p", "val", "-gpu", "0"])
        assert "iter 3" in caplog.text and "physical units" in caplog.text
        # one fitted std for the whole winter group: physical RMSE = std * standardised RMSE
        ratio = float(phys_units["RMSE"]) / float(std_units["RMSE"])
        assert 5.0 < ratio < 15.0
        out = tmp_path / "out"
        sample.main(["-c", str(cfg), "-o", str(out), "-gpu", "0", "-d", "2000-02-05-06"])
        sr, phys = np.load(out / "sr.npy"), np.load(out / "sr_physical.npy")
        assert sr.shape == phys.shape == (1, 1, 32, 64) and np.isfinite(phys).all()
        assert 150.0 < phys.mean() < 400.0 and abs(sr.mean()) < 20.0
    finally:
        os.chdir(cwd)
