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


def test_resume_continues_iterations_epochs_and_adam_state(tmp_path, caplog):
    """a resumed run starts at the loaded iteration / epoch (reference train.py:55-64 + model.get_loaded_iter/epoch),
    stops at n_iter, names its checkpoints by the real counters, keeps the Adam moments and step count, and uses the one-launch
    flat Adam (attach_flat) in the train.py path."""
    cfgmod = wsr.sub("configs.config")
    cfg = _small_config(tmp_path, "sr3", 2)
    opt = json.loads(open(cfg).read())
    opt["train"].update(n_iter=2, save_checkpoint_freq=2, val_freq=2, full_val_freq=1000)
    open(cfg, "w").write(json.dumps(opt))
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        train = wsr.sub("train")
        nat = wsr.pkg.native
        import logging
        with caplog.at_level(logging.INFO, logger="base"):
            assert train.main(["-c", cfg, "-p", "train", "-gpu", "0"]) == (2, 1)
        assert "validation (standardised units)" in caplog.text          # val_freq validation inside the training loop
        ck = str(tmp_path / "checkpoint")
        assert sorted(os.listdir(ck)) == ["I2_E1_gen.pth", "I2_E1_opt.pth"]
        o1 = torch.load(os.path.join(ck, "I2_E1_opt.pth"), map_location="cpu")
        st1 = o1["optimizer"]["state"]
        assert all(float(s["step"]) == 2.0 for s in st1.values())
        assert any(float(s["exp_avg"].abs().sum()) > 0 for s in st1.values())
        # resume: two more iterations
        opt["train"].update(n_iter=4)
        opt["path"]["resume_state"] = os.path.join(ck, "I2_E1")
        open(cfg, "w").write(json.dumps(opt))
        l0 = nat.launches
        assert train.main(["-c", cfg, "-p", "train", "-gpu", "0"]) == (4, 2)
        assert "I4_E2_gen.pth" in os.listdir(ck) and "I2_E1_gen.pth" in os.listdir(ck)
        o2 = torch.load(os.path.join(ck, "I4_E2_opt.pth"), map_location="cpu")
        st2 = o2["optimizer"]["state"]
        assert all(float(s["step"]) == 4.0 for s in st2.values())
        # the loaded moments were carried into the flat buffers (not reset to zero): after two more steps with b2 = 0.999 the second
        # moment still holds (0.999^2 of) the loaded one
        k = max(st1, key=lambda i: float(st1[i]["exp_avg_sq"].sum()))
        assert float(st2[k]["exp_avg_sq"].sum()) > 0.99 * float(st1[k]["exp_avg_sq"].sum())
        del l0, cfgmod
    finally:
        os.chdir(cwd)
