"""SURVEY 8f N4 on the device: ``wsr_image_compare_loss`` (value and gradient from one launch) against the oracle's FFT + Haar
restatement, one SimpleCNN pre-training step against the REAL reference's gradients (tests/golden/simple_cnn_pretrain.npz), a short
optimisation run, and the ``pretrain.py`` entry point on a store in the reference's on-disk layout."""
import json
import os

import pytest
import torch

import wsr
from conftest import ROOT, load_golden, rel_l2
from oracle import edge
from oracle.cases import CASES
from oracle.weights import fill_module, seeded_randn

pytestmark = pytest.mark.gpu
loss_mod = wsr.sub("models.simple_cnn.loss")


@pytest.mark.parametrize("shape", [(3, 1, 32, 64), (2, 3, 128, 256), (1, 1, 16, 16)])
def test_loss_kernel_value_and_gradient(shape):
    x = seeded_randn("pl.x", shape, 5).requires_grad_(True)
    y = seeded_randn("pl.y", shape, 6)
    ref = edge.image_compare_loss(x, y)
    ref.backward()
    xd = x.detach().cuda().requires_grad_(True)
    got = loss_mod.image_compare_loss(xd, y.cuda())
    (got * 3.0).backward()                                 # upstream gradient != 1
    assert float(got) == pytest.approx(float(ref), rel=2e-6)
    assert rel_l2(xd.grad.cpu() / 3.0, x.grad) < 2e-6
    assert float(loss_mod.fft_mse_loss(xd, y.cuda())) == pytest.approx(float(edge.fft_mse_loss(x, y)), rel=2e-6)
    assert float(loss_mod.dwt_mse_loss(xd, y.cuda())) == pytest.approx(float(edge.dwt_mse_loss(x, y)), rel=2e-6)


def test_loss_refuses_what_it_cannot_do():
    with pytest.raises(wsr.pkg.native.WsrError):
        loss_mod.image_compare_loss(torch.zeros(1, 1, 16, 16), torch.zeros(1, 1, 16, 16))          # host tensors
    with pytest.raises(wsr.pkg.native.WsrError, match="multiple of 16"):
        loss_mod.image_compare_loss(torch.zeros(1, 1, 24, 16).cuda(), torch.zeros(1, 1, 24, 16).cuda())


def test_simple_cnn_pretrain_step_vs_reference():
    g, spec = load_golden("simple_cnn_pretrain"), CASES["simple_cnn_pretrain"]
    S = wsr.sub("models.simple_cnn.Simple_CNN").SimpleCNN
    net = fill_module(S(scale_factor=4, channels=1), spec["seed"]).cuda().train()
    pred = net(g["lr"].cuda())
    assert pred.requires_grad and rel_l2(pred.detach().cpu(), g["pred"]) < 1e-5
    loss = loss_mod.image_compare_loss(pred, g["hr"].cuda())
    loss.backward()
    assert float(loss) == pytest.approx(float(g["loss"]), rel=1e-5)
    for n, p in net.named_parameters():
        err = rel_l2(p.grad.cpu(), g["grad." + n])
        print("[parity] simple_cnn pretrain grad %-14s rel-L2 %.2e" % (n, err))
        assert err < 1e-4, n
    # a second backward accumulates (torch semantics), zero_grad clears
    loss2 = loss_mod.image_compare_loss(net(g["lr"].cuda()), g["hr"].cuda())
    loss2.backward()
    assert rel_l2(net.conv2.weight.grad.cpu(), 2 * g["grad.conv2.weight"]) < 1e-4


def test_pretraining_reduces_the_loss():
    torch.manual_seed(0)
    S = wsr.sub("models.simple_cnn.Simple_CNN").SimpleCNN
    net = S(4, 1).cuda().train()
    opt = wsr.sub("autograd_glue").FusedAdam(net.parameters(), lr=1e-3)
    lr = seeded_randn("pt.lr", (8, 1, 16, 32), 9).cuda()
    hr = wsr.sub("data.dataset_builder").bicubic_sr(lr, 4) + 0.3 * seeded_randn("pt.n", (8, 1, 64, 128), 9).cuda()
    losses = []
    for _ in range(30):
        loss = loss_mod.image_compare_loss(net(lr), hr)
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(float(loss))
    assert losses[-1] < 0.9 * losses[0], losses[::6]


def test_pretrain_entry_point_on_the_store(tmp_path, caplog):
    from oracle import store
    root = store.write_store(str(tmp_path / "store"), variables=("t2m",), hours=24 * 36)
    opt = {"name": "pretrain_simplesr_test", "gpu_id": [0], "phase": "train",
           "path": {"log": "logs", "results": "results", "checkpoint": "checkpoint", "resume_state": None, "validation_results_path": "validation"},
           "data": {"name": "WeatherBench", "dataroot": root, "batch_size": 16, "val_batch_size": 8, "num_workers": 2, "use_shuffle": True,
                    "train_min_date": "2000-01-01-00", "train_max_date": "2000-02-04-00", "transformation": "GlobalStandardScaling",
                    "months_subset": [1, 2], "transform_groups": {"winter": [1, 2]}, "val_min_date": "2000-02-04-00",
                    "val_max_date": "2000-02-05-00", "variables": ["t2m"], "height": 32},
           "model": {"name": "SimpleSR", "in_channel": 1, "out_channel": 1},
           "train": {"epoch": 2, "optimizer": {"type": "adam", "amsgrad": False, "lr": 1e-3}, "save_checkpoint_freq_epoch": 1},
           "save_images": 0}
    cfg = tmp_path / "pretrain.json"
    cfg.write_text(json.dumps(opt))
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        pre = wsr.sub("pretrain")
        history = pre.main(["-c", str(cfg), "-p", "train", "-gpu", "0"])
        assert len(history) == 2 and history[1][0] < history[0][0]                  # training loss falls from epoch 1 to 2
        assert 0.0 < history[1][1]["RMSE"] < 20.0                                   # Kelvin-like units via the fitted std
        ckpts = [f for _, _, fs in os.walk(tmp_path) for f in fs if f.startswith("pretrain_") and f.endswith("_gen.pth")]
        assert len(ckpts) == 2
        opt["model"] = {"name": "RRDBNet", "in_channel": 1, "out_channel": 1, "hidden_size": 64, "num_block": 1}
        cfg.write_text(json.dumps(opt))
        res = pre.main(["-c", str(cfg), "-p", "val", "-gpu", "0"])
        assert set(res) == {"MSE", "RMSE", "MAE", "MR"}
        opt["train"]["epoch"] = 1
        cfg.write_text(json.dumps(opt))
        hist = pre.main(["-c", str(cfg), "-p", "train", "-gpu", "0"])              # encoder pre-training with F.l1_loss
        assert len(hist) == 1 and hist[0][0] > 0.0
        opt["train"]["optimizer"]["amsgrad"] = True
        cfg.write_text(json.dumps(opt))
        with pytest.raises(NotImplementedError):
            pre.main(["-c", str(cfg), "-p", "train", "-gpu", "0"])
    finally:
        os.chdir(cwd)



@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_rrdb_pretrain_step_vs_reference(precision):
    """One encoder pre-training step (pretrain.py:37-48, criterion F.l1_loss): all 72 parameter gradients of a 2-block RRDBNet against
    the REAL reference's summaries (norm / probe per tensor) and against the oracle's autograd (full tensors)."""
    import math
    import torch.nn.functional as F
    from oracle.cases import calibrate_rrdb_head, grad_summary
    from test_pretrain_cpu import _rrdb_oracle_grads
    g, spec = load_golden("rrdb_pretrain"), CASES["rrdb_pretrain"]
    R = wsr.sub("models.rrdb_encoder.RRDBNet").RRDBNet
    net = calibrate_rrdb_head(fill_module(R(1, 1, 64, spec["nb"], 32, precision=precision), spec["seed"])).cuda().train()
    pred = net(g["lr"].cuda())
    assert pred.requires_grad
    loss = F.l1_loss(pred, g["hr"].cuda())
    loss.backward()
    tol_pred, tol_t, tol_all = (1e-5, 2e-4, 1e-4) if precision == "fp32" else (3e-2, 2.5e-1, 8e-2)
    assert rel_l2(pred.detach().cpu(), g["pred"]) < tol_pred
    assert float(loss.detach()) == pytest.approx(float(g["loss"]), rel=1e-5 if precision == "fp32" else 2e-2)
    _, _, oracle = _rrdb_oracle_grads(g, spec)
    named = dict(net.named_parameters())
    num = den = 0.0
    worst = ("", 0.0)
    gscale = math.sqrt(sum(float(v.double().pow(2).sum()) for v in oracle.values()))
    for n, ref in oracle.items():
        got = named[n].grad.detach().cpu()
        num += float((got.double() - ref.double()).pow(2).sum()); den += float(ref.double().pow(2).sum())
        err = float((got - ref).norm())
        rel = err / max(float(ref.norm()), 1e-30)
        if rel > worst[1] and err > 1e-3 * tol_t * gscale:
            worst = (n, rel)
    total = math.sqrt(num / den)
    print("\n[parity] rrdb pretrain step %s: whole-gradient rel-L2 %.3e, worst tensor %s %.3e" % (precision, total, worst[0], worst[1]))
    assert total < tol_all and worst[1] < tol_t, worst
    summ = grad_summary([(n, named[n].grad) for n in oracle], spec["seed"], full_below=128)
    for n in oracle:
        ref_norm = float(g["norm/" + n])
        if ref_norm > 1e-4 * gscale:
            assert abs(float(summ["norm/" + n]) - ref_norm) <= (2e-3 if precision == "fp32" else 2e-1) * ref_norm, n


def test_rrdb_pretraining_reduces_the_loss_and_eval_path_is_unchanged():
    import torch.nn.functional as F
    from oracle.cases import calibrate_rrdb_head
    R = wsr.sub("models.rrdb_encoder.RRDBNet").RRDBNet
    net = calibrate_rrdb_head(fill_module(R(1, 1, 64, 1, 32, precision="bf16"), 3)).cuda().train()
    opt = wsr.sub("autograd_glue").FusedAdam(net.parameters(), lr=2e-4)
    lr = 0.5 * seeded_randn("rp.lr", (4, 1, 16, 32), 9).cuda()
    hr = wsr.sub("data.dataset_builder").bicubic_sr(lr, 4).clamp(-1, 1)
    losses = []
    for _ in range(12):
        loss = F.l1_loss(net(lr), hr)
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(float(loss.detach()))
    assert losses[-1] < 0.97 * losses[0] and all(b < a for a, b in zip(losses, losses[1:])), losses
    net.eval()
    with torch.no_grad():
        out, feas = net(lr, True)
    assert not out.requires_grad and len(feas) == 2 and bool(torch.isfinite(out).all())
