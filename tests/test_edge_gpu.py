"""GPU tests of the device-side collate / scaling / metric kernels through the reference-named host
classes, against the fixture made by the reference's own classes and against the oracle at the benchmark shapes.
Tolerances: fp32 elementwise arithmetic -> rel-L2 <= 1e-6 (bicubic: fused multiply-adds vs torch's plain ops);
metric accumulators are doubles -> 1e-6 relative."""
import numpy as np
import pytest
import torch

import wsr
from conftest import load_golden, rel_l2
from oracle import edge

pytestmark = pytest.mark.gpu

builder = wsr.sub("data.dataset_builder")
transforms = wsr.sub("data.transforms")
metrics = wsr.sub("training.metrics")


def test_bicubic_collate_vs_reference_fixture():
    g = load_golden("edge")
    sr = builder.bicubic_sr(g["lr"].cuda(), 4)
    assert rel_l2(sr.cpu(), g["sr"]) < 1e-6
    batch = builder.collate_batch(g["lr"].cuda(), g["sr"].cuda())
    assert set(batch) == {"HR", "LR", "SR"} and batch["SR"].shape == (3, 2, 32, 64)


@pytest.mark.parametrize("shape,scale", [((64, 1, 32, 64), 4), ((8, 3, 16, 32), 8), ((2, 1, 5, 7), 4), ((1, 1, 1, 1), 4)])
def test_bicubic_at_benchmark_shapes_vs_oracle(shape, scale):
    torch.manual_seed(3)
    lr = torch.randn(shape)
    got = builder.bicubic_sr(lr.cuda(), scale).cpu()
    assert got.shape == (shape[0], shape[1], shape[2] * scale, shape[3] * scale)
    assert rel_l2(got, edge.collate_sr(lr, scale)) < 1e-6
    # size-independent property: the four tap weights sum to one -> constants are preserved
    const = builder.bicubic_sr(torch.full(shape, 2.25).cuda(), scale)
    assert float((const - 2.25).abs().max()) < 1e-5


def test_bicubic_refuses_cpu_tensors():
    with pytest.raises(wsr.pkg.native.WsrError):
        builder.bicubic_sr(torch.zeros(1, 1, 4, 4), 4)


def test_standard_scaling_and_batch_inverse():
    g = load_golden("edge")
    t = transforms.StandardScaling.from_stats(float(g["mean1"]), float(g["std1"]))
    xs = t.transform(g["x"].cuda())
    assert rel_l2(xs.cpu(), g["x_std1"]) < 1e-6
    assert rel_l2(t.revert(xs).cpu(), g["x_back1"]) < 1e-6
    # batch inverse with per-sample months and two variables vs the reference's per-sample loop (transforms.py:116-138)
    tdict = {v: {"hr": {1: transforms.StandardScaling.from_stats(270.0 + i, 9.0 + i), 7: transforms.StandardScaling.from_stats(291.0 - i, 5.5 + i)}}
             for i, v in enumerate(["t2m", "z500"])}
    months = [1, 7, 7, 1]
    mean, std = transforms.batch_statistics(tdict, ["t2m", "z500"], "hr", months)
    torch.manual_seed(5)
    x = torch.randn(4, 2, 128, 256)
    got = transforms.inverse_batch(x.cuda(), mean, std).cpu()
    assert rel_l2(got, edge.inverse_tensor(x, mean, std)) < 1e-6
    back = transforms.transform_batch(got.cuda(), mean, std).cpu()
    assert rel_l2(back, x) < 1e-5


def test_metric_accumulators_vs_reference_fixture_and_fused_inverse():
    g = load_golden("edge")
    vm = metrics.ValidationMetrics(metrics.create_metric_dict("cuda:0"))
    vm.update(g["pred"][:2].cuda(), g["target"][:2].cuda())
    vm.update(g["pred"][2:].cuda(), g["target"][2:].cuda())
    out = vm.compute_metrics()
    for name in ("MAE", "MSE", "RMSE", "MR"):
        ref = float(g["metric_" + name])
        assert abs(float(out[name]) - ref) <= 1e-5 * abs(ref) + 1e-7, name
    assert "RMSE" in vm.metrics2str()
    # stand-alone metric objects behave like the reference's (own accumulators); reset clears
    m = metrics.RMSE("cuda:0")
    m.update(g["pred"].cuda(), g["target"].cuda())
    assert abs(float(m.compute()) - float(g["metric_RMSE"])) < 1e-5
    m.reset()
    assert m.compute() == 0.0
    # RMSE in physical units without materialising the inverse transform: scale = per-(sample, variable) std
    torch.manual_seed(9)
    pred, target = torch.randn(4, 2, 64, 128), torch.randn(4, 2, 64, 128)
    mean = torch.tensor([[270.0, 5000.0]] * 4)
    std = torch.tensor([[9.0, 300.0], [5.5, 280.0], [9.0, 300.0], [5.5, 280.0]])
    ref = edge.error_metrics([(edge.inverse_tensor(pred, mean, std), edge.inverse_tensor(target, mean, std))])
    vm.reset()
    vm.update(pred.cuda(), target.cuda(), scale=std)
    got = vm.compute_metrics()
    for name in ("MAE", "MSE", "RMSE", "MR"):
        assert abs(float(got[name]) - ref[name]) <= 2e-5 * abs(ref[name]) + 1e-6, name


def test_edge_kernels_fallback_paths():
    """Shapes the vectorised kernels do not take (plane sizes that are not multiples of 4, bases that are not 16-byte aligned, a
    scale other than 4 / 8) run on the scalar kernels with the same results."""
    torch.manual_seed(11)
    # generic bicubic kernel (scale 2 and 3)
    for scale in (2, 3):
        lr = torch.randn(3, 2, 6, 10)
        assert rel_l2(builder.bicubic_sr(lr.cuda(), scale).cpu(), edge.collate_sr(lr, scale)) < 1e-6
    # 5 x 7 planes: hw = 35
    x = torch.randn(4, 2, 5, 7) * 7 + 280
    mean, std = torch.tensor([[270.0, 5000.0]] * 4), torch.tensor([[9.0, 300.0]] * 4)
    back = transforms.transform_batch(transforms.inverse_batch(x.cuda(), mean, std), mean, std).cpu()
    assert rel_l2(transforms.inverse_batch(x.cuda(), mean, std).cpu(), edge.inverse_tensor(x, mean, std)) < 1e-6 and rel_l2(back, x) < 1e-5
    # misaligned base: a view that starts one element into its storage
    flat = torch.randn(2 * 1 * 16 * 32 + 1).cuda()
    v = flat[1:].view(2, 1, 16, 32)
    assert v.data_ptr() % 16 != 0
    m1, s1 = torch.tensor([[1.5], [-2.0]]), torch.tensor([[2.0], [0.5]])
    assert rel_l2(transforms.inverse_batch(v, m1, s1).cpu(), edge.inverse_tensor(v.cpu(), m1, s1)) < 1e-6
    # error sums on odd plane sizes and with a misaligned operand
    pred, target = torch.randn(4, 2, 5, 7), torch.randn(4, 2, 5, 7)
    ref = edge.error_metrics([(pred, target)])
    vm = metrics.ValidationMetrics(metrics.create_metric_dict("cuda:0"))
    vm.update(pred.cuda(), target.cuda())
    got = vm.compute_metrics()
    for name in ("MAE", "MSE", "RMSE", "MR"):
        assert abs(float(got[name]) - ref[name]) <= 1e-5 * abs(ref[name]) + 1e-7, name
    ref2 = edge.error_metrics([(v.cpu(), torch.zeros_like(v.cpu()))])
    vm.reset()
    vm.update(v, torch.zeros_like(v))
    got2 = vm.compute_metrics()
    assert abs(float(got2["RMSE"]) - ref2["RMSE"]) <= 1e-5 * ref2["RMSE"]
