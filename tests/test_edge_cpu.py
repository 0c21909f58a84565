"""CPU tests of the N2 oracle (oracle/edge.py) against the fixture produced by the reference's own classes
(tests/golden/edge.npz <- oracle/make_golden.py edge): collate interpolation, GlobalStandardScaling, MAE/MSE/RMSE/MR."""
import numpy as np
import torch

from conftest import load_golden, rel_l2
from oracle import edge


def test_collate_bicubic_matches_reference_and_the_published_formula():
    g = load_golden("edge")
    sr = edge.collate_sr(g["lr"], 4)
    assert torch.equal(sr, g["sr"])
    assert rel_l2(edge.bicubic_manual(g["lr"], 4), g["sr"]) < 1e-6
    # known answers: a constant field stays constant (the four weights sum to 1); on a unit ramp the interior output rises by
    # exactly 1 per source pixel (= per 4 outputs) -- A = -0.75 is not linear-preserving sample by sample, only period by period
    assert torch.allclose(edge.bicubic_manual(torch.full((1, 1, 4, 4), 3.5), 4), torch.full((1, 1, 16, 16), 3.5), atol=1e-6)
    ramp = torch.arange(8, dtype=torch.float32).view(1, 1, 1, 8).repeat(1, 1, 4, 1)
    up = edge.bicubic_manual(ramp, 4)[0, 0, 8, 8:24]
    assert torch.allclose(up[4:] - up[:-4], torch.ones(12), atol=1e-5) and bool((up[1:] > up[:-1]).all())


def test_standard_scaling_statistics_and_round_trip():
    g = load_golden("edge")
    for month in (1, 7):
        mean, std = edge.global_standard_stats(list(g["fit%d" % month]))
        assert abs(float(mean) - float(g["mean%d" % month])) < 1e-4 * abs(float(g["mean%d" % month]))
        assert abs(float(std) - float(g["std%d" % month])) < 1e-5 * float(g["std%d" % month])
    m, s = g["mean1"], g["std1"]
    xs = edge.standard_transform(g["x"], m, s)
    assert rel_l2(xs, g["x_std1"]) < 1e-6
    assert rel_l2(edge.standard_revert(xs, m, s), g["x_back1"]) < 1e-6
    assert rel_l2(g["x_back1"], g["x"]) < 1e-6            # transform -> revert is the identity


def test_error_metrics_match_reference():
    g = load_golden("edge")
    got = edge.error_metrics([(g["pred"][:2], g["target"][:2]), (g["pred"][2:], g["target"][2:])])
    for name in ("MAE", "MSE", "RMSE", "MR"):
        ref = float(g["metric_" + name])
        assert abs(got[name] - ref) <= 1e-5 * abs(ref) + 1e-7, name
    assert abs(got["RMSE"] - np.sqrt(got["MSE"])) < 1e-12
