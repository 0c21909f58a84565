"""CPU tests of the host-side logic: config parser, C-ABI library exports, state_dict compatibility, registry errors,
schedule buffers of the product GaussianDiffusion vs the reference golden fixture, no-CPU-fallback behaviour."""
import argparse
import copy
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest
import torch

import wsr
from conftest import ROOT, load_golden, manifest

nat = wsr.pkg.native


def test_c_abi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "wsr.h")).read()
    declared = set(re.findall(r"\b(wsr_[A-Za-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations found"
    lib_path = nat.LIB_PATH
    if not os.path.exists(lib_path):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(lib_path)
    for name in sorted(declared):
        assert hasattr(lib, name), "libwsr.so does not export %s" % name
    assert declared == set(nat.SIGNATURES) | {"wsr_last_error"}, declared ^ (set(nat.SIGNATURES) | {"wsr_last_error"})
    assert lib.wsr_version() >= 100
    out = subprocess.run(["nm", "-D", lib_path], capture_output=True, text=True).stdout
    assert " T wsr_conv_tc" in out and " T wsr_attention_tc" in out


def test_no_cpu_fallback():
    Engine = wsr.sub("engine").Engine
    with pytest.raises(nat.WsrError):
        Engine("cpu", "bf16")
    U = wsr.sub("models.diffusion_models.resdiff.unet").UNet
    net = U(in_channel=5, out_channel=1, inner_channel=64, channel_mults=[1, 2], attn_res=[], res_blocks=1,
            image_height=16, image_width=16, image_channels=1)
    with pytest.raises(nat.WsrError):
        net(torch.zeros(1, 2, 16, 16), torch.zeros(1, 1))
    with pytest.raises(RuntimeError):
        net.downs[1](torch.zeros(1, 64, 16, 16), None)          # layers have no stand-alone PyTorch path


def test_invalid_arguments_are_rejected_before_any_launch():
    d = nat.ConvDesc()
    rc = nat.lib().wsr_conv_simt(ctypes.byref(d), None)
    assert rc == -1 and "null" in nat.last_error()
    with pytest.raises(nat.WsrError):
        nat.call("wsr_softmax_rows", 0, 0, 0, 0, 0, 1.0, 0, 0, 0, None)


@pytest.mark.parametrize("tag,modname,cls,kw", [
    ("resdiff", "models.diffusion_models.resdiff.unet", "UNet", dict(in_channel=5)),
    ("srdiff", "models.diffusion_models.srdiff.unet", "UNet", dict(in_channel=1)),
    ("sr3", "models.diffusion_models.sr3.unet", "UNet", dict(in_channel=2)),
    ("phydiff", "models.diffusion_models.phydiff.unet", "UNet", dict(in_channel=5)),
])
def test_state_dict_keys_and_shapes_match_reference(tag, modname, cls, kw):
    U = getattr(wsr.sub(modname), cls)
    net = U(out_channel=1, norm_groups=32, inner_channel=64, channel_mults=[1, 2, 4, 8, 8], attn_res=[16], res_blocks=2,
            dropout=0.2, image_height=128, image_width=256, image_channels=1, **kw)
    got = [(k, tuple(v.shape)) for k, v in net.state_dict().items()]
    assert got == manifest(tag)


def test_prior_state_dicts_match_reference():
    R = wsr.sub("models.rrdb_encoder.RRDBNet").RRDBNet
    S = wsr.sub("models.simple_cnn.Simple_CNN").SimpleCNN
    assert [(k, tuple(v.shape)) for k, v in R(1, 1, 64, 17, 32).state_dict().items()] == manifest("rrdb")
    assert [(k, tuple(v.shape)) for k, v in S(4, 1).state_dict().items()] == manifest("simple_cnn")


def test_schedule_buffers_match_reference_fixture():
    from oracle.cases import LINEAR_1000
    from oracle.schedule import BUFFER_NAMES
    D = wsr.sub("models.diffusion_models.resdiff.resdiff_diffusion").ResDiffDiffusion
    diff = D(torch.nn.Identity(), image_height=8, image_width=8, channels=1)
    g = load_golden("schedule")
    for tag, opt in {"linear1000": LINEAR_1000,
                     "cosine20": {"schedule": "cosine", "n_timestep": 20, "linear_start": 1e-4, "linear_end": 2e-2},
                     "warmup10_40": {"schedule": "warmup10", "n_timestep": 40, "linear_start": 1e-4, "linear_end": 2e-2}}.items():
        diff.set_new_noise_schedule(opt, "cpu")
        for n in BUFFER_NAMES:
            ref = g["%s.%s" % (tag, n)].numpy()
            got = getattr(diff, n).numpy()
            fin = np.isfinite(ref)
            np.testing.assert_allclose(got[fin], ref[fin], rtol=2e-6, err_msg=tag + "." + n)
        np.testing.assert_allclose(diff.sqrt_alphas_cumprod_prev, g["%s.sqrt_alphas_cumprod_prev" % tag].numpy(), rtol=1e-12)
    assert set(BUFFER_NAMES) <= set(dict(diff.named_buffers()))
    with pytest.raises(NotImplementedError):
        diff.set_new_noise_schedule({"schedule": "nope", "n_timestep": 4, "linear_start": 1e-4, "linear_end": 1e-2}, "cpu")


def test_registry_and_config_parser(tmp_path):
    cfgmod = wsr.sub("configs.config")
    src = os.path.join(ROOT, "configs_examples", "resdiff_eval_b200.json")
    ns = argparse.Namespace(config=src, gpu_ids="0", phase=None)
    opt = cfgmod.Config(ns, experiment=False).params
    assert opt["model"]["architecture"] == "resdiff" and opt["gpu_ids"] == "0" and opt["distributed"] is False
    assert opt["data"]["transform_groups"] == [[1]]
    assert cfgmod.strip_comments('{"a": 1, // c\n"b": 2}') == '{"a": 1, \n"b": 2}'
    networks = wsr.sub("models.diffusion_models.networks")
    bad = {"model": {"architecture": "nope"}}
    with pytest.raises(NotImplementedError):
        networks.define_diffusion(bad)
    for arch, cls, inc in (("sr3", "SR3Diffusion", 2), ("phydiff", "PhyDiffDiffusion", 5)):
        o2 = copy.deepcopy(opt)
        o2["model"]["architecture"] = arch
        o2["model"]["unet"]["in_channel"] = inc
        m2 = networks.define_diffusion(o2)
        assert type(m2).__name__ == cls and type(m2.denoise_fn).__module__.endswith(arch + ".unet")
    o3 = copy.deepcopy(opt)
    o3["model"]["architecture"] = "physrdiff"               # broken in the reference itself
    with pytest.raises(NotImplementedError):
        networks.define_diffusion(o3)
    opt["phase"] = "train"
    m = networks.define_diffusion(opt)                      # builds on CPU: parameters only, orthogonal init
    assert type(m).__name__ == "ResDiffDiffusion" and sum(p.numel() for p in m.parameters()) == 98870734
    w = m.denoise_fn.downs[1].res_block.block1.block[3].weight.detach().reshape(64, -1)
    assert torch.allclose(w @ w.t(), torch.eye(64), atol=1e-4)        # orthogonal rows, zero bias
    assert float(m.denoise_fn.downs[1].res_block.block1.block[3].bias.abs().max()) == 0.0
