"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Deterministic, order-independent synthetic weights: every tensor is drawn from its own CPU generator seeded by
(seed, crc32(name)), so the reference model (in the build container) and the CUDA model (on the GPU box) get
bit-identical parameters from nothing but the list of (name, shape) pairs.
"""
import zlib

import torch


def seeded_tensor(name, shape, seed):
    g = torch.Generator(device="cpu")
    g.manual_seed((int(seed) * 1000003 + zlib.crc32(name.encode())) % (2 ** 63 - 1))
    shape = tuple(shape)
    leaf = name.rsplit(".", 1)[-1]
    if len(shape) >= 2:                                   # conv / linear weight: variance-preserving
        fan_in = 1
        for s in shape[1:]:
            fan_in *= s
        return torch.randn(shape, generator=g) * (1.0 / fan_in) ** 0.5
    if leaf == "weight":                                  # GroupNorm gamma
        return 1.0 + 0.1 * torch.randn(shape, generator=g)
    return 0.1 * torch.randn(shape, generator=g)          # biases / GroupNorm beta


def seeded_state_dict(named_shapes, seed):
    """named_shapes: iterable of (name, shape).  Returns {name: fp32 tensor}."""
    return {n: seeded_tensor(n, s, seed) for n, s in named_shapes}


def fill_module(module, seed, skip_prefixes=()):
    """Load seeded weights into every parameter of ``module`` (buffers untouched)."""
    with torch.no_grad():
        for n, p in module.named_parameters():
            if any(n.startswith(s) for s in skip_prefixes):
                continue
            p.copy_(seeded_tensor(n, p.shape, seed))
    return module


def seeded_randn(tag, shape, seed):
    g = torch.Generator(device="cpu")
    g.manual_seed((int(seed) * 7919 + zlib.crc32(tag.encode())) % (2 ** 63 - 1))
    return torch.randn(tuple(shape), generator=g)
