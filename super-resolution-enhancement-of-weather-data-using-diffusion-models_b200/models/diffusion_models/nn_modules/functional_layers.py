"""Parameter containers mirroring the reference's models/diffusion_models/nn_modules/functional_layers.py
(same class names, constructor arguments and state_dict keys).  They own fp32 parameters in the reference layout;
their arithmetic is executed by the fused engine (see ``unet_plan.py``), not by ``forward``.
"""
from inspect import isfunction

from torch import nn


def exists(x):
    return x is not None


def default(val, d):
    if exists(val):
        return val
    return d() if isfunction(d) else d


class EngineOnly(nn.Module):
    """Base for layers whose arithmetic lives in the CUDA engine: calling them stand-alone is an error, not a
    silent PyTorch fallback."""

    def forward(self, *a, **k):
        raise RuntimeError("%s is executed by the fused B200 engine as part of UNet.forward; it has no stand-alone "
                           "PyTorch path" % type(self).__name__)


class PositionalEncoding(EngineOnly):
    """functional_layers.py:21-41 -- sinusoidal encoding of the continuous noise level (kernel: wsr_noise_embed)."""

    def __init__(self, dim):
        super().__init__()
        self.dim = dim


class Swish(EngineOnly):
    """functional_layers.py:44-47 (fused into wsr_gn_apply / wsr_noise_embed)."""


class Mish(EngineOnly):
    """functional_layers.py:49-52."""


class Upsample(EngineOnly):
    """functional_layers.py:54-67 -- nearest x2 + conv3x3 (kernel: wsr_conv_tc with upsample=1, no materialised copy)."""

    def __init__(self, dim):
        super().__init__()
        self.conv = nn.Conv2d(dim, dim, 3, padding=1)


class Downsample(EngineOnly):
    """functional_layers.py:70-82 -- conv3x3 stride 2."""

    def __init__(self, dim):
        super().__init__()
        self.conv = nn.Conv2d(dim, dim, 3, 2, 1)
