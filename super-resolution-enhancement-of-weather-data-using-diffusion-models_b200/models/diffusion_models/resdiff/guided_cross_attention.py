"""Parameter container for the high-frequency guided cross-attention (reference resdiff/guided_cross_attention.py:6-44).
Q = 1x1 conv of the Haar detail image (condition-only, hoisted), K/V = 1x1 conv of GroupNorm(feature)."""
from torch import nn

from ..nn_modules.functional_layers import EngineOnly


class HF_guided_CA(EngineOnly):
    def __init__(self, in_channel, norm_groups=32, image_channels=3, wavelet_components=1):
        super().__init__()
        self.norm = nn.GroupNorm(norm_groups, in_channel)
        self.q = nn.Conv2d(image_channels * wavelet_components, in_channel, 1, bias=False)
        self.kv = nn.Conv2d(in_channel, in_channel * 2, 1, bias=False)
        self.out = nn.Conv2d(in_channel, in_channel, 1)
