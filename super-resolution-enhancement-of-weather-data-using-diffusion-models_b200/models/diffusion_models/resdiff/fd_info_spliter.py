"""Parameter containers for the frequency-domain splitter (reference resdiff/fd_info_spliter.py:5-148).
Kernels: wsr_fd_precompute (condition-only branch, hoisted), wsr_fd_gate + wsr_stem_assemble (per step)."""
from torch import nn

from ..nn_modules.functional_layers import EngineOnly


class ResSE(EngineOnly):
    def __init__(self, ch_in, reduction=2):
        super().__init__()
        self.avg_pool = nn.AdaptiveAvgPool2d(1)
        self.fc = nn.Sequential(
            nn.Linear(ch_in, ch_in // reduction, bias=False),
            nn.ReLU(inplace=True),
            nn.Linear(ch_in // reduction, ch_in, bias=False),
            nn.Sigmoid(),
        )


class FD_Info_Spliter(EngineOnly):
    def __init__(self, dim, in_channels, out_channels, image_height=128, image_width=128):
        super().__init__()
        self.in_channels = in_channels
        self.dim = dim
        self.image_height = image_height
        self.image_width = image_width
        self.noise_func = nn.Linear(dim, self.image_width)
        reduction = 1 if in_channels == 1 else 2
        self.noise_resSE = ResSE(in_channels, reduction=reduction)
        self.sigma_resSE = ResSE(in_channels * 2)
        self.avg_pool = nn.AdaptiveAvgPool2d(1)
        self.HF_guided_resSE = ResSE(in_channels * 2)
        self.channel_transform = nn.Conv2d(in_channels * 2, out_channels, 1)
