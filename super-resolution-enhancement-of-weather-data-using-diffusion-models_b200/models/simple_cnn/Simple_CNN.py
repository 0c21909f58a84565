"""SimpleCNN prior -- parameter container + CUDA forward, mirroring the reference's models/simple_cnn/Simple_CNN.py:10-32
(3 convs, ReLU, PixelShuffle(4), + bicubic x4 skip).  Optional pre-step that fills ``x_in['SR']``."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from ... import _native as nat
from ...engine import Engine


class SimpleCNN(nn.Module):
    def __init__(self, scale_factor=4, channels=1):
        super().__init__()
        self.scale_factor = scale_factor
        self.channels = channels
        self.conv1 = nn.Conv2d(channels, 64, kernel_size=3, stride=1, padding=1, bias=True)
        self.relu1 = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(64, 32, kernel_size=3, stride=1, padding=1, bias=True)
        self.relu2 = nn.ReLU(inplace=True)
        self.conv3 = nn.Conv2d(32, channels * scale_factor ** 2, kernel_size=3, stride=1, padding=1, bias=True)
        self.pixel_shuffle = nn.PixelShuffle(scale_factor)

    @torch.no_grad()
    def forward(self, x):
        """x: LR (B,C,h,w) on a CUDA device.  The three convolutions run on the fp32 SIMT kernel (0.1 GFLOP per sample,
        once per sample); bicubic interpolation and the pixel shuffle are index plumbing done by torch."""
        eng = Engine(x.device, "fp32")
        B, Cc, h, w = x.shape
        a0 = eng.nchw_to_act(x.to(torch.float32), eng.new_act(B, h, w, Cc))
        p1 = eng.pack_conv(self.conv1.weight, self.conv1.bias)
        p2 = eng.pack_conv(self.conv2.weight, self.conv2.bias)
        p3 = eng.pack_conv(self.conv3.weight, self.conv3.bias)
        a1 = eng.conv(a0, p1, eng.new_act(B, h, w, 64), act=nat.ACT_RELU)
        a2 = eng.conv(a1, p2, eng.new_act(B, h, w, 32), act=nat.ACT_RELU)
        a3 = eng.conv(a2, p3, eng.new_act(B, h, w, Cc * self.scale_factor ** 2))
        y = a3.to_nchw(eng)
        up = F.interpolate(x.to(torch.float32), scale_factor=4, mode='bicubic', align_corners=False)
        return self.pixel_shuffle(y) + up
