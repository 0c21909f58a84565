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

    def _run(self, x, keep=False):
        eng = Engine(x.device, "fp32")
        B, Cc, h, w = x.shape
        xf = x.detach().to(torch.float32).contiguous()
        a0 = eng.nchw_to_act(xf, eng.new_act(B, h, w, Cc))
        p1 = eng.pack_conv(self.conv1.weight, self.conv1.bias)
        p2 = eng.pack_conv(self.conv2.weight, self.conv2.bias)
        p3 = eng.pack_conv(self.conv3.weight, self.conv3.bias)
        a1 = eng.conv(a0, p1, eng.new_act(B, h, w, 64), act=nat.ACT_RELU)
        a2 = eng.conv(a1, p2, eng.new_act(B, h, w, 32), act=nat.ACT_RELU)
        a3 = eng.conv(a2, p3, eng.new_act(B, h, w, Cc * self.scale_factor ** 2))
        y = a3.to_nchw(eng)
        from ...data.dataset_builder import bicubic_sr
        up = bicubic_sr(xf, 4)                                  # the reference hard-codes 4 here (Simple_CNN.py:25)
        out = self.pixel_shuffle(y) + up
        return (out, (eng, a0, a1, a2)) if keep else out

    def forward(self, x):
        """x: LR (B,C,h,w) on a CUDA device.  The three convolutions run on the fp32 SIMT kernel (0.1 GFLOP per sample);
        the pixel shuffle is index plumbing done by torch.  With gradients enabled the result carries a ``grad_fn`` whose backward
        is the hand-written pass below (pre-training, reference pretrain.py:19-53)."""
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            return _SimpleCNNFn.apply(self.conv1.weight, self, x)
        with torch.no_grad():
            return self._run(x)

    @torch.no_grad()
    def backward_pass(self, saved, d_out):
        """Parameter gradients (accumulated into ``p.grad``) from d loss / d output (B, C, 4h, 4w): pixel-unshuffle, then for each
        convolution the weight / bias gradient (``wsr_conv_wgrad_simt``) and the data gradient (the forward kernel on transposed,
        flipped weights) with the ReLU masks in between."""
        from ... import taps as tp
        eng, a0, a1, a2 = saved
        B, h, w = a0.N, a0.H, a0.W
        d3 = eng.nchw_to_act(F.pixel_unshuffle(d_out.to(torch.float32).contiguous(), self.scale_factor).contiguous(),
                             eng.new_act(B, h, w, self.conv3.out_channels))
        table = tp.forward_taps(3, 1, h, w)
        st = eng.stream

        def grads_of(conv):
            for p in (conv.weight, conv.bias):
                if p.grad is None:
                    p.grad = torch.zeros_like(p)
            return conv.weight.grad, conv.bias.grad

        def wgrad(conv, x_act, dy_act):
            gw, gb = grads_of(conv)
            co, ci, kh, kw = conv.weight.shape
            eng.wgrad(x_act, dy_act, table, gw, (1, ci * kh * kw, kh * kw), gb, 1, force_simt=True)

        def dgrad(conv, dy_act, y_prev):
            pc = eng.pack_conv(tp.dgrad_weight(conv.weight).contiguous(), None, key=("dgrad", conv.weight.data_ptr(), conv.weight._version))
            dx = eng.conv(dy_act, pc, eng.new_act(B, h, w, conv.in_channels), bias=False, force_simt=True)
            nat.call("wsr_relu_mask", y_prev.ptr, dx.ptr, B * h * w * conv.in_channels, st)      # y_prev = relu output feeding `conv`
            return dx

        wgrad(self.conv3, a2, d3)
        d2 = dgrad(self.conv3, d3, a2)
        wgrad(self.conv2, a1, d2)
        d1 = dgrad(self.conv2, d2, a1)
        wgrad(self.conv1, a0, d1)
        eng._keep.clear()


class _SimpleCNNFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, anchor, net, x):
        out, saved = net._run(x, keep=True)
        ctx.net, ctx.saved = net, saved
        return out

    @staticmethod
    def backward(ctx, d_out):
        ctx.net.backward_pass(ctx.saved, d_out)
        return None, None, None
