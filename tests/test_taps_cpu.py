"""Host logic of the training step (CPU): the tap tables and repacked weights of ``taps.py`` reproduce the data and
weight gradients of every convolution flavour on the path.  The tables are interpreted here by a few lines of torch
(the same index arithmetic as wsr_conv_taps_simt / wsr_conv_wgrad_simt, include/wsr.h) and compared with autograd."""
import pytest
import torch
import torch.nn.functional as F

import wsr

T = wsr.sub("taps")


def run_taps(x, w_oihw, t, out=None):
    """x (N,C,H,W); w_oihw (Co,Ci,KH,KW) addressed as tap = ky*KW + kx; returns y (N,Co,OH,OW) (accumulating into out)."""
    N, Ci, H, W = x.shape
    Co = w_oihw.shape[0]
    wt = w_oihw.reshape(Co, Ci, -1)
    y = torch.zeros(N, Co, t.OH, t.OW, dtype=x.dtype) if out is None else out
    for gy in range(t.GH):
        for gx in range(t.GW):
            acc = torch.zeros(N, Co, dtype=x.dtype)
            for i in range(t.ntaps):
                iy = t.in_sub * (gy + t.dy[i]) + t.py[i]
                ix = t.in_sub * (gx + t.dx[i]) + t.px[i]
                if 0 <= iy < H and 0 <= ix < W:
                    acc += x[:, :, iy, ix] @ wt[:, :, t.wtap[i]].t()
            y[:, :, gy * t.out_mul + t.out_py, gx * t.out_mul + t.out_px] += acc
    return y


def run_wgrad(x, dy, t, shape, up=1):
    """dw[co][ci][tap] from the table (x possibly read through a nearest x2 upsampling)."""
    N, Ci, H, W = x.shape
    dw = torch.zeros(shape[0], shape[1], shape[2] * shape[3], dtype=x.dtype)
    for gy in range(t.GH):
        for gx in range(t.GW):
            g = dy[:, :, gy * t.out_mul + t.out_py, gx * t.out_mul + t.out_px]
            for i in range(t.ntaps):
                uy = t.in_sub * (gy + t.dy[i]) + t.py[i]
                ux = t.in_sub * (gx + t.dx[i]) + t.px[i]
                if 0 <= uy < H * up and 0 <= ux < W * up:
                    dw[:, :, t.wtap[i]] += g.t() @ x[:, :, uy // up, ux // up]
    return dw.reshape(shape)


def _case(k, stride, upsample, H=6, W=8, ci=3, co=4, n=2, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, ci, H, W, generator=g, dtype=torch.float64, requires_grad=True)
    w = torch.randn(co, ci, k, k, generator=g, dtype=torch.float64, requires_grad=True)
    xin = F.interpolate(x, scale_factor=2, mode="nearest") if upsample else x
    y = F.conv2d(xin, w, stride=stride, padding=(k - 1) // 2)
    dy = torch.randn(y.shape, generator=g, dtype=torch.float64)
    y.backward(dy)
    return x.detach(), w.detach(), y.detach(), dy, x.grad, w.grad


@pytest.mark.parametrize("k,stride", [(3, 1), (1, 1), (3, 2)])
def test_forward_tables_reproduce_conv_and_wgrad(k, stride):
    x, w, y, dy, _, dw = _case(k, stride, False)
    t = T.forward_taps(k, stride, x.shape[2], x.shape[3])
    assert torch.allclose(run_taps(x, w, t), y, atol=1e-10)
    assert torch.allclose(run_wgrad(x, dy, t, w.shape), dw, atol=1e-10)


def test_upsample_wgrad_table():
    x, w, y, dy, _, dw = _case(3, 1, True)
    t = T.forward_upsample_taps(x.shape[2], x.shape[3])
    assert torch.allclose(run_wgrad(x, dy, t, w.shape, up=2), dw, atol=1e-10)


@pytest.mark.parametrize("k", [1, 3])
def test_dgrad_stride1(k):
    x, w, y, dy, dx, _ = _case(k, 1, False)
    t = T.forward_taps(k, 1, x.shape[2], x.shape[3])
    assert torch.allclose(run_taps(dy, T.dgrad_weight(w), t), dx, atol=1e-10)


def test_dgrad_downsample_phases():
    x, w, y, dy, dx, _ = _case(3, 2, False)
    out = torch.zeros_like(dx)
    wd = T.dgrad_weight(w)
    tables = T.dgrad_down_taps(x.shape[2], x.shape[3])
    assert [t.ntaps for t in tables] == [1, 2, 2, 4]
    for t in tables:
        run_taps(dy, wd, t, out=out)
    assert torch.allclose(out, dx, atol=1e-10)


def test_dgrad_upsample_4x4():
    x, w, y, dy, dx, _ = _case(3, 1, True)
    t = T.dgrad_upsample_taps(x.shape[2], x.shape[3])
    assert t.ntaps == 16 and t.in_sub == 2
    wd = T.upsample_dgrad_weight(w).to(torch.float64)
    assert torch.allclose(run_taps(dy, wd, t), dx, atol=1e-5)
