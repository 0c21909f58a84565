"""Tap tables (``WsrTapTable``, include/wsr.h) of the convolutions on the path and of their data gradients.

A convolution is described to the kernels as a list of taps: tap t multiplies the weight matrix ``w[wtap_t]`` with the
input pixel ``in_sub * (g + d_t) + p_t`` of loop position ``g`` and accumulates into output pixel ``g * out_mul + out_p``.
The forward tables are used by the weight-gradient kernel, the ``dgrad_*`` tables turn the gradient w.r.t. the input
of each convolution flavour into a forward convolution over the output gradient:

  * Conv2d(k, stride 1, pad (k-1)/2)                  -> same geometry, weights transposed and flipped (no table needed)
  * Downsample = Conv2d(3, stride 2, pad 1)           -> transposed convolution = four output-phase launches
    (functional_layers.py:79)                            (1, 2, 2, 4 taps) on the flipped weights
  * Upsample = nearest x2 + Conv2d(3, pad 1)          -> one 4x4 stride-2 convolution whose 16 taps are sums of the
    (functional_layers.py:62-67)                         3x3 taps that read the same source pixel
"""
import torch

from . import _native as nat


def _table(GH, GW, OH, OW, in_sub, out_mul, out_py, out_px, taps):
    """taps: list of (py, px, dy, dx, wtap)."""
    assert 0 < len(taps) <= nat.MAX_TAPS
    t = nat.TapTable()
    t.GH, t.GW, t.OH, t.OW = GH, GW, OH, OW
    t.in_sub, t.out_mul, t.out_py, t.out_px = in_sub, out_mul, out_py, out_px
    t.ntaps = len(taps)
    for i, (py, px, dy, dx, wt) in enumerate(taps):
        t.py[i], t.px[i], t.dy[i], t.dx[i], t.wtap[i] = py, px, dy, dx, wt
    return t


def _split(off):
    """offset on the full-resolution grid -> (phase, offset on the 2x-subsampled grid): off = 2*d + p."""
    p = off % 2
    return p, (off - p) // 2


def forward_taps(k, stride, H, W):
    """Conv2d(k, stride, pad (k-1)/2) on an (H, W) input: output (H/stride, W/stride)."""
    pad = (k - 1) // 2
    OH, OW = H // stride, W // stride
    taps = []
    for ky in range(k):
        for kx in range(k):
            if stride == 1:
                taps.append((0, 0, ky - pad, kx - pad, ky * k + kx))
            else:
                py, dy = _split(ky - pad)
                px, dx = _split(kx - pad)
                taps.append((py, px, dy, dx, ky * k + kx))
    return _table(OH, OW, OH, OW, stride, 1, 0, 0, taps)


def forward_upsample_taps(H, W):
    """nearest x2 + Conv2d(3, pad 1) seen from the UPSAMPLED grid (2H, 2W): used with ``up=2`` by the weight gradient."""
    taps = [(0, 0, ky - 1, kx - 1, ky * 3 + kx) for ky in range(3) for kx in range(3)]
    return _table(2 * H, 2 * W, 2 * H, 2 * W, 1, 1, 0, 0, taps)


def dgrad_down_taps(H, W):
    """Data gradient of Conv2d(3, stride 2, pad 1) with input (H, W): four tables, one per phase (py, px) of dX.
    The loop grid is the (H/2, W/2) grid of dY; weights are the transposed + flipped pack (tap ky'*3 + kx')."""
    OHf, OWf = H // 2, W // 2
    out = []
    for py in range(2):
        for px in range(2):
            taps = []
            for ky in range(3):
                if (py + ky - 1) % 2:
                    continue
                for kx in range(3):
                    if (px + kx - 1) % 2:
                        continue
                    taps.append((0, 0, (py + ky - 1) // 2, (px + kx - 1) // 2, ky * 3 + kx))
            out.append(_table(OHf, OWf, H, W, 1, 2, py, px, taps))
    return out


def dgrad_upsample_taps(H, W):
    """Data gradient of nearest x2 + Conv2d(3, pad 1) with input (H, W): a 4x4 stride-2 convolution over dY (2H, 2W).
    Tap (r, s), r, s in {-1, 0, 1, 2}, reads dY[2u + r, 2v + s]; weights from ``upsample_dgrad_weight``."""
    taps = []
    for r in range(-1, 3):
        py, dy = _split(r)
        for s in range(-1, 3):
            px, dx = _split(s)
            taps.append((py, px, dy, dx, (r + 1) * 4 + (s + 1)))
    return _table(H, W, H, W, 2, 1, 0, 0, taps)


def dgrad_weight(weight):
    """OIHW weight of Conv2d(k, pad (k-1)/2) -> OIHW weight of the convolution that computes its data gradient."""
    return weight.detach().permute(1, 0, 2, 3).flip(2, 3)


_ROWSETS = {-1: (2,), 0: (1, 2), 1: (0, 1), 2: (0,)}     # dY row 2u + r receives x[u] through these ky


def upsample_dgrad_weight(weight):
    """(Cout, Cin, 3, 3) weight of the Upsample conv -> (Cin, Cout, 4, 4) weight of its data-gradient convolution."""
    w = weight.detach().to(torch.float32)
    cout, cin = w.shape[:2]
    out = torch.zeros((cin, cout, 4, 4), dtype=torch.float32, device=w.device)
    for r in range(-1, 3):
        for s in range(-1, 3):
            acc = 0
            for ky in _ROWSETS[r]:
                for kx in _ROWSETS[s]:
                    acc = acc + w[:, :, ky, kx]
            out[:, :, r + 1, s + 1] = acc.t()
    return out


def conv_transpose_k8s4_wgrad_taps(h, w):
    """Weight gradient of ConvTranspose2d(k=8, s=4, p=2) (srdiff/unet.py:43-45) with input (h, w) and output (4h, 4w):
        dW[ci][co][ky][kx] = sum_{n,i,j} x[n,ci,i,j] * dY[n,co,4i+ky-2,4j+kx-2]
    which IS the weight gradient of a stride-4 8x8 convolution whose input is dY and whose output gradient is x -- so the
    generic tap-table kernel applies with the operand roles swapped (loop grid = the low-resolution grid, in_sub = 4).
    64 taps = four tables of 16 (two kernel rows each); wtap = ky*8 + kx indexes the reference's weight layout directly."""
    out = []
    for q in range(4):
        taps = []
        for ky in (2 * q, 2 * q + 1):
            py = (ky - 2) % 4
            dy = (ky - 2 - py) // 4
            for kx in range(8):
                px = (kx - 2) % 4
                dx = (kx - 2 - px) // 4
                taps.append((py, px, dy, dx, ky * 8 + kx))
        out.append(_table(h, w, h, w, 4, 1, 0, 0, taps))
    return out
