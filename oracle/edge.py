"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

CPU restatement of the steps either side of the sampling loop, each citing the reference line it
follows.  The bicubic arithmetic itself lives in a third-party dependency (torch, pinned 2.0.1 in the reference's
requirements.txt:11; ``F.interpolate(mode='bicubic')`` = cubic convolution, A = -0.75, align_corners=False, clamped taps):
the restatement calls the same torch operator on the CPU, which is what the reference's collate does.
"""
import torch
import torch.nn.functional as F


def collate_sr(lr, scale=4):
    """data/dataset_builder.py:374-380: per-sample ``interpolate(lr, scale_factor=4, mode='bicubic')`` then cat."""
    return torch.cat([F.interpolate(lr[i:i + 1], scale_factor=scale, mode="bicubic") for i in range(lr.shape[0])])


def bicubic_manual(lr, scale=4):
    """The published algorithm spelled out (ATen UpSample.h: cubic_convolution1/2, upsample_get_value_bounded), pure torch
    indexing -- used to cross-check that ``collate_sr`` and the CUDA kernel implement the same formula."""
    a = -0.75
    b, c, h, w = lr.shape

    def coeffs(t):
        x0, x1, x2, x3 = t + 1, t, 1 - t, 2 - t
        return [((a * x0 - 5 * a) * x0 + 8 * a) * x0 - 4 * a, ((a + 2) * x1 - (a + 3)) * x1 * x1 + 1,
                ((a + 2) * x2 - (a + 3)) * x2 * x2 + 1, ((a * x3 - 5 * a) * x3 + 8 * a) * x3 - 4 * a]

    def axis(n):
        o = torch.arange(n * scale, dtype=torch.float32)
        s = (o + 0.5) / scale - 0.5
        f = torch.floor(s)
        return f.long(), coeffs(s - f)

    iy, cy = axis(h)
    ix, cx = axis(w)
    out = torch.zeros(b, c, h * scale, w * scale)
    for p in range(4):
        yy = (iy - 1 + p).clamp(0, h - 1)
        row = torch.zeros(b, c, h * scale, w * scale)
        for q in range(4):
            xx = (ix - 1 + q).clamp(0, w - 1)
            row = row + lr[:, :, yy][:, :, :, xx] * cx[q][None, None, None, :]
        out = out + row * cy[p][None, None, :, None]
    return out


def global_standard_stats(batches, unbiased=True):
    """data/transforms.py:281-420 + GlobalStandardScaling._compute_stats (:452-463): running mean / squared differences over
    (batch, lat, lon) merged batch by batch (Chan's update, :_update_stats), std = sqrt(ssd / (count - bias))."""
    count, mean, ssd = 0, None, None
    for data in batches:
        n = data.shape[0] * data.shape[2] * data.shape[3]
        m = data.mean(dim=(0, 2, 3), keepdim=True)
        s = ((data - m) ** 2).sum(dim=(0, 2, 3), keepdim=True)
        if mean is None:
            count, mean, ssd = n, m, s
        else:
            tot = count + n
            delta = m - mean
            ssd = ssd + s + delta ** 2 * count * n / tot
            mean = mean + delta * n / tot
            count = tot
    return mean, torch.sqrt(ssd / (count - int(unbiased)))


def standard_transform(x, mean, std):
    """transforms.py:391-399."""
    return (x - mean) / std


def standard_revert(x, mean, std):
    """transforms.py:401-409."""
    return std * x + mean


def inverse_tensor(tensor, mean_bc, std_bc):
    """transforms.py:116-138: per variable, per sample ``revert`` with the statistics of that sample's month."""
    out = torch.empty_like(tensor)
    for c in range(tensor.shape[1]):
        for b in range(tensor.shape[0]):
            out[b, c] = standard_revert(tensor[b, c], mean_bc[b, c], std_bc[b, c])
    return out


def error_metrics(pairs):
    """training/metrics.py:75-201: MAE = sum|d| / n, MSE = sum d^2 / n, RMSE = sqrt(MSE), MR = sum d / n over all updates."""
    s_abs = s_sq = s_d = 0.0
    n = 0
    for pred, target in pairs:
        d = (pred - target).double()
        s_abs += float(d.abs().sum())
        s_sq += float((d ** 2).sum())
        s_d += float(d.sum())
        n += pred.numel()
    return {"MAE": s_abs / n, "MSE": s_sq / n, "RMSE": (s_sq / n) ** 0.5, "MR": s_d / n}


def haar_bands(x, levels=4):
    """pytorch_wavelets.DWTForward(J, 'haar', 'symmetric') detail bands per level (see oracle/ref_shims.py for the convention):
    list of (B, C, 3, H / 2^j, W / 2^j)."""
    out, ll = [], x
    for _ in range(levels):
        a, b, c, d = ll[..., 0::2, 0::2], ll[..., 0::2, 1::2], ll[..., 1::2, 0::2], ll[..., 1::2, 1::2]
        out.append(torch.stack([(a + b - c - d) / 2, (a - b + c - d) / 2, (a - b - c + d) / 2], dim=2))
        ll = (a + b + c + d) / 2
    return out


def fft_mse_loss(x, y):
    """models/simple_cnn/loss.py:9-28, as written (real and imaginary parts of the orthonormal FFT)."""
    fx, fy = torch.fft.fftn(x, dim=(2, 3), norm="ortho"), torch.fft.fftn(y, dim=(2, 3), norm="ortho")
    return torch.mean((fx.imag - fy.imag) ** 2) + torch.mean((fx.real - fy.real) ** 2)


def dwt_mse_loss(x, y, J=4):
    """models/simple_cnn/loss.py:31-57."""
    bx, by = haar_bands(x, J), haar_bands(y, J)
    total = 0.0
    for i in range(J):
        for k in range(3):
            total = total + torch.mean((bx[i][:, :, k] - by[i][:, :, k]) ** 2)
    return total


def image_compare_loss(x, y, alpha=0.2, beta=0.1):
    """models/simple_cnn/loss.py:60-76."""
    return alpha * fft_mse_loss(x, y) + beta * dwt_mse_loss(x, y)
