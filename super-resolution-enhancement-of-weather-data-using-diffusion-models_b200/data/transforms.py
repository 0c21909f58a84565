"""StandardScaling and the batch inverse transform of the reference (data/transforms.py:281-409, 81-138) on the device.

The reference keeps one fitted ``StandardScaling`` per (variable, lr/hr, month) and inverts a batch with a Python loop over
samples and variables (``_inverse_tensor``, :116-138).  Here the statistics of a batch are gathered into two (B, C) tensors
and one kernel launch transforms the whole batch."""
import torch

from .. import _native as nat


class StandardScaling:
    """``transform`` / ``revert`` with the reference's arithmetic: (x - mean) / std and std * x + mean (:391-409).  The
    statistics come from ``from_stats`` (a fitted reference object exposes them as ``_mean`` and ``_std()``)."""

    def __init__(self, mean=0.0, std=1.0):
        self._mean, self._stdv = float(mean), float(std)

    @classmethod
    def from_stats(cls, mean, std):
        return cls(float(mean), float(std))

    def _std(self):
        return self._stdv

    def _apply(self, data, inverse):
        if not data.is_cuda:
            raise nat.WsrError("StandardScaling runs on the CUDA path only (got a %s tensor)" % data.device)
        x = data.to(torch.float32).contiguous()
        planes = 1
        hw = x.numel()
        m = torch.full((planes,), self._mean, device=x.device, dtype=torch.float32)
        s = torch.full((planes,), self._stdv, device=x.device, dtype=torch.float32)
        y = torch.empty_like(x)
        nat.call("wsr_standard_scale", x.data_ptr(), planes, hw, m.data_ptr(), s.data_ptr(), inverse, y.data_ptr(),
                 torch.cuda.current_stream(x.device).cuda_stream)
        return y

    def transform(self, data):
        return self._apply(data, 0)

    def revert(self, data):
        return self._apply(data, 1)


def batch_statistics(transformation_dict, variables, data_type, months):
    """(mean, std) tensors of shape (B, C) for a batch: sample b uses the transform fitted for ``months[b]``
    (reference _inverse_tensor, transforms.py:116-138).  ``transformation_dict[variable][data_type][month]`` must expose
    ``_mean`` and ``_std()`` (the reference's fitted StandardScaling, or the class above)."""
    mean = torch.empty((len(months), len(variables)), dtype=torch.float32)
    std = torch.empty_like(mean)
    for c, var in enumerate(variables):
        for b, mo in enumerate(months):
            t = transformation_dict[var][data_type][mo]
            mean[b, c] = float(t._mean)
            std[b, c] = float(t._std())
    return mean, std


def inverse_batch(tensor, mean, std):
    """Physical units of a (B, C, H, W) batch in ONE launch: out[b, c] = std[b, c] * tensor[b, c] + mean[b, c]."""
    if not tensor.is_cuda:
        raise nat.WsrError("inverse_batch runs on the CUDA path only (got a %s tensor)" % tensor.device)
    x = tensor.to(torch.float32).contiguous()
    b, c, h, w = x.shape
    m = mean.to(device=x.device, dtype=torch.float32).contiguous()
    s = std.to(device=x.device, dtype=torch.float32).contiguous()
    assert m.shape == (b, c) and s.shape == (b, c)
    y = torch.empty_like(x)
    nat.call("wsr_standard_scale", x.data_ptr(), b * c, h * w, m.data_ptr(), s.data_ptr(), 1, y.data_ptr(),
             torch.cuda.current_stream(x.device).cuda_stream)
    return y


def transform_batch(tensor, mean, std):
    """Standardised units of a (B, C, H, W) batch: (tensor - mean[b, c]) / std[b, c]."""
    x = tensor.to(torch.float32).contiguous()
    if not x.is_cuda:
        raise nat.WsrError("transform_batch runs on the CUDA path only (got a %s tensor)" % tensor.device)
    b, c, h, w = x.shape
    m = mean.to(device=x.device, dtype=torch.float32).contiguous()
    s = std.to(device=x.device, dtype=torch.float32).contiguous()
    y = torch.empty_like(x)
    nat.call("wsr_standard_scale", x.data_ptr(), b * c, h * w, m.data_ptr(), s.data_ptr(), 0, y.data_ptr(),
             torch.cuda.current_stream(x.device).cuda_stream)
    return y
