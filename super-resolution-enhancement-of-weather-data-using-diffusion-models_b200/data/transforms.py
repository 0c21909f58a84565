"""Data transforms of the reference (data/transforms.py:14-493): ``IdentityTransform``, ``StandardScaling`` with its
``LocalStandardScaling`` / ``GlobalStandardScaling`` statistics, ``get_transformation_by_name`` and ``DataTransformer`` (one
fitted transform per (variable, lr/hr, month group); batch inverse transform to physical units).

Where the work happens:
  * per-batch transform / inverse transform of DEVICE tensors: one ``wsr_standard_scale`` launch for the whole batch with
    per-(sample, variable) statistics (``transform_batch`` / ``inverse_batch``) -- the reference loops over samples and
    variables in Python (``_inverse_tensor``, :116-138);
  * fitting (once per run, bounded by reading the store): running count / mean / squared differences merged chunk by chunk
    with the reference's update rule (:281-300).  For a plain ``TimeVariateData`` the chunks are read with
    ``WNPYReader.read_into`` (payloads straight into one buffer, thread pool) and reduced in float64;
  * host tensors (what ``DDPM.get_images`` returns) are transformed with the same formula in torch -- host-side plumbing
    of the data pipeline, not a fallback of the model path."""
import os
from collections import OrderedDict
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

from .. import _native as nat
from .datasets import TimeVariateData
from .npy_reader import WNPYReader
from .utils import date_to_str, find_group_idx, month_windows, validate_group_months_subset

_FIT_CHUNK = 4096        # samples per bulk read while fitting


class Transform:
    def __init__(self, requires_fit, exclude_at_evaluation=False):
        self.requires_fit = requires_fit
        self.exclude_at_evaluation = exclude_at_evaluation

    def transform(self, data):
        raise NotImplementedError()

    def out_channels(self, in_channels):
        return in_channels

    def forward(self, data):
        return self.transform(data)

    __call__ = forward

    def is_data_adaptive(self):
        return self.requires_fit

    def summarize(self):
        return {"transform_type": self.__class__.__name__}


class IdentityTransform(Transform):
    def __init__(self):
        super().__init__(requires_fit=False, exclude_at_evaluation=False)
        self._data_source = None

    def transform(self, data):
        return data

    def revert(self, data):
        return data

    def fit(self, dataset, batch_size=None, previous_transforms=None, disable_fitting_mode=False):
        return self

    def _update_parameters(self, data):
        return None

    def clear_data_source(self):
        self._data_source = None

    def summarize(self):
        s = super().summarize()
        s.update({"identity_transform": True})
        return s


def _scale_on_device(x, mean, std, inverse):
    """x (..., H, W) CUDA tensor, mean / std broadcastable scalars per leading plane -> one kernel launch."""
    x = x.to(torch.float32).contiguous()
    planes = mean.numel()
    hw = x.numel() // planes
    y = torch.empty_like(x)
    nat.call("wsr_standard_scale", x.data_ptr(), planes, hw, mean.data_ptr(), std.data_ptr(), inverse, y.data_ptr(),
             torch.cuda.current_stream(x.device).cuda_stream)
    return y


class StandardScaling(Transform):
    """(x - mean) / std and std * x + mean (reference :391-409); ``fit`` accumulates count, mean and the sum of squared
    differences over chunks (:281-300).  ``_compute_stats`` (what the statistics are taken over) is defined by the
    Local / Global subclasses, as in the reference."""

    def __init__(self, unbiased=True, exclude_at_evaluation=False):
        super().__init__(requires_fit=True, exclude_at_evaluation=exclude_at_evaluation)
        self._count = 0
        self._bias_correction = int(unbiased)
        self._mean = None
        self._squared_differences = None
        self._data_source = None
        self._m64 = self._s64 = None          # float64 shadows of the running statistics (bulk fitting path)

    @classmethod
    def from_stats(cls, mean, std, count=2):
        """A fitted transform with the given scalar mean and standard deviation."""
        t = cls()
        t._count = count
        t._mean = torch.tensor(float(mean), dtype=torch.float32).reshape(1, 1, 1, 1)
        t._squared_differences = torch.tensor(float(std) ** 2 * (count - t._bias_correction), dtype=torch.float32).reshape(1, 1, 1, 1)
        return t

    # ---- fitting ---------------------------------------------------------------------------------------------------
    def fit(self, dataset, batch_size=None, previous_transforms=None, disable_fitting_mode=False):
        if self._data_source is not None:
            raise Exception("[ERROR] Fit should only be called once on adaptive transform objects.")
        if previous_transforms is not None:
            assert isinstance(previous_transforms, list)
            for t in previous_transforms:
                assert isinstance(t, Transform)
        if not dataset.is_time_variate():
            self._fit_to_batch(dataset, [0], previous_transforms)
        else:
            in_fitting_mode = dataset.get_fitting_mode()
            if in_fitting_mode != disable_fitting_mode:
                dataset.set_fitting_mode(disable_fitting_mode)
            if self._bulk_ok(dataset, previous_transforms):
                self._fit_bulk(dataset, batch_size or _FIT_CHUNK)
            elif batch_size is None:
                self._fit_to_batch(dataset, np.arange(len(dataset)), previous_transforms)
            else:
                assert isinstance(batch_size, int)
                idx = np.arange(len(dataset))
                for idx_batch in np.array_split(idx, np.ceil(len(idx) / batch_size)):
                    self._fit_to_batch(dataset, idx_batch, previous_transforms)
            dataset.set_fitting_mode(in_fitting_mode)
        self._fill_data_source(dataset, previous_transforms)
        return self

    def _bulk_ok(self, dataset, previous_transforms):
        return (previous_transforms is None and isinstance(dataset, TimeVariateData) and not dataset.get_transform()
                and hasattr(self, "_compute_stats_np"))

    def _fit_bulk(self, dataset, chunk):
        reader = dataset.wnpy_reader
        n = len(dataset)
        buf = np.empty((min(chunk, n),) + reader.sample_shape(), dtype=np.float32)
        with ThreadPoolExecutor(max_workers=min(16, os.cpu_count() or 1)) as pool:
            for lo in range(0, n, chunk):
                idx = np.arange(lo, min(n, lo + chunk))
                view = buf[:len(idx)]
                reader.read_into(view, dataset.stamps_of(idx), pool)
                count, mean, sq = self._compute_stats_np(view.astype(np.float64))
                self._merge(count, torch.from_numpy(mean).to(torch.float32), torch.from_numpy(sq).to(torch.float32),
                            mean64=mean, sq64=sq)

    def _merge(self, count, mean, sq, mean64=None, sq64=None):
        """Chan et al. pairwise update; the float64 shadows keep the bulk path exact across many chunks."""
        if self._mean is None:
            self._count, self._mean, self._squared_differences = count, mean, sq
            self._m64, self._s64 = mean64, sq64
            return self
        if mean64 is not None and getattr(self, "_m64", None) is not None:
            new = self._count + count
            self._s64 = self._s64 + sq64 + (mean64 - self._m64) ** 2 * ((count * self._count) / new)
            self._m64 = (self._count * self._m64 + count * mean64) / new
            self._count = new
            self._mean = torch.from_numpy(np.asarray(self._m64)).to(torch.float32)
            self._squared_differences = torch.from_numpy(np.asarray(self._s64)).to(torch.float32)
            return self
        return self._update_stats(count, mean, sq)

    def _update_stats(self, data_count, data_mean, data_squared_differences):
        new_count = self._count + data_count
        self._squared_differences = self._squared_differences + data_squared_differences \
            + (data_mean - self._mean) ** 2 * ((data_count * self._count) / new_count)
        self._mean = ((self._count * self._mean) + (data_count * data_mean)) / new_count
        self._count = new_count
        self._m64 = self._s64 = None
        return self

    def _fit_to_batch(self, dataset, batch, previous_transforms):
        for data in dataset.get_batch(batch):
            if previous_transforms is not None:
                for t in previous_transforms:
                    data = t.transform(data)
            self._update_parameters(data)

    def _update_parameters(self, data):
        stats = self._compute_stats(data)
        if self._mean is None:
            self._count, self._mean, self._squared_differences = stats
            self._m64 = self._s64 = None
            return self
        return self._update_stats(*stats)

    def _fill_data_source(self, dataset, previous_transforms):
        self._data_source = dataset.summarize()
        if previous_transforms is not None:
            self._data_source.update({"previous_transforms": [t.summarize() for t in reversed(previous_transforms)]})

    def clear_data_source(self):
        self._data_source = None

    # ---- application -----------------------------------------------------------------------------------------------
    def _std(self):
        return torch.sqrt(self._squared_differences / (self._count - self._bias_correction))

    def _apply(self, data, inverse):
        mean, std = self._mean, self._std()
        if data.is_cuda and mean.numel() == 1:
            m = mean.reshape(1).to(device=data.device, dtype=torch.float32)
            s = std.reshape(1).to(device=data.device, dtype=torch.float32)
            return _scale_on_device(data, m, s, inverse)
        mean, std = mean.to(data.device), std.to(data.device)
        return (std * data) + mean if inverse else (data - mean) / std

    def transform(self, data):
        return self._apply(data, 0)

    def revert(self, data):
        return self._apply(data, 1)

    def summarize(self):
        s = super().summarize()
        fitted = self._mean is not None and self._count
        s.update({"mean": self._mean if fitted else None, "std": self._std() if fitted else None})
        return s


class LocalStandardScaling(StandardScaling):
    """Statistics per grid point: over the sample axis only (reference :423-438)."""

    def _compute_stats(self, data):
        mean = torch.mean(data, dim=0, keepdim=True)
        return data.shape[0], mean, torch.sum(torch.square(data - mean), dim=0, keepdim=True)

    def _compute_stats_np(self, data):
        mean = data.mean(axis=0, keepdims=True)
        return data.shape[0], mean, np.square(data - mean).sum(axis=0, keepdims=True)


class GlobalStandardScaling(StandardScaling):
    """One mean / standard deviation per channel: over samples, latitude and longitude (reference :441-463)."""

    def _compute_stats(self, data):
        shape = data.shape
        mean = torch.mean(data, dim=(0, 2, 3), keepdim=True)
        return shape[0] * shape[2] * shape[3], mean, torch.sum(torch.square(data - mean), dim=(0, 2, 3), keepdim=True)

    def _compute_stats_np(self, data):
        shape = data.shape
        mean = data.mean(axis=(0, 2, 3), keepdims=True)
        return shape[0] * shape[2] * shape[3], mean, np.square(data - mean).sum(axis=(0, 2, 3), keepdims=True)


def get_transformation_by_name(name):
    if name == "GlobalStandardScaling":
        return GlobalStandardScaling
    if name == "LocalStandardScaling":
        return LocalStandardScaling
    if name == "IdentityTransform":
        return IdentityTransform
    raise Exception("[ERROR] Unknown transformation <{}>.".format(name))


class DataTransformer:
    """``transformation_dict[variable][lr|hr][month]`` = the transform fitted on that month's GROUP (months of a group share
    one object), reference :14-180."""

    def __init__(self, variables: list, dataroot: str, months_subset, groups=None):
        self.transformation_dict = {}
        self.variables = variables
        self.dataroot = dataroot
        self.groups = groups
        self.months_subset = months_subset

    def transform(self, min_date: str, max_date: str, data_type: str, variable: str, transformation) -> dict:
        validate_group_months_subset(self.months_subset, self.groups)
        fitted = {g: self._fit_month(ds, transformation) for g, ds in self._create_group_datasets(min_date, max_date, data_type, variable).items()}
        mapped = {}
        for idx, group in enumerate(self.groups):
            for month in group:
                mapped[month] = fitted[idx + 1]
        self.transformation_dict.setdefault(variable, {})[data_type] = mapped
        return mapped

    def get_transform(self, variable: str, data_type: str) -> dict:
        return self.transformation_dict[variable][data_type]

    def _create_group_datasets(self, min_date, max_date, data_type, variable):
        """group index (1-based) -> list of per-month ``TimeVariateData`` windows inside [min_date, max_date)."""
        reader = WNPYReader(os.path.join(self.dataroot, data_type, variable))
        out = {}
        for start, end in month_windows(min_date, max_date):
            g = find_group_idx(start.month, self.groups)
            if g is not None:
                out.setdefault(g, []).append(TimeVariateData(reader, name=f"{variable}_{data_type}{date_to_str(start)}", lead_time=0,
                                                             min_date=date_to_str(start), max_date=date_to_str(end)))
        return out

    @staticmethod
    def _fit_month(datasets, transformation):
        t = transformation()
        for d in datasets:
            t.fit(d)
            t.clear_data_source()
        return t

    # ---- batches ---------------------------------------------------------------------------------------------------
    def batch_statistics(self, data_type, months):
        return batch_statistics(self.transformation_dict, self.variables, data_type, months)

    def inverse_transform(self, data: dict, batch_months: list) -> dict:
        out = OrderedDict()
        for key, tensor in data.items():
            out[key] = self._inverse_tensor(tensor, "lr" if key == "LR" else "hr", batch_months)
        return out

    def _inverse_tensor(self, tensor, data_type, months_subset):
        """(B, V, H, W) standardised -> physical units; sample b uses the transform of ``months_subset[b]``."""
        per = [[self.transformation_dict[v][data_type][m] for m in months_subset] for v in self.variables]
        scalar = all(isinstance(t, StandardScaling) and t._mean.numel() == 1 for row in per for t in row)
        if scalar and tensor.shape[1] == len(self.variables):
            mean, std = self.batch_statistics(data_type, months_subset)
            if tensor.is_cuda:
                return inverse_batch(tensor, mean, std)
            b, c = mean.shape
            return std.view(b, c, 1, 1) * tensor + mean.view(b, c, 1, 1)
        cols = []
        for vi, row in enumerate(per):
            col = tensor[:, vi].unsqueeze(1)
            cols.append(torch.cat([row[b].revert(col[b]) for b in range(tensor.shape[0])]).reshape(col.shape))
        return torch.cat(cols, dim=1)


def batch_statistics(transformation_dict, variables, data_type, months):
    """(mean, std) tensors of shape (B, C) for a batch: sample b uses the transform fitted for ``months[b]``
    (reference _inverse_tensor, transforms.py:116-138).  ``transformation_dict[variable][data_type][month]`` must expose
    ``_mean`` and ``_std()`` (the reference's fitted StandardScaling, or the class above)."""
    mean = torch.empty((len(months), len(variables)), dtype=torch.float32)
    std = torch.empty_like(mean)
    for c, var in enumerate(variables):
        cache = {}
        for b, mo in enumerate(months):
            if mo not in cache:
                t = transformation_dict[var][data_type][mo]
                cache[mo] = (float(t._mean), float(t._std()))
            mean[b, c], std[b, c] = cache[mo]
    return mean, std


def _planes(tensor, mean, std):
    if not tensor.is_cuda:
        raise nat.WsrError("the batch transform runs on the CUDA path only (got a %s tensor)" % tensor.device)
    x = tensor.to(torch.float32).contiguous()
    b, c = x.shape[:2]
    m = mean.to(device=x.device, dtype=torch.float32).contiguous()
    s = std.to(device=x.device, dtype=torch.float32).contiguous()
    assert m.shape == (b, c) and s.shape == (b, c)
    return x, m, s


def inverse_batch(tensor, mean, std):
    """Physical units of a (B, C, H, W) batch in ONE launch: out[b, c] = std[b, c] * tensor[b, c] + mean[b, c]."""
    x, m, s = _planes(tensor, mean, std)
    return _scale_on_device(x, m, s, 1)


def transform_batch(tensor, mean, std):
    """Standardised units of a (B, C, H, W) batch: (tensor - mean[b, c]) / std[b, c]."""
    x, m, s = _planes(tensor, mean, std)
    return _scale_on_device(x, m, s, 0)
