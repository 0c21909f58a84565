"""Date-indexed datasets over the on-disk store (reference data/datasets.py:97-861): ``TimeVariateData``
(one variable, one or several disjoint date ranges -- the month-subset datasets are unions of monthly ranges),
``ConstantData`` and ``WeatherBenchData`` (groups 'lr' / 'hr' of variables with common date bounds).  Same class names,
constructor arguments, item format ``(tensor (1, C, H, W), name, month)`` and error behaviour as the reference.

Differences in construction, not in behaviour: the sample index is ONE ``datetime64`` array (the reference builds a Python
dict ``{i: time stamp}`` with one entry per hour of data) and the disjoint date ranges are a sorted list (the reference
needs the third-party ``intervaltree`` package for the same overlap query).  ``stamps_of`` / ``months_of`` expose the index
in bulk for the device loader (``dataset_builder.DeviceBatchLoader``)."""
from collections import OrderedDict
from datetime import datetime, timezone

import numpy as np
import torch
from torch.utils.data import Dataset

from ..configs.config import DataConfig
from .npy_reader import WNPYReader
from .utils import str_to_date

config = DataConfig()
TEMPORAL_RESOLUTION = np.timedelta64(config.temporal_resolution_value, config.temporal_resolution_unit)
DATETIME_FORMAT = config.datetime_format
_EPOCH_ALIGNED = np.datetime64("2020-01-01T00")


def _parse_date_input(date_input, datetime_format=None):
    """str / datetime / datetime64 / None -> datetime64 / None."""
    if date_input is None:
        return None
    kind = type(date_input)
    if kind == np.datetime64:
        return date_input
    if kind == datetime:
        return np.datetime64(date_input)
    if kind == str:
        fmt = DATETIME_FORMAT if datetime_format is None else datetime_format
        try:
            return np.datetime64(datetime.strptime(date_input, fmt))
        except Exception:
            raise Exception("[ERROR] Encountered invalid date string input (input: {}, datetime format: {}).".format(date_input, fmt))
    raise Exception("[ERROR] Encountered invalid date input.")


def _verify_date_bounds(min_date, max_date):
    assert (isinstance(min_date, np.datetime64) or min_date is None) and (isinstance(max_date, np.datetime64) or max_date is None), \
        "[ERROR] Date bounds must be given as numpy.datetime64 objects."
    for d in (min_date, max_date):
        if d is not None:
            assert (d - _EPOCH_ALIGNED) % TEMPORAL_RESOLUTION == np.timedelta64(0, "ms"), \
                "[ERROR] Date bounds must be consistent with the temporal resolution of the data set ({}).".format(TEMPORAL_RESOLUTION)
    if min_date is not None and max_date is not None:
        assert max_date > min_date, "[ERROR] Lower date bound ({}) must be earlier than upper ({}).".format(min_date, max_date)


def month_of(time_stamp):
    """1..12 of a datetime64 (scalar or array)."""
    return np.asarray(time_stamp).astype("datetime64[M]").astype(np.int64) % 12 + 1


class DefaultIdentityMapping(dict):
    """month -> transform mapping whose missing keys are the identity."""

    def __missing__(self, key):
        return lambda x: x


class DateRanges:
    """Sorted set of disjoint half-open [begin, end) date ranges."""

    def __init__(self):
        self._ranges = []

    def overlaps(self, begin, end):
        return any(b < end and e > begin for b, e in self._ranges)

    def add(self, begin, end):
        self._ranges.append((begin, end))
        self._ranges.sort(key=lambda r: r[0])

    def __len__(self):
        return len(self._ranges)

    def __iter__(self):
        return iter(self._ranges)

    def span(self):
        return self._ranges[0][0], max(e for _, e in self._ranges)


class TimeVariateData(Dataset):
    def __init__(self, source: WNPYReader, name=None, lead_time=None, delays=None, min_date=None, max_date=None, transform: dict = None):
        assert isinstance(source, WNPYReader)
        assert source.is_time_variate()
        if name is not None:
            assert isinstance(name, str)
        self.name = name if name is not None else source.name
        self.wnpy_reader = source
        self._lead_time = TEMPORAL_RESOLUTION * lead_time if lead_time is not None else None
        if delays is not None:
            assert isinstance(delays, list), "[ERROR] Delay parameter must be given as list."
            for d in delays:
                assert isinstance(d, int), "[ERROR] Delay parameter must be given as list of ints."
            if 0 not in delays:
                delays = [0] + delays
            delays = np.array(delays)
            assert len(delays) == len(np.unique(delays)), "[ERROR] Delays must be unique."
            self._delays = TEMPORAL_RESOLUTION * delays
        else:
            self._delays = None
        self.min_date = self.max_date = None
        self._sample_index = None
        self.set_date_range(min_date, max_date)
        self._fitting_mode = False
        self._transform = transform if transform else DefaultIdentityMapping()
        self.date_ranges = DateRanges()
        self.date_ranges.add(self.min_date, self.max_date)

    # ---- date ranges -----------------------------------------------------------------------------------------------
    def _admissible(self, with_offsets=True):
        stamps = self.wnpy_reader.get_valid_time_stamps()
        lo, hi = np.min(stamps), np.max(stamps) + TEMPORAL_RESOLUTION
        if with_offsets and self._lead_time is not None:
            lo, hi = lo - self._lead_time, hi - self._lead_time
        if with_offsets and self._delays is not None:
            lo, hi = lo - np.min(self._delays), hi - np.max(self._delays)
        return lo, hi

    @staticmethod
    def _check_admissible(min_date, max_date, lo, hi):
        if min_date is not None:
            assert min_date >= lo, "[ERROR] Requested minimum date ({}) is beyond the range of admissible dates ([{}] – [{}]).".format(min_date, lo, hi)
        if max_date is not None:
            assert max_date <= hi, "[ERROR] Requested maximum date ({}) is beyond the range of admissible dates ([{}] – [{}]).".format(max_date, lo, hi)

    def set_date_range(self, min_date=None, max_date=None, datetime_format=None):
        min_date = _parse_date_input(min_date, datetime_format)
        max_date = _parse_date_input(max_date, datetime_format)
        _verify_date_bounds(min_date, max_date)
        lo, hi = self._admissible()
        self._check_admissible(min_date, max_date, lo, hi)
        self.min_date = lo if min_date is None else min_date
        self.max_date = hi if max_date is None else max_date
        self._build_sample_index()
        return self

    def add_data_by_date(self, min_date, max_date, datetime_format=None):
        """Append another disjoint range (reference :209-254); its samples are indexed after the existing ones."""
        min_date = _parse_date_input(min_date, datetime_format)
        max_date = _parse_date_input(max_date, datetime_format)
        _verify_date_bounds(min_date, max_date)
        assert min_date is not None, "[ERROR] Requested minimum date is None."
        assert max_date is not None, "[ERROR] Requested maximum date is None."
        assert not self.date_ranges.overlaps(min_date, max_date), \
            f"[ERROR] Requested date range ({min_date}, {max_date}) overlaps with existing date ranges."
        lo, hi = self._admissible(with_offsets=False)
        self._check_admissible(min_date, max_date, lo, hi)
        self.min_date = min(self.min_date, min_date)
        self.max_date = max(self.max_date, max_date)
        self.date_ranges.add(min_date, max_date)
        self._sample_index = np.concatenate([self._sample_index, np.arange(min_date, max_date, TEMPORAL_RESOLUTION).astype(self._sample_index.dtype)])

    def _build_sample_index(self):
        self._sample_index = np.arange(self.min_date, self.max_date, TEMPORAL_RESOLUTION).astype("datetime64[us]")

    def get_time_intervals(self):
        return ((b, e) for b, e in self.date_ranges)

    def get_valid_time_stamps(self):
        return sorted(self._sample_index)

    # ---- bulk access for the device loader -------------------------------------------------------------------------
    def stamps_of(self, indices):
        """Time stamps (lead time applied) of dataset positions, as one datetime64 array."""
        stamps = self._sample_index[np.asarray(indices, dtype=np.int64)]
        return stamps + self._lead_time if self._lead_time is not None else stamps

    def months_of(self, indices):
        """Month keys of dataset positions as the per-sample interface reports them (0 = no transform fitted for that month)."""
        months = month_of(self._sample_index[np.asarray(indices, dtype=np.int64)])
        return [int(m) if int(m) in self._transform else 0 for m in months]

    # ---- the reference's per-sample interface ----------------------------------------------------------------------
    def set_transform(self, transform: dict):
        self._transform = transform

    def get_transform(self):
        return self._transform

    def __getitem__(self, item):
        time_stamp = item if isinstance(item, np.datetime64) else self._sample_index[item]
        month = int(month_of(time_stamp))
        if month not in self._transform:
            month = 0
        if self._lead_time is not None:
            time_stamp = time_stamp + self._lead_time
        if self._fitting_mode or self._delays is None:
            return self._transform[month](self.wnpy_reader[time_stamp]), self.name, month
        return tuple((self._transform[month](self.wnpy_reader[t]), self.name, month) for t in (time_stamp + self._delays))

    def __len__(self):
        return len(self._sample_index)

    def get_channel_count(self):
        c = self.wnpy_reader.get_channel_count()
        return len(self._delays) * c if self._delays is not None else c

    def get_batch(self, indices, chunk_size=50000):
        """Yields concatenated chunks of at most ``chunk_size`` samples (used by the transform fitting)."""
        indices = list(indices)
        for start in range(0, len(indices), chunk_size):
            items = [self[i] for i in indices[start:start + chunk_size]]
            if self._delays is not None and not self._fitting_mode:
                yield tuple(torch.cat(d[0], dim=0) for d in items)
            else:
                yield torch.cat([it[0] for it in items], dim=0)

    def enable_fitting_mode(self):
        return self.set_fitting_mode(True)

    def disable_fitting_mode(self):
        return self.set_fitting_mode(False)

    def set_fitting_mode(self, mode):
        assert isinstance(mode, bool)
        self._fitting_mode = mode
        return self

    def get_fitting_mode(self):
        return self._fitting_mode

    @staticmethod
    def is_time_variate():
        return True

    def summarize(self):
        lo, hi = self.date_ranges.span()
        return {"data_type": "TimeVariateData", "path": self.wnpy_reader.path,
                "date_range": [_numpy_date_to_datetime(lo).strftime(DATETIME_FORMAT), _numpy_date_to_datetime(hi).strftime(DATETIME_FORMAT)],
                "lead_time": self._lead_time, "delays": self._delays, "name": self.name, "number_of_intervals": len(self.date_ranges)}


def _numpy_date_to_datetime(time_stamp):
    seconds = (time_stamp - np.datetime64("1970-01-01T00:00:00")) / np.timedelta64(1, "s")
    return datetime.fromtimestamp(float(seconds), tz=timezone.utc).replace(tzinfo=None)


class ConstantData(Dataset):
    def __init__(self, source, name=None, min_date=None, max_date=None, datetime_format=None):
        assert isinstance(source, WNPYReader)
        assert not source.is_time_variate()
        if name is not None:
            assert isinstance(name, str)
        self.name = name if name is not None else source.name
        self.wnpy_reader = source
        self._num_samples = None
        self.set_date_range(min_date, max_date, datetime_format)
        self._fitting_mode = False

    def set_date_range(self, min_date=None, max_date=None, datetime_format=None):
        min_date = _parse_date_input(min_date, datetime_format)
        max_date = _parse_date_input(max_date, datetime_format)
        _verify_date_bounds(min_date, max_date)
        self.min_date, self.max_date = min_date, max_date
        self._num_samples = 1 if (min_date is None or max_date is None) else int((max_date - min_date) / TEMPORAL_RESOLUTION)
        return self

    def __getitem__(self, item):
        if item < self._num_samples:
            return self.wnpy_reader[item]
        raise Exception("[ERROR] Requested item ({}) is beyond the range of valid item numbers ([0, {}]).".format(item, self._num_samples))

    def __len__(self):
        return self._num_samples

    def get_channel_count(self):
        return self.wnpy_reader.get_channel_count()

    def enable_fitting_mode(self):
        return self.set_fitting_mode(True)

    def disable_fitting_mode(self):
        return self.set_fitting_mode(False)

    def set_fitting_mode(self, mode):
        assert isinstance(mode, bool)
        self._fitting_mode = mode
        return self

    def get_fitting_mode(self):
        return self._fitting_mode

    @staticmethod
    def is_time_variate():
        return False

    def summarize(self):
        return {"data_type": "ConstantData", "path": self.wnpy_reader.path}


class WeatherBenchData(Dataset):
    """Groups ('lr', 'hr') of datasets that share their date bounds; item = tuple over groups of tuples over variables of
    ``(tensor, name, month)`` (reference :623-861)."""

    def __init__(self, min_date=None, max_date=None, datetime_format=None):
        min_date = _parse_date_input(min_date, datetime_format)
        max_date = _parse_date_input(max_date, datetime_format)
        _verify_date_bounds(min_date, max_date)
        self.min_date, self.max_date = min_date, max_date
        self.data_groups = OrderedDict({})

    def add_data_group(self, group_key, datasets, _except_on_changing_date_bounds=False):
        self._verify_data_group_inputs(group_key, datasets)
        if not isinstance(datasets, list):
            datasets = [datasets]
        mins = [d.min_date for d in datasets if d.min_date is not None]
        maxs = [d.max_date for d in datasets if d.max_date is not None]
        common_min = np.max(mins) if mins else None
        common_max = np.min(maxs) if maxs else None
        if _except_on_changing_date_bounds:
            assert common_min == self.min_date, "[ERROR] Encountered missing time stamps."
            assert common_max == self.max_date, "[ERROR] Encountered missing time stamps."
        else:
            if common_min is not None and (self.min_date is None or common_min > self.min_date):
                self.min_date = common_min
            if common_max is not None and (self.max_date is None or common_max < self.max_date):
                self.max_date = common_max
        self.data_groups.update({group_key: {d.name: d for d in datasets}})
        self._check_groups_date_bounds()
        return self

    def _check_groups_date_bounds(self):
        assert self.min_date is not None and self.max_date is not None, "[ERROR] Date bounds must be set."
        for group in self.data_groups.values():
            for d in group.values():
                assert d.min_date == self.min_date, "[ERROR] Date bounds are not same for all groups."
                assert d.max_date == self.max_date, "[ERROR] Date bounds are not same for all groups."

    def _verify_data_group_inputs(self, group_key, datasets):
        assert isinstance(group_key, str), "[ERROR] Group keys must be of type string."
        assert group_key not in self.data_groups, "[ERROR] Group keys must be unique. Key <{}> is already existing.".format(group_key)
        if not isinstance(datasets, list):
            datasets = [datasets]
        for d in datasets:
            assert isinstance(d, (ConstantData, TimeVariateData)), \
                "[ERROR] Datasets must be given as TimeVariateData or ConstantData objects or a list thereof."
        names = [d.name for d in datasets]
        assert len(names) == len(np.unique(names)), "[ERROR] Dataset names must be unique within a common parameter group."

    def remove_data_group(self, group_key):
        self.data_groups.pop(group_key, None)
        return self

    def set_date_range(self, min_date=None, max_date=None, datetime_format=None):
        min_date = _parse_date_input(min_date, datetime_format)
        max_date = _parse_date_input(max_date, datetime_format)
        _verify_date_bounds(min_date, max_date)
        self.min_date, self.max_date = min_date, max_date
        for group in self.data_groups.values():
            for d in group.values():
                d.set_date_range(min_date, max_date)
        return self

    def _first(self):
        return next(iter(next(iter(self.data_groups.values())).values()))

    def __len__(self):
        return 0 if len(self.data_groups) == 0 else len(self._first())

    def __getitem__(self, item):
        return tuple(tuple(d[item] for d in group.values()) for group in self.data_groups.values())

    def get_data_names(self):
        return {k: tuple(d.name for d in g.values()) for k, g in self.data_groups.items()}

    def get_named_item(self, item):
        return {k: {d.name: d[item] for d in g.values()} for k, g in self.data_groups.items()}

    def get_channel_count(self, group_key=None):
        if group_key is None:
            return {k: self.get_channel_count(group_key=k) for k in self.data_groups}
        if group_key in self.data_groups:
            return np.sum([d.get_channel_count() for d in self.data_groups[group_key].values()])
        raise Exception("[ERROR] Dataset does not contain a data group named <{}>.".format(group_key))

    def get_valid_time_stamps(self):
        return np.arange(self.min_date, self.max_date, TEMPORAL_RESOLUTION)

    def get_data_by_date(self, date):
        date = str_to_date(date)
        assert self.min_date <= date <= self.max_date, \
            "[ERROR] Requested date is beyond the range of valid dates. Use a date between {} and {} for this validation dataset configuration.".format(self.min_date, self.max_date)
        date = np.datetime64(date)
        return tuple(tuple(d[date] for d in group.values()) for group in self.data_groups.values())

    def summarize(self):
        return {"data_type": "WeatherBenchData",
                "date_range": [_numpy_date_to_datetime(self.min_date).strftime(DATETIME_FORMAT),
                               _numpy_date_to_datetime(self.max_date).strftime(DATETIME_FORMAT)],
                "data_groups": {k: {d.name: d.summarize() for d in g.values()} for k, g in self.data_groups.items()}}
