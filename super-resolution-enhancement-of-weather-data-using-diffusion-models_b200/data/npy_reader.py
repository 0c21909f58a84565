"""Reader of the reference's on-disk store (reference data/npy_reader.py:22-277, written by
data/conversions/netcdf_to_npy.py:134-246):

    <root>/<lr|hr>/<variable>/meta/metadata.json                      name, time_variate, dims, shape, coords, attrs
    <root>/<lr|hr>/<variable>/samples/<year>/<YYYY-MM-DD-HH>.npy       one field per time stamp (time-variate variables)
    <root>/<lr|hr>/<variable>/samples/constant.npy                    constant variables

``WNPYReader`` keeps the reference's interface (``reader[i]`` / ``reader[np.datetime64]`` -> (1, C, H, W) tensor,
``get_valid_time_stamps``, ``get_channel_count``, ``meta_data``, ``name``) and the same directory / completeness checks.

What is new is the batch path the device loader uses: every sample file of a variable has the same header, so the payload
offset, dtype and shape are parsed ONCE and ``read_into`` then copies payloads straight from the files into a caller-owned
(pinned) buffer with ``readinto`` -- no per-sample ``np.load`` / ``torch.tensor`` / ``torch.cat`` allocations -- optionally
from a thread pool (file reads release the GIL)."""
import json
import os
from datetime import datetime

import numpy as np
import torch

from ..configs.config import DataConfig

config = DataConfig()
DATETIME_FORMAT = config.datetime_format
TEMPORAL_RESOLUTION = np.timedelta64(config.temporal_resolution_value, config.temporal_resolution_unit)
DIRECTORY_NAME_META_DATA = config.directory_name_meta_data
FILE_NAME_META_DATA = config.file_name_meta_data
DIRECTORY_NAME_SAMPLE_DATA = config.directory_name_sample_data
FILE_NAME_CONSTANT_DATA = config.file_name_constant_data


class WNPYReader(object):
    def __init__(self, path, domain_dimension=2, sample_index=None):
        self._verify_path(path)
        self.path = os.path.abspath(path)
        self.domain_dimension = domain_dimension
        self.meta_data = None
        self._load_meta_data()
        assert len(self.meta_data["dims"]) >= domain_dimension
        self.name = self.meta_data["name"]
        self._is_time_variate = self.meta_data["time_variate"]
        self._samples = None
        self._layout = None
        if sample_index is None:
            self._read_sample_directory()
        else:
            self._samples = sample_index

    # ---- directory structure ---------------------------------------------------------------------------------------
    @staticmethod
    def _verify_path(path):
        assert os.path.isdir(path), "[ERROR] <{}> is not a valid directory path.".format(path)
        contents = os.listdir(path)
        assert (len(contents) == 2 and os.path.isdir(os.path.join(path, DIRECTORY_NAME_META_DATA))
                and os.path.isdir(os.path.join(path, DIRECTORY_NAME_SAMPLE_DATA))), \
            "[ERROR] <{}> does not follow the expected folder structure of a WeatherBench parameter directory.".format(path)

    def _load_meta_data(self):
        with open(os.path.join(self.path, DIRECTORY_NAME_META_DATA, FILE_NAME_META_DATA + ".json"), "r") as fh:
            self.meta_data = json.load(fh)
        for c in self.meta_data["coords"]:
            c.update({"values": np.array(c["values"])})

    def _read_sample_directory(self):
        sample_directory = os.path.join(self.path, DIRECTORY_NAME_SAMPLE_DATA)
        if self._is_time_variate:
            self._verify_data_completeness(self._build_sample_index(sample_directory))
        else:
            self._load_constant_data(sample_directory)

    def _build_sample_index(self, sample_directory):
        """(first time stamp, paths sorted by time stamp); returns the sorted time stamps."""
        paths, stamps = [], []
        with os.scandir(sample_directory) as years:
            year_dirs = sorted(e.path for e in years if e.is_dir())
        for d in year_dirs:
            with os.scandir(d) as files:
                for e in files:
                    t = self._sample_time_stamp(e.name)
                    if t is not None:
                        paths.append(e.path)
                        stamps.append(t)
        assert paths, "[ERROR] <{}> contains no sample files.".format(sample_directory)
        stamps = np.array(stamps, dtype="datetime64[us]")
        order = np.argsort(stamps, kind="stable")
        self._samples = (stamps[order[0]], np.array(paths)[order])
        return stamps[order]

    @staticmethod
    def _verify_data_completeness(sample_time_stamps):
        first, last = sample_time_stamps[0], sample_time_stamps[-1]
        assert len(sample_time_stamps) == int((last - first) / TEMPORAL_RESOLUTION) + 1, "[ERROR] encountered missing data values."
        assert np.all(np.diff(sample_time_stamps) == TEMPORAL_RESOLUTION)

    @staticmethod
    def _sample_time_stamp(file_name):
        """datetime64 of ``<DATETIME_FORMAT>.npy`` or None when the name does not follow the convention."""
        parts = file_name.split(".")
        if len(parts) != 2 or parts[1] != "npy":
            return None
        try:
            return np.datetime64(datetime.strptime(parts[0], DATETIME_FORMAT))
        except ValueError:
            return None

    def _matches_sample_file_convention(self, f):
        return self._sample_time_stamp(f) is not None

    @staticmethod
    def _file_name_to_datetime(f):
        return np.datetime64(datetime.strptime(f.split(".")[0], DATETIME_FORMAT))

    def _load_constant_data(self, sample_directory):
        data = torch.tensor(np.load(os.path.join(sample_directory, FILE_NAME_CONSTANT_DATA + ".npy")))
        self._samples = self._to_pytorch_standard_shape(data)

    # ---- the reference's per-sample interface ----------------------------------------------------------------------
    def _to_pytorch_standard_shape(self, data):
        """(..., H, W) -> (1, C, H, W): leading axes are flattened into channels (reference :196-217)."""
        dim, dd = data.dim(), self.domain_dimension
        if dim == dd:
            data = data.unsqueeze(0)
        elif dim > dd + 1:
            data = torch.flatten(data, start_dim=0, end_dim=-(dd + 1))
        return data.unsqueeze(0)

    def __len__(self):
        return len(self._samples[1]) if self._is_time_variate else 1

    def index_of(self, item):
        """Position of a time stamp (or the int itself) in the sorted sample list."""
        if isinstance(item, (int, np.integer)):
            return int(item)
        return int((item - self._samples[0]) / TEMPORAL_RESOLUTION)

    def __getitem__(self, item):
        if not self._is_time_variate:
            return self._samples
        return self._to_pytorch_standard_shape(torch.tensor(np.load(self._samples[1][self.index_of(item)])))

    def get_valid_time_stamps(self):
        if not self._is_time_variate:
            return None
        first = self._samples[0]
        return np.arange(first, first + len(self._samples[1]) * TEMPORAL_RESOLUTION, TEMPORAL_RESOLUTION)

    def is_time_variate(self):
        return self._is_time_variate

    def get_channel_count(self):
        count = 1
        for axis_length in self.meta_data["shape"][0:-self.domain_dimension]:
            count *= axis_length
        return int(count)

    # ---- batch path: payloads straight into a caller-owned buffer --------------------------------------------------
    def payload_layout(self):
        """(byte offset of the array data, dtype, shape) shared by all sample files, parsed from the first one."""
        if self._layout is None:
            assert self._is_time_variate
            with open(self._samples[1][0], "rb") as fh:
                major, minor = np.lib.format.read_magic(fh)
                if (major, minor) == (1, 0):
                    shape, fortran, dtype = np.lib.format.read_array_header_1_0(fh)
                else:
                    shape, fortran, dtype = np.lib.format.read_array_header_2_0(fh)
                assert not fortran and not dtype.hasobject, "[ERROR] unsupported sample file layout."
                self._layout = (fh.tell(), dtype, tuple(shape))
        return self._layout

    def sample_shape(self):
        """(C, H, W) of one sample in the standard shape."""
        shape = self.payload_layout()[2]
        dd = self.domain_dimension
        lead = int(np.prod(shape[:-dd], dtype=np.int64)) if len(shape) > dd else 1
        return (lead,) + tuple(shape[-dd:])

    def read_into(self, dst, items, pool=None):
        """dst: float32 ndarray (n, C, H, W) with contiguous rows (e.g. the numpy view of a pinned tensor); items: n ints or
        datetime64 time stamps.  Row i receives sample items[i].  float32 stores are read with ``readinto`` (zero copies on
        the Python side); other dtypes go through one conversion."""
        offset, dtype, shape = self.payload_layout()
        n = len(items)
        assert dst.shape == (n,) + self.sample_shape() and dst.dtype == np.float32
        count = int(np.prod(shape, dtype=np.int64))
        nbytes = count * dtype.itemsize
        direct = dtype == np.dtype("<f4")
        paths = self._samples[1]
        # each ROW must be contiguous (dst may be a channel slice of a wider (n, C_total, H, W) staging buffer)
        assert n == 0 or dst[0].flags.c_contiguous, "[ERROR] rows of dst must be contiguous."
        flat = [dst[i].reshape(count) for i in range(n)]

        def one(i):
            p = paths[self.index_of(items[i])]
            with open(p, "rb", buffering=0) as fh:
                fh.seek(offset)
                if direct:
                    got = fh.readinto(memoryview(flat[i]).cast("B"))
                    assert got == nbytes, "[ERROR] truncated sample file <{}>.".format(p)
                else:
                    raw = fh.read(nbytes)
                    assert len(raw) == nbytes, "[ERROR] truncated sample file <{}>.".format(p)
                    flat[i] = np.frombuffer(raw, dtype=dtype, count=count)

        if pool is None or n < 4:
            for i in range(n):
                one(i)
        else:
            list(pool.map(one, range(n)))
        return dst
