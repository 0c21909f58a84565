"""``DataHandler`` (reference data/dataset_builder.py:14-382): builds the train / validation ``WeatherBenchData`` over the
on-disk store for a date range and a month subset, fits the per-month-group transforms, and hands out batch loaders whose
items are the reference's collate output ``({'HR', 'LR', 'SR'}, months)``.

The loader is where this differs from the reference (``torch.utils.data.DataLoader`` with worker processes, per-sample
``np.load`` + transform + ``torch.cat``, bicubic interpolation per sample on the CPU):

``DeviceBatchLoader``  a reader thread fills pinned staging buffers with the RAW fields of the next batches
(``WNPYReader.read_into``: file payloads straight into pinned memory, thread pool); the consumer copies a staged batch to the
device on a side stream and then needs three launches for the whole batch: standardise LR, standardise HR
(``wsr_standard_scale`` with the per-(sample, variable) statistics of each sample's month) and the bicubic x4 condition
(``wsr_bicubic_upsample``).  Reading batch i+1 and copying it overlap the model's work on batch i.  Without a CUDA device
(host-only use of the data pipeline, e.g. inspecting a store) the same batches are produced as host tensors with the
reference's formulas."""
import logging
import os
import queue
import threading
from concurrent.futures import ThreadPoolExecutor
from types import SimpleNamespace

import numpy as np
import torch
from torch.nn.functional import interpolate

from .. import _native as nat
from .datasets import TimeVariateData, WeatherBenchData
from .npy_reader import WNPYReader
from .transforms import DataTransformer, IdentityTransform, StandardScaling
from .utils import date_to_str, is_full_year, log_dataset_info, month_windows, save_object, validate_month_subset


def bicubic_sr(lr, scale_factor=4):
    """LR (B, C, h, w) fp32 CUDA tensor -> SR (B, C, h*scale, w*scale), same arithmetic as F.interpolate(mode='bicubic')."""
    if not lr.is_cuda:
        raise nat.WsrError("bicubic_sr runs on the CUDA path only (got a %s tensor)" % lr.device)
    x = lr.to(torch.float32).contiguous()
    b, c, h, w = x.shape
    s = int(scale_factor)
    if s != scale_factor or s < 1:
        raise NotImplementedError("integer scale factors only (the reference uses 4 everywhere)")
    out = torch.empty((b, c, h * s, w * s), device=x.device, dtype=torch.float32)
    nat.call("wsr_bicubic_upsample", x.data_ptr(), b * c, h, w, s, out.data_ptr(), torch.cuda.current_stream(x.device).cuda_stream)
    return out


def collate_batch(lr, hr):
    """The dict the reference's collate returns (dataset_builder.py:372-382) from already-stacked LR / HR device tensors."""
    return {"HR": hr, "LR": lr, "SR": bicubic_sr(lr, 4)}


class DeviceBatchLoader:
    """Iterable over ``({'HR', 'LR', 'SR'}, months)`` batches of a ``WeatherBenchData`` with groups 'lr' and 'hr'.

    device: a CUDA device -> device-resident batches (see the module docstring); None -> host tensors.
    drop_last is always on and the order is a fresh permutation per epoch when ``shuffle`` (reference :139-147)."""

    def __init__(self, dataset, batch_size, shuffle=False, num_workers=4, device=None, scale=4, prefetch=2, seed=0, shard=None):
        """shard = (rank, world): under one-process-per-GPU data parallelism every rank walks the SAME batch sequence (same seed) but
        reads, copies and transforms only its contiguous 1 / world slice of each global batch (``batch_size`` stays the GLOBAL size,
        as in the reference's config; what ``nn.DataParallel`` scattered after loading is never loaded here)."""
        self.dataset, self.batch_size, self.shuffle, self.scale = dataset, int(batch_size), bool(shuffle), scale
        self.shard = (0, 1) if shard is None else (int(shard[0]), int(shard[1]))
        assert 0 <= self.shard[0] < self.shard[1] and self.batch_size % self.shard[1] == 0, "global batch must divide by the world size"
        self.device = torch.device(device) if device is not None else None
        self.prefetch = max(1, int(prefetch))
        self.workers = max(1, int(num_workers or 1))
        self._rng = np.random.default_rng(seed)
        self._groups = {k: list(dataset.data_groups[k].values()) for k in ("lr", "hr")}
        self._bulk = all(self._bulk_ok(d) for g in self._groups.values() for d in g)
        if self._bulk:
            self._shapes = {}
            for k, g in self._groups.items():
                shp = [d.wnpy_reader.sample_shape() for d in g]
                assert all(s[1:] == shp[0][1:] for s in shp), "[ERROR] variables of one group must share the grid."
                self._shapes[k] = (sum(s[0] for s in shp),) + shp[0][1:]

    @staticmethod
    def _bulk_ok(d):
        """The staged path handles identity and scalar-statistics transforms; anything else goes sample by sample."""
        if not isinstance(d, TimeVariateData) or d._delays is not None:
            return False
        return all(isinstance(t, IdentityTransform) or (isinstance(t, StandardScaling) and t._mean is not None and t._mean.numel() == 1)
                   for t in d.get_transform().values())

    def __len__(self):
        return len(self.dataset) // self.batch_size

    def _order(self):
        n = len(self.dataset)
        return self._rng.permutation(n) if self.shuffle else np.arange(n)

    # ---- host side: raw fields + statistics of one batch ---------------------------------------------------------------
    def _stats(self, group, months):
        """(B, C_total) mean / std of this batch (identity = 0 / 1), one column per channel."""
        cols_m, cols_s = [], []
        for d in self._groups[group]:
            tr = d.get_transform()
            m = np.zeros(len(months), dtype=np.float32)
            s = np.ones(len(months), dtype=np.float32)
            for mo in set(months):
                t = tr.get(mo) if mo in tr else None
                if isinstance(t, StandardScaling):
                    sel = np.asarray(months) == mo
                    m[sel], s[sel] = float(t._mean), float(t._std())
            c = d.wnpy_reader.sample_shape()[0]
            cols_m += [m] * c
            cols_s += [s] * c
        return torch.from_numpy(np.stack(cols_m, 1)), torch.from_numpy(np.stack(cols_s, 1))

    def _stage(self, idx, bufs, pool):
        for k, g in self._groups.items():
            dst = bufs[k].numpy()
            c0 = 0
            for d in g:
                c = d.wnpy_reader.sample_shape()[0]
                d.wnpy_reader.read_into(dst[:, c0:c0 + c], d.stamps_of(idx), pool)
                c0 += c
        months = self._groups["lr"][0].months_of(idx)
        return months, self._stats("lr", months), self._stats("hr", months)

    def _new_bufs(self):
        pin = self.device is not None
        n = self.batch_size // self.shard[1]
        return {k: torch.empty((n,) + self._shapes[k], dtype=torch.float32, pin_memory=pin) for k in ("lr", "hr")}

    # ---- iteration ---------------------------------------------------------------------------------------------------
    def __iter__(self):
        order = self._order()
        r, w = self.shard
        per = self.batch_size // w
        batches = [order[i * self.batch_size + r * per:i * self.batch_size + (r + 1) * per] for i in range(len(self))]
        if not self._bulk:
            for idx in batches:
                yield form_batch([self.dataset[int(i)] for i in idx], self.scale, self.device)
            return
        free, ready = queue.Queue(), queue.Queue()
        for _ in range(self.prefetch + 1):
            free.put((self._new_bufs(), None))
        stop = threading.Event()

        def reader():
            try:
                with ThreadPoolExecutor(max_workers=self.workers) as pool:
                    for idx in batches:
                        bufs, ev = free.get()
                        if stop.is_set():
                            return
                        if ev is not None:
                            ev.synchronize()            # the previous copy out of these pinned buffers has finished
                        ready.put((bufs, self._stage(idx, bufs, pool)))
                ready.put(None)
            except BaseException as exc:                # surfaced in the consumer
                ready.put(exc)

        th = threading.Thread(target=reader, daemon=True)
        th.start()
        copy_stream = torch.cuda.Stream(self.device) if self.device is not None else None
        try:
            while True:
                item = ready.get()
                if item is None:
                    break
                if isinstance(item, BaseException):
                    raise item
                bufs, (months, lr_stats, hr_stats) = item
                if self.device is None:
                    lr = (bufs["lr"] - lr_stats[0][:, :, None, None]) / lr_stats[1][:, :, None, None]
                    hr = (bufs["hr"] - hr_stats[0][:, :, None, None]) / hr_stats[1][:, :, None, None]
                    free.put((bufs, None))
                    yield {"HR": hr, "LR": lr, "SR": interpolate(lr, scale_factor=self.scale, mode="bicubic")}, months
                    continue
                from .transforms import transform_batch
                cur = torch.cuda.current_stream(self.device)
                with torch.cuda.stream(copy_stream):
                    lr_raw = bufs["lr"].to(self.device, non_blocking=True)
                    hr_raw = bufs["hr"].to(self.device, non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(copy_stream)
                cur.wait_event(ev)
                lr_raw.record_stream(cur)
                hr_raw.record_stream(cur)
                free.put((bufs, ev))
                lr = transform_batch(lr_raw, *lr_stats)
                hr = transform_batch(hr_raw, *hr_stats)
                yield {"HR": hr, "LR": lr, "SR": bicubic_sr(lr, self.scale)}, months
        finally:
            stop.set()
            free.put((None, None))
            th.join(timeout=5)


def form_batch(samples, scale=4, device=None):
    """The reference's collate (``DataHandler._form_batch``, :344-382) for a list of ``WeatherBenchData`` items; with a CUDA
    ``device`` the stacked LR / HR go to the device first and the condition is interpolated there."""
    lr = torch.cat([torch.cat([v[0] for v in low], dim=1) for low, _ in samples])
    hr = torch.cat([torch.cat([v[0] for v in high], dim=1) for _, high in samples])
    months = [low[0][2] for low, _ in samples]
    if device is not None and torch.device(device).type == "cuda":
        lr, hr = lr.to(device, torch.float32), hr.to(device, torch.float32)
        return {"HR": hr, "LR": lr, "SR": bicubic_sr(lr, scale)}, months
    return {"HR": hr, "LR": lr, "SR": interpolate(lr, scale_factor=scale, mode="bicubic")}, months


class DataHandler:
    def __init__(self, dataroot, variables, storage_root, months_subset, groups, transformation, train_min_date=None,
                 train_max_date=None, val_min_date=None, val_max_date=None, val_batch_size=None, train_batch_size=None,
                 shuffle_data=True, num_workers=None, device="auto", shard=None):
        self.metadata = {}
        self.dataroot, self.variables, self.storage_root = dataroot, variables, storage_root
        self.months_subset, self.groups, self.transformation = months_subset, groups, transformation
        self.train_min_date, self.train_max_date = train_min_date, train_max_date
        self.val_min_date, self.val_max_date = val_min_date, val_max_date
        self.val_batch_size, self.train_batch_size = val_batch_size, train_batch_size
        self.shuffle_data, self.num_workers = shuffle_data, num_workers
        validate_month_subset(months_subset)
        self.data_transformer = DataTransformer(variables, dataroot, months_subset, groups)
        self.train_loader = self.val_loader = self.train_dataset = self.val_dataset = None
        if device == "auto":
            device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else None
        self.device = device
        self.shard = shard                 # (rank, world) for the TRAINING loader under torchrun; validation is not sharded
        self._readers = {}

    # ---- accessors -----------------------------------------------------------------------------------------------
    def get_datasets(self):
        return self.train_dataset, self.val_dataset

    def get_data_loaders(self):
        return self.train_loader, self.val_loader

    def get_metadata(self):
        return SimpleNamespace(**self.metadata)

    def get_data_transformer(self):
        return self.data_transformer

    def get_all(self):
        return self.train_dataset, self.val_dataset, self.get_metadata(), self.data_transformer

    def get_tensor_by_date(self, date):
        pass

    # ---- datasets ------------------------------------------------------------------------------------------------
    def create_train_set(self, train_min_date=None, train_max_date=None):
        self.train_min_date = train_min_date or self.train_min_date
        self.train_max_date = train_max_date or self.train_max_date
        self.train_dataset = self._create_set(self.train_min_date, self.train_max_date, train=True)

    def create_val_set(self, val_min_date=None, val_max_date=None, transform=None):
        self.val_min_date = val_min_date or self.val_min_date
        self.val_max_date = val_max_date or self.val_max_date
        self.val_dataset = self._create_set(self.val_min_date, self.val_max_date, train=False)
        return self.val_dataset

    def _reader(self, data_type, variable):
        """One directory scan per (lr/hr, variable): the sorted sample index is shared by every reader of that directory."""
        key = (data_type, variable)
        path = os.path.join(self.dataroot, data_type, variable)
        if key not in self._readers:
            self._readers[key] = WNPYReader(path)
            return self._readers[key]
        return WNPYReader(path, sample_index=self._readers[key]._samples)

    def _create_set(self, min_date=None, max_date=None, train=True):
        datasets = {"lr": [], "hr": []}
        for variable in self.variables:
            for data_type in ("lr", "hr"):
                reader = self._reader(data_type, variable)
                if train:
                    transform = self.data_transformer.transform(min_date, max_date, data_type, variable, self.transformation)
                    self._update_metadata(data_type, reader)
                else:
                    transform = self.data_transformer.get_transform(variable, data_type)
                name = f"{data_type}_{variable}"
                if is_full_year(self.months_subset):
                    data = TimeVariateData(reader, name=name, lead_time=0, min_date=min_date, max_date=max_date, transform=transform)
                else:
                    data = self._create_dataset_by_month_subset(reader, name, 0, min_date, max_date, transform)
                datasets[data_type].append(data)
        dataset = WeatherBenchData(min_date=min_date, max_date=max_date)
        dataset.add_data_group("lr", datasets["lr"])
        dataset.add_data_group("hr", datasets["hr"])
        return dataset

    def _create_dataset_by_month_subset(self, source, name, lead_time, min_date, max_date, transform):
        """Union of the month windows of [min_date, max_date) whose month is in ``months_subset`` (reference :296-342)."""
        dataset = None
        for start, end in month_windows(min_date, max_date):
            if start.month not in self.months_subset:
                continue
            if dataset is None:
                dataset = TimeVariateData(source, name=name, lead_time=lead_time, min_date=date_to_str(start),
                                          max_date=date_to_str(end), transform=transform)
            else:
                dataset.add_data_by_date(date_to_str(start), date_to_str(end))
        return dataset

    def _update_metadata(self, data_type, wbd_reader):
        for dimension in wbd_reader.meta_data["coords"]:
            self.metadata[f"{data_type}_{dimension['name']}"] = dimension["values"]

    def _save_metadata_and_transformations(self):
        save_object(self.metadata, self.storage_root, "metadata")
        save_object(self.data_transformer.transformation_dict, self.storage_root, "transformations")

    # ---- loaders -------------------------------------------------------------------------------------------------
    def create_train_loader(self, batch_size, use_shuffle, num_workers):
        if self.train_dataset is None:
            raise ValueError("Training dataset is not created. Call create_train_set() first.")
        self.train_loader = DeviceBatchLoader(self.train_dataset, batch_size, shuffle=use_shuffle, num_workers=num_workers, device=self.device,
                                              shard=self.shard)
        return self.train_loader

    def create_val_loader(self, batch_size, use_shuffle, num_workers):
        if self.val_dataset is None:
            raise ValueError("Validation dataset is not created. Call create_val_set() first.")
        self.val_loader = DeviceBatchLoader(self.val_dataset, batch_size, shuffle=False, num_workers=num_workers, device=self.device)

    def log_info(self):
        logger = logging.getLogger("base")
        log_dataset_info(self.train_dataset, "Train WeatherDataset", logger)
        log_dataset_info(self.val_dataset, "Validation WeatherDataset", logger)
        logger.info("Finished.\n")

    def process_data(self):
        self.create_train_set()
        self.create_val_set()
        self._save_metadata_and_transformations()
        self.create_train_loader(self.train_batch_size, self.shuffle_data, self.num_workers)
        self.create_val_loader(self.val_batch_size, self.shuffle_data, self.num_workers)
        self.log_info()
        return self.train_loader, self.val_loader, self.get_metadata(), self.get_data_transformer()

    def get_data_by_date(self, date):
        return self._form_batch([self.val_dataset.get_data_by_date(date)])

    def _form_batch(self, samples: list):
        return form_batch(samples, 4, self.device)
