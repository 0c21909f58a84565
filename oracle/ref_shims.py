"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Analysis shims that let the REAL reference (``/root/reference``, read-only, present only in the build container)
be imported on a CPU-only machine without its missing third-party packages.  Used by ``make_golden.py`` and by the
"oracle vs live reference" tests (which skip when ``/root/reference`` is absent, e.g. on the GPU box).

1. ``pytorch_wavelets.DWTForward(J, wave='haar', mode=...)`` stand-in (package pinned at 1.3.0 in the reference's
   requirements.txt:6, not installable here).  Haar analysis on 2x2 blocks ``a b / c d``:
       LL = (a+b+c+d)/2, yh[:,:,0] = LH = (a+b-c-d)/2, yh[:,:,1] = HL = (a-b+c-d)/2, yh[:,:,2] = HH = (a-b-c+d)/2
   (pywt 'haar': dec_lo = [1,1]/sqrt2, dec_hi = [-1,1]/sqrt2; pytorch_wavelets correlates with the reversed filters,
   so high-pass = (even - odd)/sqrt2; rows then columns; band order LH, HL, HH).  This convention is restated from
   the published package, NOT verified against it -- the one un-verifiable assumption of the oracle.
   Known answer: [[1,2],[3,4]] -> LL 5, LH -2, HL -1, HH 0.
2. ``matplotlib.pyplot`` stub (imported, unused, at resdiff/fd_info_spliter.py:105).
3. CPU only: ``nn.Module.cuda()`` / ``.to('cuda')`` become no-ops (resdiff/unet.py:129,
   guided_cross_attention.py:19).
"""
import os
import sys
import types

import torch
from torch import nn

REFERENCE_ROOT = os.environ.get("WSR_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "models", "diffusion_models"))


class _HaarDWTForward(nn.Module):
    def __init__(self, J=1, wave="haar", mode="symmetric"):
        super().__init__()
        assert wave == "haar"
        self.J = J

    def forward(self, x):
        highs = []
        ll = x
        for _ in range(self.J):
            a = ll[:, :, 0::2, 0::2]
            b = ll[:, :, 0::2, 1::2]
            c = ll[:, :, 1::2, 0::2]
            d = ll[:, :, 1::2, 1::2]
            highs.append(torch.stack([(a + b - c - d) / 2, (a - b + c - d) / 2, (a - b - c + d) / 2], dim=2))
            ll = (a + b + c + d) / 2
        return ll, highs


_installed = False


def install():
    """Idempotently install the shims and put the reference on sys.path."""
    global _installed
    if _installed:
        return
    if not reference_available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    pw = types.ModuleType("pytorch_wavelets")
    pw.DWTForward = _HaarDWTForward
    sys.modules.setdefault("pytorch_wavelets", pw)
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib.pyplot  # noqa: F401
        except Exception:
            mpl = types.ModuleType("matplotlib")
            plt = types.ModuleType("matplotlib.pyplot")
            mpl.pyplot = plt
            sys.modules["matplotlib"] = mpl
            sys.modules["matplotlib.pyplot"] = plt
    if not torch.cuda.is_available():
        nn.Module.cuda = lambda self, device=None: self
        _orig_to = nn.Module.to

        def _to(self, *args, **kwargs):
            args = tuple(a for a in args if not (isinstance(a, str) and a.startswith("cuda")))
            if isinstance(kwargs.get("device"), str) and kwargs["device"].startswith("cuda"):
                kwargs.pop("device")
            if not args and not kwargs:
                return self
            return _orig_to(self, *args, **kwargs)

        nn.Module.to = _to
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _installed = True


def import_reference_edge():
    """The reference's collate interpolation call, StandardScaling classes and error metrics.  Their modules
    import packages that are not installed here (torcheval, skimage, intervaltree) for code paths that are NOT used by these
    classes: empty stand-ins are registered so that the imports succeed."""
    install()
    for name, attrs in {"torcheval": [], "torcheval.metrics": ["PeakSignalNoiseRatio", "MeanSquaredError", "StructuralSimilarity"],
                        "skimage": [], "skimage.metrics": ["structural_similarity"], "intervaltree": ["IntervalTree", "Interval"]}.items():
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                m = types.ModuleType(name)
                for a in attrs:
                    setattr(m, a, type(a, (), {"__init__": lambda self, *a, **k: None}))
                sys.modules[name] = m
    ns = types.SimpleNamespace()
    import training.metrics as metrics
    import data.transforms as transforms
    ns.metrics, ns.transforms = metrics, transforms
    return ns


def import_reference():
    """Returns a namespace with the reference classes on the hot path."""
    install()
    ns = types.SimpleNamespace()
    from models.diffusion_models.resdiff.unet import UNet as ResDiffUNet
    from models.diffusion_models.resdiff.resdiff_diffusion import ResDiffDiffusion
    from models.diffusion_models.srdiff.unet import UNet as SRDiffUNet
    from models.diffusion_models.srdiff.srdiff_diffusion import SRDiffDiffusion
    from models.diffusion_models.sr3.unet import UNet as SR3UNet
    from models.diffusion_models.sr3.sr3_diffusion import SR3Diffusion
    from models.diffusion_models.phydiff.unet import UNet as PhyDiffUNet
    from models.diffusion_models.phydiff.phydiff_diffusion import PhyDiffDiffusion
    from models.rrdb_encoder.RRDBNet import RRDBNet
    from models.simple_cnn.Simple_CNN import SimpleCNN
    from models.diffusion_models import networks
    from models.diffusion_models.sheduler import make_beta_schedule
    ns.ResDiffUNet, ns.ResDiffDiffusion = ResDiffUNet, ResDiffDiffusion
    ns.SRDiffUNet, ns.SRDiffDiffusion = SRDiffUNet, SRDiffDiffusion
    ns.SR3UNet, ns.SR3Diffusion, ns.PhyDiffUNet, ns.PhyDiffDiffusion = SR3UNet, SR3Diffusion, PhyDiffUNet, PhyDiffDiffusion
    ns.RRDBNet, ns.SimpleCNN, ns.networks, ns.make_beta_schedule = RRDBNet, SimpleCNN, networks, make_beta_schedule
    return ns


class _Interval(tuple):
    """(begin, end, data) with attribute access, ordered like intervaltree.Interval."""

    def __new__(cls, begin, end, data=None):
        return super().__new__(cls, (begin, end, data))

    begin = property(lambda self: self[0])
    end = property(lambda self: self[1])
    data = property(lambda self: self[2])


class _IntervalTree:
    """The subset of ``intervaltree.IntervalTree`` (3.1.0, requirements.txt) the reference's data/datasets.py uses: slice
    assignment ``tree[a:b] = data``, overlap query ``tree[a:b]`` (half-open intervals: overlap iff begin < b and end > a),
    iteration, ``len``, ``items()``."""

    def __init__(self):
        self._items = set()

    def __setitem__(self, index, data):
        self._items.add(_Interval(index.start, index.stop, data))

    def __getitem__(self, index):
        if isinstance(index, slice):
            return {iv for iv in self._items if iv.begin < index.stop and iv.end > index.start}
        return {iv for iv in self._items if iv.begin <= index < iv.end}

    def __iter__(self):
        return iter(self._items)

    def __len__(self):
        return len(self._items)

    def items(self):
        return set(self._items)


def import_reference_data():
    """The reference's data pipeline (npy_reader, datasets, transforms, dataset_builder) for SURVEY 8f N3 fixtures.  The one
    missing dependency that its code path really uses is ``intervaltree`` (a functional stand-in is installed above)."""
    install()
    if not isinstance(getattr(sys.modules.get("intervaltree"), "IntervalTree", None), type) or \
            not hasattr(sys.modules["intervaltree"].IntervalTree, "items"):
        try:
            import intervaltree  # noqa: F401
            real = hasattr(intervaltree.IntervalTree, "items")
        except Exception:
            real = False
        if not real:
            m = types.ModuleType("intervaltree")
            m.IntervalTree, m.Interval = _IntervalTree, _Interval
            sys.modules["intervaltree"] = m
            for name in ("data.datasets", "data.dataset_builder", "data.transforms"):
                sys.modules.pop(name, None)          # re-import against the functional stand-in
    ns = types.SimpleNamespace()
    import data.npy_reader as npy_reader
    import data.datasets as datasets
    import data.transforms as transforms
    import data.dataset_builder as dataset_builder
    ns.npy_reader, ns.datasets, ns.transforms, ns.dataset_builder = npy_reader, datasets, transforms, dataset_builder
    return ns
