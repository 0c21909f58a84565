"""B200-native (sm_100a) implementation of the diffusion super-resolution hot path of
jellikus/Super-Resolution-Enhancement-of-Weather-Data-Using-Diffusion-Models.

Layout
------
csrc/        hand-written CUDA kernels + the C ABI (include/wsr.h) -> _lib/libwsr.so
_native.py   ctypes binding (no fallback: raises if the library is missing or a call fails)
engine.py    device buffers, packed weights, the per-step launch schedule and its CUDA graph
models/      host-side mirror of the reference's model classes (same names, signatures, state_dict keys)
configs/     the reference's JSON-with-comments config parser
train.py / sample.py   entry points with the reference's command-line flags
"""
from . import _native as native  # noqa: F401

__all__ = ["native"]
