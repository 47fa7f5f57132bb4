"""multivolumes_b200 — B200-native (sm_100a) cube-map-space multi-volume ray marching.

The compute lives in libmv_b200.so (hand-written CUDA behind the C-ABI of include/mv.h); this package
is the Python mirror of the reference's MultiRayCaster operator surface plus the harness helpers
(scene set-up, multi-GPU driver). Nothing here computes on the CPU.
"""
from . import scene  # noqa: F401
from .caster import MultiRayCaster, PinnedBuffer, binding, parse_dds, parse_obj, write_png, LIB_PATH  # noqa: F401

__all__ = ["MultiRayCaster", "PinnedBuffer", "binding", "parse_dds", "parse_obj", "write_png", "scene", "LIB_PATH"]
