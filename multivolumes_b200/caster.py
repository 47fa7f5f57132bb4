"""Product binding: the MultiRayCaster operator surface over libmv_b200.so (CUDA, sm_100a).

There is no CPU path. Importing this module without the built library, or creating a caster without a
B200-class device, raises — it never falls back to anything else.
"""
import ctypes as C
import os

import numpy as np

from ._abi import Binding, CasterBase, P, f32, u32, _vp

HERE = os.path.dirname(os.path.abspath(__file__))
# MV_B200_LIB lets the tuning scripts under tools/ load an experimental build of the same library
LIB_PATH = os.environ.get("MV_B200_LIB") or os.path.join(HERE, "libmv_b200.so")

FLAG_COUNT_SAMPLES = 1
FLAG_TIME_PASSES = 2
FLAG_DENSITY_ONLY = 4   # MV_FLAG_DENSITY_ONLY: R16F density volumes, colour (1, 1, 1)


class Timings(C.Structure):
    _fields_ = [(k, f32) for k in ("cull", "ray_march_light", "ray_march_view", "resolve_oit", "postprocess", "total")]

    def as_dict(self):
        return {k: float(getattr(self, k)) for k, _ in self._fields_}


class ExchangeLayout(C.Structure):
    _fields_ = [(k, C.c_uint64) for k in ("block_bytes", "arena_offset", "arena_bytes", "light_staging_offset", "light_staging_bytes",
                                          "back_buffer_offset", "back_buffer_bytes", "flags_offset", "flags_bytes")] + \
               [("light_slab_depth", u32), ("reserved", u32), ("light_staging2_offset", C.c_uint64),
                ("history_offset", C.c_uint64 * 2), ("history_bytes", C.c_uint64)]


class DdsInfo(C.Structure):
    _fields_ = [(k, u32) for k in ("width", "height", "depth", "format", "bytes_per_texel", "data_offset")]


u64p = P(C.c_uint64)
_EXTRA = {
    "last_error": (C.c_char_p, []),
    "abi_version": (u32, []),
    "set_targets_device": (C.c_int, [_vp, _vp, _vp, u32, _vp, _vp]),
    "set_volume_world_matrices": (C.c_int, [_vp, u32, u32, _vp]),
    "create_sharded": (C.c_int, [_vp, u32, u32, u32, P(_vp)]),
    "set_peer_block": (C.c_int, [_vp, u32, _vp]),
    "get_timings": (C.c_int, [_vp, P(Timings)]),
    "sync": (C.c_int, [_vp]),
    "set_flags": (C.c_int, [_vp, u32]),
    "host_alloc": (_vp, [C.c_size_t]),
    "host_free": (None, [_vp]),
    "set_row_stripes": (C.c_int, [_vp, u32]),
    "exchange_block": (C.c_int, [_vp, P(_vp), u64p]),
    "exchange_layout_get": (C.c_int, [_vp, P(ExchangeLayout)]),
    "cube_region": (C.c_int, [_vp, u32, u32, u64p, u64p, u64p, u64p]),
    "ipc_export": (C.c_int, [_vp, _vp]),
    "ipc_import": (C.c_int, [_vp, u32, _vp]),
    "peer_barrier": (C.c_int, [_vp]),
    "light_commit": (C.c_int, [_vp]),
    "set_stream": (C.c_int, [_vp, _vp]),
    "get_stream": (C.c_int, [_vp, P(_vp)]),
    "frame_buffers": (C.c_int, [_vp, P(_vp), P(_vp), P(_vp)]),
    "dds_parse": (C.c_int, [C.c_char_p, P(DdsInfo)]),
    "volume_load_dds": (C.c_int, [_vp, u32, C.c_char_p]),
    "obj_parse": (C.c_int, [C.c_char_p, P(P(f32)), P(u32), P(P(u32)), P(u32)]),
    "obj_free": (None, [P(f32), P(u32)]),
    "mesh_load_obj": (C.c_int, [_vp, C.c_char_p]),
    "write_png": (C.c_int, [C.c_char_p, _vp, u32, u32]),
    "screenshot": (C.c_int, [_vp, C.c_char_p]),
    "present_async": (C.c_int, [_vp, _vp, u32]),
    "present_rows_async": (C.c_int, [_vp, _vp, u32]),
    "host_register": (C.c_int, [_vp, C.c_size_t]),
    "host_unregister": (C.c_int, [_vp]),
    "present_wait": (C.c_int, [_vp, u32]),
}

_binding = None


def binding():
    """Loads libmv_b200.so (built by `make -C multivolumes_b200/csrc` or __graft_entry__.build())."""
    global _binding
    if _binding is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `make -C multivolumes_b200/csrc` "
                               "(the product has no CPU fallback)")
        b = Binding(LIB_PATH, "mv_", _EXTRA)
        if b.missing:
            raise RuntimeError(f"libmv_b200.so does not export: {b.missing}")
        _binding = b
    return _binding


class PinnedBuffer:
    """Page-locked host memory from mv_host_alloc, exposed as a numpy array."""

    def __init__(self, shape, dtype):
        self.b = binding()
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self.ptr = self.b.host_alloc(n)
        if not self.ptr:
            raise MemoryError(self.b.last_error().decode())
        buf = (C.c_char * n).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=dtype).reshape(shape)

    def free(self):
        if self.ptr:
            self.array = None
            self.b.host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def parse_obj(path):
    """XUSGObjLoader::Import(forDX = true) through mv_obj_parse (host only, no device needed): z negated, index list reversed."""
    b = binding()
    pos, idx, nv, ni = P(f32)(), P(u32)(), u32(0), u32(0)
    rc = b.obj_parse(os.fsencode(path), C.byref(pos), C.byref(nv), C.byref(idx), C.byref(ni))
    if rc != 0:
        raise RuntimeError(f"mv_obj_parse failed (rc={rc}): {b.last_error().decode()}")
    try:
        positions = np.ctypeslib.as_array(pos, shape=(nv.value * 3,)).copy().reshape(-1, 3) if nv.value else np.zeros((0, 3), np.float32)
        indices = np.ctypeslib.as_array(idx, shape=(ni.value,)).copy() if ni.value else np.zeros((0,), np.uint32)
    finally:
        b.obj_free(pos, idx)
    return positions, indices


def write_png(path, rgba8):
    """mv_write_png (host only): an (H, W, 4) uint8 image as a PNG file."""
    b = binding()
    img = np.ascontiguousarray(rgba8, np.uint8)
    assert img.ndim == 3 and img.shape[2] == 4
    rc = b.write_png(os.fsencode(path), img.ctypes.data, img.shape[1], img.shape[0])
    if rc != 0:
        raise RuntimeError(f"mv_write_png failed (rc={rc}): {b.last_error().decode()}")


def parse_dds(path):
    """Header of a 3-D scalar DDS through mv_dds_parse (host only): dict of width, height, depth, format, bytes_per_texel, data_offset."""
    b = binding()
    info = DdsInfo()
    rc = b.dds_parse(os.fsencode(path), C.byref(info))
    if rc != 0:
        raise RuntimeError(f"mv_dds_parse failed (rc={rc}): {b.last_error().decode()}")
    return {k: int(getattr(info, k)) for k, _ in DdsInfo._fields_}


class MultiRayCaster(CasterBase):
    """MultiVolumes/Content/MultiRayCaster.h:28-50 on one B200. Method names follow the reference class."""

    def __init__(self, device=0, count_samples=True, time_passes=False, density_only=False, shard_volumes=None, **kw):
        """shard_volumes = (rank, world, proxy_grid): volume-sharded storage (mv_create_sharded, include/mv.h)."""
        flags = (FLAG_COUNT_SAMPLES if count_samples else 0) | (FLAG_TIME_PASSES if time_passes else 0) | \
                (FLAG_DENSITY_ONLY if density_only else 0)
        b = binding()
        create = None
        if shard_volumes is not None:
            rank, world, proxy = shard_volumes
            create = lambda d, h: b.create_sharded(d, rank, world, proxy, h)
        super().__init__(b, opt0=device, opt1=flags, create=create, **kw)
        self.shard_volumes = shard_volumes
        self.device = device
        self.density_only = bool(density_only)

    # --- product-only calls ---
    def SetRenderTargetsDevice(self, depth=0, shadow=0, shadow_size=0, color=0, velocity=0):
        self._ck(self.b.set_targets_device(self.h, depth or None, shadow or None, shadow_size, color or None, velocity or None),
                 "set_targets_device")

    def SetInstrumentation(self, count_samples, time_passes):
        self._ck(self.b.set_flags(self.h, (FLAG_COUNT_SAMPLES if count_samples else 0) | (FLAG_TIME_PASSES if time_passes else 0)), "set_flags")

    def Sync(self):
        self._ck(self.b.sync(self.h), "sync")

    def GetTimings(self):
        t = Timings()
        self._ck(self.b.get_timings(self.h, C.byref(t)), "get_timings")
        return t.as_dict()

    def ReadPostInto(self, rgba8_ptr=None, taa_ptr=None):
        """Read-back into caller-owned (pinned) memory; pointers are integers."""
        self._ck(self.b.read_post(self.h, taa_ptr, rgba8_ptr), "read_post")

    def LoadVolumeFile(self, i, path):
        """MultiRayCaster::LoadVolumeData(cmdList, i, fileName): DDS import + CSR32FToRGBA16F."""
        self._ck(self.b.volume_load_dds(self.h, i, os.fsencode(path)), "volume_load_dds")

    def LoadMeshObj(self, path):
        self._ck(self.b.mesh_load_obj(self.h, os.fsencode(path)), "mesh_load_obj")

    def Screenshot(self, path):
        """MultiVolumes::SaveImage: the RGBA8 back buffer as a PNG file."""
        self._ck(self.b.screenshot(self.h, os.fsencode(path)), "screenshot")

    def PresentAsync(self, rgba8_ptr, slot):
        """Swap-chain Present: asynchronous read-back of the back buffer into pinned memory (slot < 3 in flight)."""
        self._ck(self.b.present_async(self.h, rgba8_ptr, slot), "present_async")

    def PresentRowsAsync(self, frame_ptr, slot):
        """Present of a sharded frame: this rank's rows into the whole-frame host buffer (shared by the ranks)."""
        self._ck(self.b.present_rows_async(self.h, frame_ptr, slot), "present_rows_async")

    def PresentWait(self, slot):
        self._ck(self.b.present_wait(self.h, slot), "present_wait")

    # --- multi-GPU ---
    def SetRowStripes(self, stripe_height):
        self._ck(self.b.set_row_stripes(self.h, stripe_height), "set_row_stripes")

    def ExchangeBlock(self):
        p, n = _vp(), C.c_uint64()
        self._ck(self.b.exchange_block(self.h, C.byref(p), C.byref(n)), "exchange_block")
        return p.value, n.value

    def ExchangeLayout(self):
        lay = ExchangeLayout()
        self._ck(self.b.exchange_layout_get(self.h, C.byref(lay)), "exchange_layout_get")
        return lay

    def CubeRegion(self, volume, mip):
        a, b_, c_, d = (C.c_uint64() for _ in range(4))
        self._ck(self.b.cube_region(self.h, volume, mip, C.byref(a), C.byref(b_), C.byref(c_), C.byref(d)), "cube_region")
        return a.value, b_.value, c_.value, d.value

    def IpcExport(self):
        buf = C.create_string_buffer(64)
        self._ck(self.b.ipc_export(self.h, buf), "ipc_export")
        return bytes(buf.raw)

    def IpcImport(self, peer, handle):
        buf = C.create_string_buffer(bytes(handle), 64)
        self._ck(self.b.ipc_import(self.h, peer, buf), "ipc_import")

    def SetPeerBlock(self, peer, block_ptr):
        """A peer caster of THIS process (same or peer-enabled device): its exchange block is addressed directly."""
        self._ck(self.b.set_peer_block(self.h, peer, block_ptr), "set_peer_block")

    def PeerBarrier(self):
        self._ck(self.b.peer_barrier(self.h), "peer_barrier")

    def LightCommit(self):
        self._ck(self.b.light_commit(self.h), "light_commit")

    def SetStream(self, stream_ptr):
        self._ck(self.b.set_stream(self.h, stream_ptr or None), "set_stream")

    def FrameBuffers(self):
        a, b_, c_ = _vp(), _vp(), _vp()
        self._ck(self.b.frame_buffers(self.h, C.byref(a), C.byref(b_), C.byref(c_)), "frame_buffers")
        return a.value, b_.value, c_.value
