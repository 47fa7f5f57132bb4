"""Test-side binding of the CPU oracle (oracle/_build/libmv_oracle.so). Never imported by the product."""
import ctypes as C
import os

import numpy as np

from multivolumes_b200._abi import Binding, CasterBase, P, f32, u32, _vp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "_build", "libmv_oracle.so")

_EXTRA = {
    "sample_volume": (None, [_vp, u32, P(f32), P(f32)]),
    "sample_lightmap": (None, [_vp, u32, P(f32), P(f32)]),
    "read_per_frame": (None, [_vp, P(f32)]),
    "debug_f32": (C.c_int, [_vp, C.c_int, _vp, _vp]),
    "debug_oit": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "quantize_r11": (f32, [f32]),
    "quantize_b10": (f32, [f32]),
    "f32_to_f16": (C.c_uint16, [f32]),
    "f16_to_f32": (f32, [C.c_uint16]),
    "eval_sh_irradiance": (None, [_vp, P(f32), P(f32)]),
    "write_cubemap": (C.c_int, [_vp, u32, u32, _vp, _vp]),
    "write_lightmap_slab": (C.c_int, [_vp, u32, u32, u32, _vp]),
    "write_rows": (C.c_int, [_vp, u32, u32, u32, _vp]),
    "set_shard_volumes": (C.c_int, [_vp, u32, u32, u32]),
    "set_min16_consts_as_half": (None, [C.c_int]),
    "cube_resolve_texel": (None, [C.c_int, C.c_int, C.c_int, C.c_int, P(C.c_int)]),
}
_binding = None


def oracle_binding():
    global _binding
    if _binding is None:
        _binding = Binding(ORACLE_SO, "mvo_", _EXTRA)
        assert not _binding.missing, _binding.missing
    return _binding


class OracleCaster(CasterBase):
    """CPU oracle behind the MultiRayCaster surface. filter_model: 1 = sm_100a texture-unit model, 0 = exact fp32."""

    def __init__(self, filter_model=1, threads=0, **kw):
        super().__init__(oracle_binding(), opt0=filter_model, opt1=threads, **kw)

    def SetShardVolumes(self, rank, world, proxy_grid):
        """Volume-sharded storage as rank `rank` of `world` sees it (call after the volumes are loaded)."""
        self._ck(self.b.set_shard_volumes(self.h, rank, world, proxy_grid), "set_shard_volumes")

    def ReadPerFrame(self):
        out = np.zeros(37, np.float32)
        self.b.read_per_frame(self.h, out.ctypes.data_as(P(f32)))
        return dict(eye=out[0:3], viewport=out[3:5], screen_to_world=out[5:21].reshape(4, 4), shadow_view_proj=out[21:37].reshape(4, 4))

    def DebugF32(self, on=True):
        """keep (on) / read back the marches' fp32 outputs before the RGBA16F / R11G11B10F stores: (cube (N, 6, G, G, 4), light (L, L, L, 3))"""
        cube = np.zeros((self.N, 6, self.G, self.G, 4), np.float32); light = np.zeros((self.L,) * 3 + (3,), np.float32)
        self._ck(self.b.debug_f32(self.h, 1 if on else 0, cube.ctypes.data, light.ctypes.data), "debug_f32")
        return cube, light

    def DebugOIT(self):
        """per-pixel fragments of the resolve: (count (H, W), info (H, W, 8, 4) u32 {depth key, volume, face, stored}, data (H, W, 8, 9) f32
        {lpt xyz, face uv, colour rgba}, result (H, W, 4) f32 before the render-target blend). Leaves the colour target as it was."""
        cnt = np.zeros((self.H, self.W), np.uint32); info = np.zeros((self.H, self.W, 8, 4), np.uint32)
        data = np.zeros((self.H, self.W, 8, 9), np.float32); res = np.zeros((self.H, self.W, 4), np.float32)
        self.all_keys = np.zeros((self.H, self.W, self.N), np.uint32)      # every fragment's depth key, draw order (0xffffffff = none)
        self._ck(self.b.debug_oit(self.h, cnt.ctypes.data, info.ctypes.data, data.ctypes.data, res.ctypes.data, self.all_keys.ctypes.data), "debug_oit")
        return cnt, info, data, res

    def SampleVolume(self, src, uvw):
        uvw = np.ascontiguousarray(uvw, np.float32).reshape(-1, 3)
        out = np.empty((len(uvw), 4), np.float32)
        for i in range(len(uvw)):
            self.b.sample_volume(self.h, src, uvw[i].ctypes.data_as(P(f32)), out[i].ctypes.data_as(P(f32)))
        return out
