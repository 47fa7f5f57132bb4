"""ctypes binding of the mv_* C-ABI (include/mv.h).

The same call surface is implemented by the product library (prefix ``mv_``, CUDA only) and, for the
tests, by the CPU oracle (prefix ``mvo_``, built from oracle/). ``Binding`` is told which shared
object and prefix to bind; the package itself only ever binds the product library.
"""
import ctypes as C
import numpy as np

u32, i32, f32 = C.c_uint32, C.c_int32, C.c_float
P = C.POINTER


class Desc(C.Structure):
    """mv_desc / mvo_desc (identical layout; opt0/opt1 = device, flags | tex_filter_model, num_threads)."""
    _fields_ = [("grid_size", u32), ("light_grid_size", u32), ("num_volumes", u32), ("num_volume_srcs", u32),
                ("width", u32), ("height", u32), ("max_ray_samples", u32), ("max_light_samples", u32),
                ("opt0", u32), ("opt1", u32)]


class Stats(C.Structure):
    _fields_ = [("view_rays", C.c_uint64), ("view_samples", C.c_uint64), ("view_light_fetches", C.c_uint64),
                ("light_voxels", C.c_uint64), ("light_dense_voxels", C.c_uint64), ("light_samples", C.c_uint64),
                ("direct_rays", C.c_uint64), ("direct_samples", C.c_uint64), ("direct_light_fetches", C.c_uint64),
                ("oit_fragments", C.c_uint64),
                ("visible_count", u32), ("cubemap_count", u32), ("light_volume", u32), ("threads", u32),
                ("view_skipped", C.c_uint64), ("direct_skipped", C.c_uint64)]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


_vp = C.c_void_p
# name -> (restype, argtypes); the handle is always the first argument
_COMMON = {
    "create": (C.c_int, [P(Desc), P(_vp)]),
    "destroy": (None, [_vp]),
    "volume_init_procedural": (C.c_int, [_vp, u32, u32, u32]),
    "volume_upload_rgba16f": (C.c_int, [_vp, u32, _vp]),
    "volume_upload_r32f": (C.c_int, [_vp, u32, _vp]),
    "volume_upload_r32f_sized": (C.c_int, [_vp, u32, _vp, u32, u32, u32]),
    "volume_read": (C.c_int, [_vp, u32, _vp]),
    "set_targets": (C.c_int, [_vp, _vp, _vp, u32, _vp, _vp]),
    "reset_color": (C.c_int, [_vp]),
    "set_sh": (C.c_int, [_vp, _vp]),
    "set_max_samples": (C.c_int, [_vp, u32, u32]),
    "set_volumes_world": (C.c_int, [_vp, f32, P(f32)]),
    "set_volume_world": (C.c_int, [_vp, u32, f32, P(f32)]),
    "set_volume_world_matrix": (C.c_int, [_vp, u32, P(f32)]),
    "set_light": (C.c_int, [_vp, P(f32), P(f32), f32]),
    "set_ambient": (C.c_int, [_vp, P(f32), f32]),
    "update_frame": (C.c_int, [_vp, P(f32), P(f32), P(f32)]),
    "render": (C.c_int, [_vp, u32]),
    "render_work_graph": (C.c_int, [_vp, u32]),
    "cull": (C.c_int, [_vp]),
    "ray_march_light": (C.c_int, [_vp, i32]),
    "ray_march_view": (C.c_int, [_vp]),
    "resolve_oit": (C.c_int, [_vp]),
    "postprocess": (C.c_int, [_vp, u32]),
    "sh_project": (C.c_int, [_vp, _vp, u32, _vp]),
    "set_environment": (C.c_int, [_vp, _vp, u32]),
    "render_environment": (C.c_int, [_vp]),
    "read_per_object": (C.c_int, [_vp, _vp]),
    "read_visible": (C.c_int, [_vp, _vp, P(u32)]),
    "read_cube_volumes": (C.c_int, [_vp, _vp, P(u32)]),
    "read_attribs": (C.c_int, [_vp, _vp]),
    "read_cubemap": (C.c_int, [_vp, u32, u32, _vp, _vp]),
    "read_lightmap": (C.c_int, [_vp, u32, _vp]),
    "read_frame": (C.c_int, [_vp, _vp]),
    "read_post": (C.c_int, [_vp, _vp, _vp]),
    "get_stats": (C.c_int, [_vp, P(Stats)]),
    "set_frame_index": (C.c_int, [_vp, u32]),
    "mesh_set": (C.c_int, [_vp, _vp, u32, _vp, u32]),
    "mesh_set_world": (C.c_int, [_vp, f32, P(f32)]),
    "mesh_render_depth": (C.c_int, [_vp, P(f32), P(f32)]),
    "mesh_render": (C.c_int, [_vp, P(f32), P(f32), P(f32), P(f32)]),
    "read_velocity": (C.c_int, [_vp, _vp]),
    "read_depth": (C.c_int, [_vp, _vp, _vp, P(u32)]),
    "set_shard": (C.c_int, [_vp, u32, u32]),
    "set_row_band": (C.c_int, [_vp, u32, u32]),
}


def _fp(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(P(f32))


class Binding:
    """Loads a shared object exporting <prefix><name> for every entry of the call surface."""

    def __init__(self, path, prefix, extra=None):
        self.lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
        self.prefix = prefix
        self.path = path
        table = dict(_COMMON)
        if extra:
            table.update(extra)
        self.missing = []
        for name, (res, args) in table.items():
            try:
                fn = getattr(self.lib, prefix + name)
            except AttributeError:
                self.missing.append(prefix + name)
                continue
            fn.restype = res
            fn.argtypes = args
            setattr(self, name, fn)


class CasterBase:
    """Shared implementation of the MultiRayCaster operator surface over a Binding.

    Method names follow the reference class (MultiVolumes/Content/MultiRayCaster.h:28-50); array
    arguments are numpy arrays in host memory.
    """

    def __init__(self, binding, grid_size=128, light_grid_size=96, num_volumes=2, num_volume_srcs=None,
                 width=1280, height=720, max_ray_samples=256, max_light_samples=96, opt0=0, opt1=0, create=None):
        self.b = binding
        self.G, self.L, self.N = grid_size, light_grid_size, num_volumes
        self.srcs = num_volume_srcs or num_volumes
        self.W, self.H = width, height
        d = Desc(grid_size, light_grid_size, num_volumes, self.srcs, width, height, max_ray_samples, max_light_samples, opt0, opt1)
        h = _vp()
        rc = create(C.byref(d), C.byref(h)) if create else binding.create(C.byref(d), C.byref(h))
        if rc != 0 or not h:
            raise RuntimeError(f"{binding.prefix}create failed (rc={rc}): {self._last_error()}")
        self.h = h

    def _last_error(self):
        fn = getattr(self.b, "last_error", None)
        return fn().decode() if fn else ""

    def _ck(self, rc, what):
        if rc != 0:
            raise RuntimeError(f"{self.b.prefix}{what} failed (rc={rc}): {self._last_error()}")

    def close(self):
        if getattr(self, "h", None):
            self.b.destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- MultiRayCaster::InitVolumeData / LoadVolumeData ---
    def InitVolumeData(self, i, mode=0, seed=0):
        self._ck(self.b.volume_init_procedural(self.h, i, mode, seed), "volume_init_procedural")

    def LoadVolumeData(self, i, data):
        data = np.ascontiguousarray(data)
        if data.dtype == np.float32 and data.ndim == 3 and data.shape != (self.G,) * 3:
            # a scalar source of another resolution: resampled by CSR32FToRGBA16F's LINEAR fetch (array is [z][y][x])
            d, h, w = data.shape
            self._ck(self.b.volume_upload_r32f_sized(self.h, i, data.ctypes.data, w, h, d), "volume_upload_r32f_sized")
        elif data.dtype == np.float32:
            assert data.size == self.G ** 3
            self._ck(self.b.volume_upload_r32f(self.h, i, data.ctypes.data), "volume_upload_r32f")
        else:
            data = data.view(np.uint16)
            assert data.size == self.G ** 3 * 4
            self._ck(self.b.volume_upload_rgba16f(self.h, i, data.ctypes.data), "volume_upload_rgba16f")

    def ReadVolume(self, i):
        out = np.empty((self.G, self.G, self.G, 4), np.uint16)
        self._ck(self.b.volume_read(self.h, i, out.ctypes.data), "volume_read")
        return out.view(np.float16)

    # --- SetRenderTargets / SetViewport ---
    def SetRenderTargets(self, depth=None, shadow=None, color=None, velocity=None):
        keep = []
        def ptr(a, dt, n):
            if a is None:
                return None
            a = np.ascontiguousarray(a).view(dt) if np.asarray(a).dtype.itemsize == np.dtype(dt).itemsize else np.ascontiguousarray(a, dtype=dt)
            assert a.size == n, (a.size, n)
            keep.append(a)
            return a.ctypes.data
        ssize = 0 if shadow is None else int(np.asarray(shadow).shape[0])
        self._ck(self.b.set_targets(self.h, ptr(depth, np.float32, self.W * self.H), ptr(shadow, np.uint16, ssize * ssize), ssize,
                                    ptr(color, np.uint16, self.W * self.H * 4), ptr(velocity, np.uint16, self.W * self.H * 2)), "set_targets")

    def ResetColor(self):
        self._ck(self.b.reset_color(self.h), "reset_color")

    def SetSH(self, coeffs):
        if coeffs is None:
            self._ck(self.b.set_sh(self.h, None), "set_sh")
        else:
            a = np.ascontiguousarray(coeffs, dtype=np.float32).reshape(27)
            self._ck(self.b.set_sh(self.h, a.ctypes.data), "set_sh")

    def SetMaxSamples(self, ray, light):
        self._ck(self.b.set_max_samples(self.h, ray, light), "set_max_samples")

    def SetVolumesWorld(self, size, center=(0, 0, 0)):
        a, p = _fp(center)
        self._ck(self.b.set_volumes_world(self.h, size, p), "set_volumes_world")

    def SetVolumeWorld(self, i, size, pos):
        a, p = _fp(pos)
        self._ck(self.b.set_volume_world(self.h, i, size, p), "set_volume_world")

    def SetVolumeWorldMatrix(self, i, world43):
        a, p = _fp(np.asarray(world43).reshape(12))
        self._ck(self.b.set_volume_world_matrix(self.h, i, p), "set_volume_world_matrix")

    def SetVolumeWorldMatrices(self, world43, first=0):
        """(count, 4, 3) matrices for consecutive volumes; one call where the library offers it."""
        m = np.ascontiguousarray(world43, np.float32).reshape(-1, 12)
        fn = getattr(self.b, "set_volume_world_matrices", None)
        if fn is not None:
            self._ck(fn(self.h, first, m.shape[0], m.ctypes.data), "set_volume_world_matrices")
        else:
            for k in range(m.shape[0]):
                self.SetVolumeWorldMatrix(first + k, m[k])

    def SetLight(self, pos, color, intensity):
        a, pa = _fp(pos); b, pb = _fp(color)
        self._ck(self.b.set_light(self.h, pa, pb, intensity), "set_light")

    def SetAmbient(self, color, intensity):
        a, pa = _fp(color)
        self._ck(self.b.set_ambient(self.h, pa, intensity), "set_ambient")

    # --- the occluder mesh: ObjectRenderer's depth-only passes (ObjectRenderer.h; .cpp:147-243, 555-570) ---
    def SetMesh(self, positions, indices):
        """positions (V, 3) float32 and indices (3 T,) uint32 of a triangle list (createVB / createIB)."""
        pos = np.ascontiguousarray(positions, np.float32).reshape(-1, 3)
        idx = np.ascontiguousarray(indices, np.uint32).reshape(-1)
        self._ck(self.b.mesh_set(self.h, pos.ctypes.data, pos.shape[0], idx.ctypes.data, idx.shape[0]), "mesh_set")

    def SetMeshWorld(self, scale, pos):
        a, pa = _fp(pos)
        self._ck(self.b.mesh_set_world(self.h, scale, pa), "mesh_set_world")

    def RenderMeshDepth(self, view_proj):
        """RenderShadow + depth pre-pass into the caster's shadow map / scene depth; returns the light's view-projection."""
        a, pa = _fp(np.asarray(view_proj).reshape(16))
        out = np.zeros(16, np.float32)
        self._ck(self.b.mesh_render_depth(self.h, pa, out.ctypes.data_as(P(f32))), "mesh_render_depth")
        return out.reshape(4, 4)

    def RenderMesh(self, view_proj, eye, clear_rgba=(0.0, 0.0, 0.0, 0.0)):
        """ObjectRenderer::UpdateFrame(viewProj, eyePt) + RenderShadow + Render: scene depth, shadow map, background colour
        and TAA velocity all come from the shaded mesh. Returns the light's view-projection (for UpdateFrame)."""
        a, pa = _fp(np.asarray(view_proj).reshape(16))
        e, pe = _fp(np.asarray(eye).reshape(3))
        cl, pcl = _fp(np.asarray(clear_rgba).reshape(4))
        out = np.zeros(16, np.float32)
        self._ck(self.b.mesh_render(self.h, pa, pe, pcl, out.ctypes.data_as(P(f32))), "mesh_render")
        return out.reshape(4, 4)

    def ReadVelocity(self):
        out = np.empty((self.H, self.W, 2), np.uint16)
        self._ck(self.b.read_velocity(self.h, out.ctypes.data), "read_velocity")
        return out

    def ReadDepth(self):
        depth = np.empty((self.H, self.W), np.float32)
        size = u32(0)
        self._ck(self.b.read_depth(self.h, None, None, C.byref(size)), "read_depth")
        shadow = np.empty((size.value, size.value), np.uint16)
        self._ck(self.b.read_depth(self.h, depth.ctypes.data, shadow.ctypes.data if size.value else None, C.byref(size)), "read_depth")
        return depth, shadow

    def UpdateFrame(self, view_proj, shadow_vp, eye):
        a, pa = _fp(np.asarray(view_proj).reshape(16))
        if shadow_vp is None:
            shadow_vp = np.eye(4)
        b, pb = _fp(np.asarray(shadow_vp).reshape(16))
        c, pc = _fp(eye)
        self._ck(self.b.update_frame(self.h, pa, pb, pc), "update_frame")

    # --- passes ---
    def Render(self, oit_method=0, use_work_graph=False):
        """MultiRayCaster::Render (MultiRayCaster.h:49-50); use_work_graph = the reference's useWorkGraph argument."""
        if use_work_graph:
            self._ck(self.b.render_work_graph(self.h, oit_method), "render_work_graph")
        else:
            self._ck(self.b.render(self.h, oit_method), "render")

    def Cull(self):
        self._ck(self.b.cull(self.h), "cull")

    def RayMarchL(self, volume=-1):
        self._ck(self.b.ray_march_light(self.h, volume), "ray_march_light")

    def RayMarchV(self):
        self._ck(self.b.ray_march_view(self.h), "ray_march_view")

    def ResolveOIT(self):
        self._ck(self.b.resolve_oit(self.h), "resolve_oit")

    def Postprocess(self, taa=True):
        self._ck(self.b.postprocess(self.h, 1 if taa else 0), "postprocess")

    def SetEnvironment(self, cube_rgb):
        """LightProbe: the radiance cube map, (6, S, S, 3) float32 in the D3D face order (None = no environment)."""
        if cube_rgb is None:
            self._ck(self.b.set_environment(self.h, None, 0), "set_environment")
            return
        cube = np.ascontiguousarray(cube_rgb, dtype=np.float32)
        assert cube.ndim == 4 and cube.shape[0] == 6 and cube.shape[1] == cube.shape[2] and cube.shape[3] == 3
        self._ck(self.b.set_environment(self.h, cube.ctypes.data, cube.shape[1]), "set_environment")

    def RenderEnvironment(self):
        """The colour RT before the volumes: the background (mesh pass) and the environment where the depth is 1."""
        self._ck(self.b.render_environment(self.h), "render_environment")

    def TransformSH(self, cube_rgb):
        cube = np.ascontiguousarray(cube_rgb, dtype=np.float32)
        size = cube.shape[1]
        assert cube.shape == (6, size, size, 3)
        out = np.empty(27, np.float32)
        self._ck(self.b.sh_project(self.h, cube.ctypes.data, size, out.ctypes.data), "sh_project")
        return out.reshape(9, 3)

    def SetShard(self, rank, world):
        self._ck(self.b.set_shard(self.h, rank, world), "set_shard")

    def SetRowBand(self, row0, row1):
        self._ck(self.b.set_row_band(self.h, row0, row1), "set_row_band")

    def SetFrameIndex(self, f):
        self._frame = f
        self._ck(self.b.set_frame_index(self.h, f), "set_frame_index")

    def AdvanceFrame(self):
        """What Render() does to m_frameIdx (MultiRayCaster.cpp:384), for callers that run the passes one by one."""
        self.SetFrameIndex(getattr(self, "_frame", 0) + 1)

    # --- read-backs ---
    def ReadPerObject(self):
        out = np.empty((self.N, 56), np.float32)
        self._ck(self.b.read_per_object(self.h, out.ctypes.data), "read_per_object")
        return out

    def _read_list(self, fn, name):
        ids = np.empty(self.N, np.uint32)
        n = u32(0)
        self._ck(fn(self.h, ids.ctypes.data, C.byref(n)), name)
        return ids[: n.value].copy()

    def ReadVisible(self):
        return self._read_list(self.b.read_visible, "read_visible")

    def ReadCubeVolumes(self):
        return self._read_list(self.b.read_cube_volumes, "read_cube_volumes")

    def ReadAttribs(self):
        out = np.empty((self.N, 4), np.uint16)
        self._ck(self.b.read_attribs(self.h, out.ctypes.data), "read_attribs")
        return out

    def ReadCubeMap(self, volume, mip):
        s = self.G >> mip
        rgba = np.empty((6, s, s, 4), np.uint16)
        depth = np.empty((6, s, s), np.float32)
        self._ck(self.b.read_cubemap(self.h, volume, mip, rgba.ctypes.data, depth.ctypes.data), "read_cubemap")
        return rgba.view(np.float16), depth

    def ReadLightMap(self, volume):
        out = np.empty((self.L, self.L, self.L, 4), np.uint16)
        self._ck(self.b.read_lightmap(self.h, volume, out.ctypes.data), "read_lightmap")
        return out.view(np.float16)

    def ReadFrame(self):
        out = np.empty((self.H, self.W, 4), np.uint16)
        self._ck(self.b.read_frame(self.h, out.ctypes.data), "read_frame")
        return out.view(np.float16)

    def ReadPost(self):
        taa = np.empty((self.H, self.W, 4), np.uint16)
        rgba8 = np.empty((self.H, self.W, 4), np.uint8)
        self._ck(self.b.read_post(self.h, taa.ctypes.data, rgba8.ctypes.data), "read_post")
        return taa.view(np.float16), rgba8

    def GetStats(self):
        s = Stats()
        self._ck(self.b.get_stats(self.h, C.byref(s)), "get_stats")
        return s.as_dict()
