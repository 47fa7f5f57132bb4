#!/usr/bin/env python
"""ORACLE TOOLING (test infrastructure): golden vectors from the reference's own compiled shaders.

    python -m oracle.dxil.make_golden [name ...]        # run HERE (needs /root/reference/Bin/*.cso and llvmlite)

Each case drives one reference shader (DXIL disassembled by container.py, executed by interp.py) on seeded inputs and
stores inputs and outputs under tests/golden/dxil_<name>.npz; tests/test_dxil_golden.py feeds the same inputs to the oracle
and compares. The host-side inputs a shader needs (PerObject records, cbPerFrame) are taken from the oracle, whose host
code has its own tests (tests/test_oracle_kat.py). /root/reference is read only here, never by the tests."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle.dxil.container import disassemble          # noqa: E402
from oracle.dxil.interp import (CBuffer, ResArray, Resources, Shader, StructuredBuffer, Texture, TypedBuffer, run_compute, F32)   # noqa: E402

BIN = "/root/reference/Bin/"
OUT = os.path.join(ROOT, "tests", "golden")
_cache = {}


def shader(name):
    if name not in _cache:
        _cache[name] = Shader(disassemble(BIN + name + ".cso"))
    return _cache[name]


def per_object_bytes(po):
    """oracle PerObject rows (N, 56: WVP 4x4, WVPI 4x4, WorldI 4x3, World 4x3, row-major, row-vector convention) -> the
    StructuredBuffer<PerObject> bytes the shader reads (HLSL matrices are column-major in memory; stride 224)."""
    po = np.asarray(po, np.float32)
    out = np.empty_like(po)
    out[:, 0:16] = po[:, 0:16].reshape(-1, 4, 4).transpose(0, 2, 1).reshape(-1, 16)
    out[:, 16:32] = po[:, 16:32].reshape(-1, 4, 4).transpose(0, 2, 1).reshape(-1, 16)
    out[:, 32:44] = po[:, 32:44].reshape(-1, 4, 3).transpose(0, 2, 1).reshape(-1, 12)
    out[:, 44:56] = po[:, 44:56].reshape(-1, 4, 3).transpose(0, 2, 1).reshape(-1, 12)
    return out


def per_frame_bytes(eye, viewport, screen_to_world=None, shadow_vp=None, light_pos=(0, 0, 0, 0), light_color=(0, 0, 0, 0), ambient=(0, 0, 0, 0), frame_idx=0):
    """cbPerFrame (Common.hlsli:39-49): 212 bytes, matrices column-major"""
    b = np.zeros(53, np.float32)
    b[0:3] = eye
    b[4:6] = viewport
    if screen_to_world is not None: b[8:24] = np.asarray(screen_to_world, np.float32).reshape(4, 4).T.reshape(16)
    if shadow_vp is not None: b[24:40] = np.asarray(shadow_vp, np.float32).reshape(4, 4).T.reshape(16)
    b[40:44], b[44:48], b[48:52] = light_pos, light_color, ambient
    raw = bytearray(b.tobytes())
    raw[208:212] = np.uint32(frame_idx).tobytes()
    return bytes(raw)


# ------------------------------------------------------------------------------------------------ CSVolumeCull
def cull_case(o, eye, max_ray_samples):
    """o: an OracleCaster after UpdateFrame. Runs CSVolumeCull.cso over its PerObject records."""
    N = o.N
    po = o.ReadPerObject()
    descs = np.array([(i % o.srcs) | (5 << 14) | (o.G << 18) for i in range(N)], np.uint32)
    cbf = CBuffer(per_frame_bytes(eye, (o.W, o.H)))
    cbs = CBuffer(np.array([max_ray_samples, 0], np.uint32).tobytes())
    visible = StructuredBuffer(np.full(N, 0xffffffff, np.uint32), 4)
    cubes = StructuredBuffer(np.full(N, 0xffffffff, np.uint32), 4)
    volumes = TypedBuffer(np.zeros((N, 4), np.uint32))
    res = Resources(srv={0: StructuredBuffer(per_object_bytes(po), 224), 1: StructuredBuffer(descs, 4)},
                    uav={0: visible, 1: volumes, 2: cubes}, cbv={0: cbf, 1: cbs})
    run_compute(shader("CSVolumeCull"), res, ((N + 3) // 4, 1, 1), threads_per_group=(8, 4, 1))
    return dict(per_object=po, eye=np.asarray(eye, np.float32), viewport=np.array([o.W, o.H], np.float32), max_ray_samples=max_ray_samples,
                grid=o.G, srcs=o.srcs, visible=visible.words[:visible.counter].copy(), cube_volumes=cubes.words[:cubes.counter].copy(),
                volume_info=volumes.a.copy())


def make_cull():
    from harness import configure
    from oracle_binding import OracleCaster
    cases = {}
    shapes = {"cfg1": dict(grid_size=64, light_grid_size=32, num_volumes=2, width=1280, height=720),
              "cfg2": dict(grid_size=128, light_grid_size=64, num_volumes=16, num_volume_srcs=2, width=1920, height=1080),
              "cfg4": dict(grid_size=256, light_grid_size=128, num_volumes=64, num_volume_srcs=2, width=3840, height=2160),   # (two sources: the cull does not read the volumes)
              "n512": dict(grid_size=64, light_grid_size=16, num_volumes=512, num_volume_srcs=4, width=2560, height=1440)}
    for name, kw in shapes.items():
        for seed, eye in ((0, (4.0, 16.0, -80.0)), (3, (4.0, 16.0, -80.0)), (5, (10.0, 40.0, -160.0)), (9, (30.0, 8.0, -30.0))):
            if (name == "cfg1" and seed == 0) or (name == "n512" and seed in (0, 9)): continue
            o = OracleCaster(filter_model=1, max_ray_samples=256, **kw)
            for i in range(o.srcs): pass
            from multivolumes_b200 import scene
            o.SetVolumesWorld(20.0, (0, 0, 0))
            if seed:
                from harness import rotation_xyz, world43
                rs = np.random.RandomState(seed)
                for i in range(o.N):
                    pos = rs.uniform(-60, 60, 3) * np.array([1, 0.3, 1])
                    o.SetVolumeWorldMatrix(i, world43(rs.uniform(4, 14), rotation_xyz(rs), pos))
            vp, e = scene.default_camera(o.W, o.H, eye=eye)
            o.UpdateFrame(vp, None, e)
            c = cull_case(o, e, 256)
            c.update(view_proj=vp, seed=seed, shape=np.array([kw["grid_size"], kw["light_grid_size"], kw["num_volumes"], kw["width"], kw["height"]]))
            o.Cull()
            vis, cub, att = o.ReadVisible(), o.ReadCubeVolumes(), o.ReadAttribs()
            same = (np.array_equal(np.sort(vis), np.sort(c["visible"])) and np.array_equal(np.sort(cub), np.sort(c["cube_volumes"]))
                    and all(np.array_equal(att[v], c["volume_info"][v].astype(np.uint16)) for v in c["visible"]))
            print(f"cull {name} seed {seed}: visible {len(c['visible'])} cube {len(c['cube_volumes'])}  oracle agrees: {same}")
            cases[f"{name}_s{seed}"] = c
    flat = {f"{k}/{f}": v for k, c in cases.items() for f, v in c.items()}
    np.savez_compressed(os.path.join(OUT, "dxil_cull.npz"), **flat)


# ------------------------------------------------------------------------------------------------ samplers of the harness
def bilinear_2d(tex, smp, coords, offs, lod, cmp_):
    """SampleLevel(g_smpLinear, uv, 0) on a 2-D texture: clamp addressing, texel coordinates in fixed point with 8 fractional
    bits, rounded to nearest (oracle/mvo_sampler.h axis_q8), fp32 blend, zero-weight taps ignored."""
    a = tex.mips[0]
    H, W = a.shape[0], a.shape[1]
    def axis(u, n):
        u = float(np.float32(u)); u = 0.0 if u != u else min(max(u, -1.0), 2.0)
        fx = np.float32(np.float64(np.float32(u)) * n - 0.5)                      # fma(u, n, -0.5)
        xq = int(np.floor(np.float32(np.float64(fx) * 256.0 + 0.5)))             # fma(fx, 256, 0.5)
        xq = min(max(xq, 0), (n - 1) * 256)
        return xq >> 8, min((xq >> 8) + 1, n - 1), F32((xq & 255) / 256.0)
    x0, x1, wx = axis(coords[0], W); y0, y1, wy = axis(coords[1], H)
    t00, t10 = a[y0, x0].astype(np.float32), a[y0, x1].astype(np.float32)
    t01, t11 = a[y1, x0].astype(np.float32), a[y1, x1].astype(np.float32)
    fma = lambda a_, b_, w: a_ if w == 0 else ((b_ - a_).astype(np.float32).astype(np.float64) * np.float64(w) + a_.astype(np.float64)).astype(np.float32)
    r = fma(fma(t00, t10, wx), fma(t01, t11, wx), wy)
    return [F32(r[k]) if k < r.size else F32(0) for k in range(4)]


# ------------------------------------------------------------------------------------------------ CSTemporalAA / PSToneMap
def taa_inputs(W, H, seed):
    rs = np.random.RandomState(seed)
    def img():
        y, x = np.mgrid[0:H, 0:W]
        base = 0.5 + 0.5 * np.sin(x[..., None] * 0.37 + y[..., None] * 0.23 + np.arange(3) * 1.3)
        c = (base * rs.uniform(0.0, 3.0, (H, W, 1)) + rs.uniform(0, 0.05, (H, W, 3))).astype(np.float16)
        return c
    cur = np.concatenate([img(), np.where(rs.uniform(size=(H, W, 1)) < 0.7, 1.0, rs.uniform(0.0, 0.99, (H, W, 1))).astype(np.float16)], -1)
    hist = np.concatenate([(img().astype(np.float32) * 0.9 + cur[..., :3].astype(np.float32) * 0.1).astype(np.float16),
                           rs.choice([0.0, 1.0 / 15.0, 0.5, 1.0], size=(H, W, 1)).astype(np.float16)], -1)
    vel = np.zeros((H, W, 2), np.float16)
    if seed % 2:
        vel[...] = (rs.uniform(-1.5, 1.5, (H, W, 2)) / np.array([W, H])).astype(np.float16)
        vel[rs.uniform(size=(H, W)) < 0.5] = 0
    return cur, hist, vel


def run_taa(cur, hist, vel):
    import oracle.dxil.interp as I
    H, W = cur.shape[:2]
    out = {}
    for mode, promote in (("f32", True), ("f16", False)):
        I.PROMOTE_HALF = promote
        sh = Shader(disassemble(BIN + "CSTemporalAA.cso"))
        rt = Texture(np.zeros((H, W, 4), np.float16))
        res = Resources(srv={0: Texture(cur), 1: Texture(hist), 2: Texture(vel)}, uav={0: rt}, sampler={0: None}, sample=bilinear_2d)
        run_compute(sh, res, ((W + 7) // 8, (H + 7) // 8, 1), threads_per_group=(8, 8, 1))
        out[mode] = rt.a.copy()
    I.PROMOTE_HALF = True
    return out


def run_tonemap(taa):
    import oracle.dxil.interp as I
    H, W = taa.shape[:2]
    out = {}
    for mode, promote in (("f32", True), ("f16", False)):
        I.PROMOTE_HALF = promote
        sh = Shader(disassemble(BIN + "PSToneMap.cso"))
        img = np.zeros((H, W, 4), np.float32)
        res = Resources(srv={0: Texture(taa)})
        for y in range(H):
            for x in range(W):
                o = {}
                sh.run_wave([sh.lane(res, {}, inputs={0: {0: F32(x + 0.5), 1: F32(y + 0.5), 2: F32(0), 3: F32(1)}}, outputs=o)])
                img[y, x] = [float(o[0][k]) for k in range(4)]
        out[mode] = np.floor(np.clip(img, 0, 1) * 255.0 + 0.5).astype(np.uint8)      # R8G8B8A8_UNORM write
    I.PROMOTE_HALF = True
    return out


def make_post():
    from oracle_binding import OracleCaster
    flat = {}
    for seed, (W, H) in ((1, (40, 24)), (2, (40, 24)), (3, (33, 19))):
        cur, hist, vel = taa_inputs(W, H, seed)
        taa = run_taa(cur, hist, vel)
        tm = run_tonemap(taa["f32"])
        from oracle_binding import oracle_binding
        res_ = {}
        for as_half in (1, 0):
            oracle_binding().set_min16_consts_as_half(as_half)
            o = OracleCaster(filter_model=0, grid_size=32, light_grid_size=16, num_volumes=1, width=W, height=H)
            o.SetRenderTargets(color=hist); o.Postprocess(False)
            o.SetRenderTargets(color=cur, velocity=vel); o.Postprocess(True)
            res_[as_half] = o.ReadPost()
        oracle_binding().set_min16_consts_as_half(1)
        got, rgba8 = res_[1]
        dh = np.abs(res_[0][0].astype(np.float32) - taa["f32"].astype(np.float32))
        print(f"   (oracle with the decimal literals of the HLSL text: max abs {dh.max():.3e}, rgba8 {np.abs(res_[0][1].astype(int) - tm['f32'].astype(int)).max()})")
        d32 = np.abs(got.astype(np.float32) - taa["f32"].astype(np.float32)); d16 = np.abs(got.astype(np.float32) - taa["f16"].astype(np.float32))
        print(f"post seed {seed} {W}x{H}: oracle vs DXIL(min16 as f32) max abs {d32.max():.3e}, vs DXIL(min16 as f16) {d16.max():.3e}; "
              f"rgba8 max diff {np.abs(rgba8.astype(int) - tm['f32'].astype(int)).max()} (f16 tone map: {np.abs(rgba8.astype(int) - tm['f16'].astype(int)).max()})")
        for k, v in dict(current=cur, history=hist, velocity=vel, taa_f32=taa["f32"], taa_f16=taa["f16"], rgba8_f32=tm["f32"], rgba8_f16=tm["f16"]).items():
            flat[f"s{seed}/{k}"] = v.view(np.uint16) if v.dtype == np.float16 else v
    np.savez_compressed(os.path.join(OUT, "dxil_post.npz"), **flat)


# ------------------------------------------------------------------------------------------------ CSRayMarchV
class OracleTex:
    """a texture whose FILTER is the oracle's model of the texture unit (the shader code is what is under test, not the unit)"""
    def __init__(self, o, kind, index):
        self.o, self.kind, self.index = o, kind, index


def make_sampler(o, depth=None):
    import ctypes as C
    from multivolumes_b200._abi import P, f32
    uvw = np.zeros(3, np.float32); out = np.zeros(4, np.float32)
    pu, po = uvw.ctypes.data_as(P(f32)), out.ctypes.data_as(P(f32))

    def sample(tex, smp, coords, offs, lod, cmp_):
        if isinstance(tex, OracleTex):
            uvw[:] = [coords[0], coords[1], coords[2]]
            if any(offs[k] not in (None, 0) for k in range(3)):      # integer texel offsets, applied in normalised space as the oracle does
                inv = F32(1.0) / F32(o.G if tex.kind == "volume" else o.L)
                for k in range(3):
                    ok = offs[k] or 0
                    ok = ok - (1 << 32) if ok >> 31 else ok
                    uvw[k] = F32(uvw[k]) + F32(F32(ok) * inv)
            (o.b.sample_volume if tex.kind == "volume" else o.b.sample_lightmap)(o.h, tex.index, pu, po)
            return [F32(out[0]), F32(out[1]), F32(out[2]), F32(out[3])]
        if cmp_ is not None:                                     # ShadowTest: SampleCmpLevelZero(g_smpShadow, uv, ref), LINEAR, LESS_EQUAL, clamp; D16 texels
            a = tex.mips[0]
            S = a.shape[0]
            fx = F32(F32(coords[0]) * F32(S)) - F32(0.5); fy = F32(F32(coords[1]) * F32(S)) - F32(0.5)
            flx, fly = np.floor(fx), np.floor(fy)
            wx, wy = F32(fx - flx), F32(fy - fly)
            ix, iy = int(flx), int(fly)
            def tap(x, y):
                d = F32(F32(a[min(max(y, 0), S - 1), min(max(x, 0), S - 1), 0]) / F32(65535.0))
                return F32(1.0) if F32(cmp_) <= d else F32(0.0)
            t00, t10, t01, t11 = tap(ix, iy), tap(ix + 1, iy), tap(ix, iy + 1), tap(ix + 1, iy + 1)
            top = F32(t00 + F32(F32(t10 - t00) * wx)); bot = F32(t01 + F32(F32(t11 - t01) * wx))
            return [F32(top + F32(F32(bot - top) * wy)), F32(0), F32(0), F32(0)]
        if smp == "point":                                       # g_txDepth.SampleLevel(g_smpPoint, uv, 0): POINT_CLAMP
            a = tex.mips[0]
            H, W = a.shape[:2]
            u, v = F32(coords[0]), F32(coords[1])
            ix = 0 if u != u else int(np.floor(F32(u * F32(W)))); iy = 0 if v != v else int(np.floor(F32(v * F32(H))))
            t = a[min(max(iy, 0), H - 1), min(max(ix, 0), W - 1)]
            return [F32(t[0]), F32(0), F32(0), F32(0)]
        return bilinear_2d(tex, smp, coords, offs, lod, cmp_)
    return sample


def march_v_case(o, eye, depth, max_ray_samples, f32=False):
    """f32: keep the UAV stores as fp32 (what the shader computed before the RGBA16F conversion)"""
    N, G = o.N, o.G
    po = o.ReadPerObject()
    cubes = o.ReadCubeVolumes()
    att = o.ReadAttribs().astype(np.uint32)
    cube_maps = [Texture(np.zeros((6, G >> m, G >> m, 4), np.float32 if f32 else np.float16)) for v in range(N) for m in range(5)]
    cube_depths = [Texture(np.full((6, G >> m, G >> m, 1), -1.0, np.float32)) for v in range(N) for m in range(5)]
    res = Resources(
        srv={0: ResArray(3, [OracleTex(o, "lightmap", v) for v in range(N)]), 1: ResArray(0, [OracleTex(o, "volume", s) for s in range(o.srcs)]),
             2: Texture(depth[..., None]), 3: StructuredBuffer(per_object_bytes(po), 224), 4: StructuredBuffer(cubes.astype(np.uint32), 4), 5: TypedBuffer(att)},
        uav={0: ResArray(0, cube_maps), 1: ResArray(0, cube_depths)},
        cbv={0: CBuffer(per_frame_bytes(eye, (o.W, o.H)))}, sampler={0: "linear", 1: "point"}, sample=make_sampler(o))
    out = {}
    sh = shader("CSRayMarchV")
    for ci, v in enumerate(cubes):
        s = G >> int(att[v, 0])
        g = (s + 7) // 8
        # Dispatch(g, g, count) restricted to this volume's z slice of groups
        run_compute_z(sh, res, (g, g), ci, threads_per_group=(8, 8, 6))
        out[int(v)] = (int(att[v, 0]), cube_maps[5 * int(v) + int(att[v, 0])].a.copy(), cube_depths[5 * int(v) + int(att[v, 0])].a[..., 0].copy())
    return out


def run_compute_z(sh, res, groups_xy, gz, threads_per_group):
    """the groups (x, y, gz) of a dispatch"""
    from oracle.dxil.interp import run_compute as rc
    tx, ty, tz = threads_per_group
    n = tx * ty * tz
    for gy in range(groups_xy[1]):
        for gx in range(groups_xy[0]):
            for w0 in range(0, n, 32):
                lanes, envs = [], []
                for l in range(w0, min(w0 + 32, n)):
                    lx, ly, lz = l % tx, (l // tx) % ty, l // (tx * ty)
                    sv = {"threadId": (gx * tx + lx, gy * ty + ly, gz * tz + lz), "groupId": (gx, gy, gz), "threadIdInGroup": (lx, ly, lz), "flattenedThreadIdInGroup": l}
                    envs.append({})
                    lanes.append(sh.lane(res, sv, lane_index=l - w0, lane_count=32, env=envs[-1]))
                sh.run_wave(lanes, envs)


def compare_f16(a, b):
    a32, b32 = a.astype(np.float32), b.astype(np.float32)
    return float(np.abs(a32 - b32).max()), int((a.view(np.uint16) != b.view(np.uint16)).sum()), a.size


def make_march_v():
    """keys: f<filter model>/s<scene>/v<volume>/{mip, rgba, depth}; filter model 0 = exact fp32 trilinear, 1 = the sm_100a texture
    unit (what the product's tex3D returns) — the shader code is the same, the texture unit is the caller's"""
    import oracle.dxil.interp as I
    from harness import DXIL_SCENES, dxil_scene
    from oracle_binding import OracleCaster, oracle_binding
    flat = {}
    I.PROMOTE_HALF = True
    for model in (0, 1):
        for name in DXIL_SCENES:
            for as_half in (1, 0):
                oracle_binding().set_min16_consts_as_half(as_half)
                o, vp, eye, depth, shadow = dxil_scene(OracleCaster, name, filter_model=model)
                if as_half:
                    dx = march_v_case(o, eye, depth, DXIL_SCENES[name]["ray"])
                    dx32 = march_v_case(o, eye, depth, DXIL_SCENES[name]["ray"], f32=True)
                    o.DebugF32(True)
                o.RayMarchV()
                if as_half:
                    cube32, _ = o.DebugF32(True)
                    bad = tot = 0
                    for v, (mip, rgba32, dep) in dx32.items():
                        s_ = o.G >> mip
                        m_ = dep >= 0
                        bad += int((cube32[v, :, :s_, :s_].view(np.uint32)[m_] != rgba32.view(np.uint32)[m_]).sum()); tot += int(m_.sum()) * 4
                        flat[f"f{model}/{name}/v{v}/rgba_f32"] = rgba32
                    print(f"march_v filter {model} scene {name}: fp32 scatter before the RGBA16F store: {bad}/{tot} values differ")
                for v, (mip, rgba, dep) in dx.items():
                    orgba, odep = o.ReadCubeMap(v, mip)
                    mask = dep >= 0                                      # texels the shader wrote
                    mx, nbits, n = compare_f16(rgba[mask], orgba.view(np.float16)[mask])
                    print(f"march_v filter {model} scene {name} volume {v} mip {mip} (oracle literals: {'shipped DXIL' if as_half else 'HLSL text'}): {int(mask.sum())} rays, "
                          f"depth equal {np.array_equal(dep[mask], odep[mask])}, colour max abs {mx:.3e}, {nbits}/{n} fp16 values differ")
                    if as_half:
                        k = f"f{model}/{name}/v{v}"
                        flat[k + "/mip"] = np.int32(mip); flat[k + "/rgba"] = rgba.view(np.uint16); flat[k + "/depth"] = dep
            flat[f"f{model}/{name}/cubes"] = np.array(sorted(dx), np.int32)
    oracle_binding().set_min16_consts_as_half(1)
    np.savez_compressed(os.path.join(OUT, "dxil_march_v.npz"), **flat)


# ------------------------------------------------------------------------------------------------ CSRayMarchL
def march_l_case(o, volume, shadow, shadow_vp, light, sh, num_light_samples):
    """CSRayMarchL.cso for one volume's light map (frame index chosen so that the round-robin picks `volume`)"""
    N, L = o.N, o.L
    po = o.ReadPerObject()
    visible = o.ReadVisible().astype(np.uint32)
    frame_idx = int(np.where(visible == volume)[0][0])
    descs = np.array([(i % o.srcs) | (5 << 14) | (o.G << 18) for i in range(N)], np.uint32)
    q = lambda k, v: o.b.quantize_r11(float(v)) if k < 2 else o.b.quantize_b10(float(v))
    maps = [Texture(np.zeros((L, L, L, 3), np.float32), quantise=q) for _ in range(N)]
    pf = per_frame_bytes(light["eye"], (o.W, o.H), shadow_vp=shadow_vp, light_pos=light["pos"], light_color=light["color"], ambient=light["ambient"], frame_idx=frame_idx)
    res = Resources(
        srv={0: ResArray(0, [OracleTex(o, "volume", s) for s in range(o.srcs)]), 1: Texture(shadow[..., None]),
             2: StructuredBuffer(np.asarray(sh, np.float32).reshape(9, 3), 12), 3: StructuredBuffer(per_object_bytes(po), 224),
             4: StructuredBuffer(descs, 4), 5: StructuredBuffer(visible, 4), 6: StructuredBuffer(np.array([len(visible)], np.uint32), 4)},
        uav={0: ResArray(0, maps)}, cbv={0: CBuffer(pf), 1: CBuffer(np.array([num_light_samples, 1 if sh is not None else 0], np.uint32).tobytes())},
        sampler={0: "linear", 1: "shadow"}, sample=make_sampler(o))
    run_compute(shader("CSRayMarchL"), res, (L // 4, L // 4, L // 4), threads_per_group=(4, 4, 4))
    return maps[volume].a


def make_march_l():
    """keys: f<filter model>/<scene>/v<volume>: (L, L, L, 3) float32 light map as CSRayMarchL.cso stores it (R11G11B10_FLOAT values)"""
    import oracle.dxil.interp as I
    from harness import DXIL_SCENES, dxil_scene, sh_coeffs
    from multivolumes_b200 import scene
    from oracle_binding import OracleCaster, oracle_binding
    flat = {}
    I.PROMOTE_HALF = True
    for model in (0, 1):
        for name, cfg in DXIL_SCENES.items():
            for as_half in (1, 0):
                oracle_binding().set_min16_consts_as_half(as_half)
                o, vp, eye, depth, shadow = dxil_scene(OracleCaster, name, filter_model=model, light_maps=False)
                light = dict(eye=eye, pos=tuple(scene.LIGHT_PT) + (1.0,), color=tuple(scene.LIGHT_COLOR) + (scene.LIGHT_INTENSITY,),
                             ambient=tuple(scene.AMBIENT_COLOR) + (scene.AMBIENT_INTENSITY,))
                for v in o.ReadVisible():
                    v = int(v)
                    k = f"f{model}/{name}/v{v}"
                    if as_half:
                        flat[k] = march_l_case(o, v, shadow, scene.shadow_view_proj(), light, sh_coeffs(), cfg["light"])
                    o.RayMarchL(v)
                    got = o.ReadLightMap(v).view(np.float16)[..., :3].astype(np.float32)
                    want = flat[k]
                    d = np.abs(got - want)
                    print(f"march_l filter {model} scene {name} volume {v} (oracle literals: {'shipped DXIL' if as_half else 'HLSL text'}): max abs {d.max():.3e}, "
                          f"{int((got != want).sum())}/{got.size} values differ, lit voxels {int((want.sum(-1) > 0).sum())}")
    oracle_binding().set_min16_consts_as_half(1)
    np.savez_compressed(os.path.join(OUT, "dxil_march_l.npz"), **flat)


# ------------------------------------------------------------------------------------------------ CSSHCubeMap / CSSHSum / CSSHNormalize
def cube_texel_sample(tex, smp, coords, offs, lod, cmp_):
    """TextureCube.SampleLevel at a texel-centre direction: that texel (D3D face / (u, v) convention)"""
    x, y, z = [float(c) for c in coords[:3]]
    ax, ay, az = abs(x), abs(y), abs(z)
    if ax >= ay and ax >= az: f, u, v, m = (0 if x > 0 else 1), (-z if x > 0 else z), -y, ax
    elif ay >= az: f, u, v, m = (2 if y > 0 else 3), x, (z if y > 0 else -z), ay
    else: f, u, v, m = (4 if z > 0 else 5), (x if z > 0 else -x), -y, az
    S = tex.shape[1]
    i = int(np.floor((u / m * 0.5 + 0.5) * S)); j = int(np.floor((v / m * 0.5 + 0.5) * S))
    t = tex[f, min(max(j, 0), S - 1), min(max(i, 0), S - 1)]
    return [F32(t[0]), F32(t[1]), F32(t[2]), F32(1)]


def sh_project_dxil(cube, order=3):
    """XUSG's SH transform as its three compiled kernels run it (no HLSL source in the reference tree): per-texel projection with
    one partial sum per 32-lane group, tree reduction 32:1 until one group is left, normalisation by 4 pi / sum of weights."""
    size = cube.shape[1]
    n = 6 * size * size
    groups = (n + 31) // 32
    shbuf = StructuredBuffer(np.zeros((groups * order * order, 3), np.float32), 12)
    wbuf = StructuredBuffer(np.zeros(groups, np.float32), 4)
    run_compute(shader("CSSHCubeMap"), Resources(srv={0: cube}, uav={0: shbuf, 1: wbuf}, cbv={0: CBuffer(np.array([order, size], np.uint32).tobytes())},
                                                sampler={0: None}, sample=cube_texel_sample), (groups, 1, 1))
    count, src_c, src_w = groups, shbuf, wbuf
    while count > 1:
        g2 = (count + 31) // 32
        dst_c = StructuredBuffer(np.zeros((g2 * order * order, 3), np.float32), 12); dst_w = StructuredBuffer(np.zeros(g2, np.float32), 4)
        run_compute(shader("CSSHSum"), Resources(srv={0: src_c, 1: src_w}, uav={0: dst_c, 1: dst_w}, cbv={0: CBuffer(np.array([order, count], np.uint32).tobytes())}),
                    (g2, order * order, 1))
        src_c, src_w, count = dst_c, dst_w, g2
    out = StructuredBuffer(np.zeros((order * order, 3), np.float32), 12)
    run_compute(shader("CSSHNormalize"), Resources(srv={0: src_c, 1: src_w}, uav={0: out}), (1, 1, 1), threads_per_group=(order * order, 1, 1))
    return out.words.view(np.float32).reshape(order * order, 3).copy()


def make_sh():
    from multivolumes_b200 import scene
    from oracle_binding import OracleCaster
    flat = {}
    o = OracleCaster(filter_model=1, grid_size=32, light_grid_size=16, num_volumes=1, width=64, height=48)
    cubes = {"noise8": np.random.RandomState(1).uniform(0, 2, (6, 8, 8, 3)).astype(np.float32), "sky16": scene.procedural_sky(16).astype(np.float32)}
    for name, cube in cubes.items():
        want = sh_project_dxil(cube)
        got = o.TransformSH(cube)
        print(f"sh {name}: oracle vs the three reference kernels: max abs {np.abs(got - want).max():.3e} (largest coefficient {np.abs(want).max():.3f})")
        flat[name + "/cube"] = cube; flat[name + "/coeffs"] = want
    np.savez_compressed(os.path.join(OUT, "dxil_sh.npz"), **flat)


# ------------------------------------------------------------------------------------------------ CSInitGridData / CSR32FToRGBA16F
def make_init():
    from oracle_binding import OracleCaster
    flat = {}
    half = lambda k, v: np.float16(v)
    for G in (16, 24):
        grid = Texture(np.zeros((G, G, G, 4), np.float16), quantise=half)
        run_compute(shader("CSInitGridData"), Resources(uav={0: grid}), ((G + 3) // 4,) * 3, threads_per_group=(4, 4, 4))
        o = OracleCaster(filter_model=1, grid_size=G, light_grid_size=8, num_volumes=1, width=64, height=48)
        o.InitVolumeData(0, 0, 0)
        got = o.ReadVolume(0)
        n = int((got.view(np.uint16) != grid.a.view(np.uint16)).sum())
        print(f"init G={G}: {n}/{got.size} halves differ from CSInitGridData.cso, max abs {np.abs(got.astype(np.float32) - grid.a.astype(np.float32)).max():.3e}")
        flat[f"g{G}/rgba"] = grid.a.view(np.uint16)
        # CSR32FToRGBA16F: density (R32F) -> RGBA16F
        rs = np.random.RandomState(G)
        dens = rs.uniform(0, 1.2, (G, G, G, 1)).astype(np.float32) * (rs.uniform(size=(G, G, G, 1)) > 0.3)
        dst = Texture(np.zeros((G, G, G, 4), np.float16), quantise=half)
        def centre(tex, smp, coords, offs, lod, cmp_):      # LINEAR fetch at a texel centre of a same-sized source: that texel
            i = [min(max(int(np.floor(F32(coords[k]) * F32(G))), 0), G - 1) for k in range(3)]
            return [F32(tex.a[i[2], i[1], i[0], 0]), F32(0), F32(0), F32(0)]
        run_compute(shader("CSR32FToRGBA16F"), Resources(srv={0: Texture(dens)}, uav={0: dst}, sampler={0: "linear"}, sample=centre), ((G + 3) // 4,) * 3,
                    threads_per_group=(4, 4, 4))
        o.LoadVolumeData(0, dens[..., 0])
        got = o.ReadVolume(0)
        n = int((got.view(np.uint16) != dst.a.view(np.uint16)).sum())
        print(f"r32f G={G}: {n}/{got.size} halves differ from CSR32FToRGBA16F.cso")
        flat[f"g{G}/density"] = dens[..., 0]; flat[f"g{G}/converted"] = dst.a.view(np.uint16)
    np.savez_compressed(os.path.join(OUT, "dxil_init.npz"), **flat)


# ------------------------------------------------------------------------------------------------ PSBasePass
def make_base_pass():
    """PSBasePass.cso per pixel of a screen-filling clip-space quad (identity world / view-projection: every interpolant is an
    affine function of the pixel centre, so the pixel shader's inputs are known without a rasteriser)."""
    import oracle.dxil.interp as I
    from harness import sh_coeffs
    from oracle_binding import OracleCaster
    I.PROMOTE_HALF = True
    flat = {}
    W, H = 48, 32
    a, b, c, d = (-1, -1, .25), (1, -1, .25), (1, 1, .25), (-1, 1, .25)
    pos = np.asarray([a, c, b, a, d, c], np.float32)
    eye, light, lrgbi, argbi = (0.3, 0.2, -3.0), (0.6, 0.9, -1.0), (1.0, 0.7, 0.3, 2.0), (0.4, 0.6, 1.0, 1.5)
    for use_sh in (0, 1):
        o = OracleCaster(filter_model=1, grid_size=32, light_grid_size=16, num_volumes=1, width=W, height=H)
        o.SetLight(light, lrgbi[:3], lrgbi[3]); o.SetAmbient(argbi[:3], argbi[3])
        o.SetSH(sh_coeffs() if use_sh else None)
        o.SetMesh(pos, np.arange(6, dtype=np.uint32)); o.SetMeshWorld(1.0, (0, 0, 0))
        ident = np.eye(4, dtype=np.float32)
        svp = o.RenderMesh(ident, eye)
        got = o.ReadFrame()
        _, shadow = o.ReadDepth()
        cbf = np.zeros(16, np.float32); cbf[0:3] = eye; cbf[4:7] = light; cbf[8:12] = lrgbi; cbf[12:16] = argbi
        res = Resources(srv={0: Texture(shadow[..., None]), 1: StructuredBuffer(sh_coeffs().reshape(9, 3), 12), 2: None},
                        cbv={0: CBuffer(cbf.tobytes()), 1: CBuffer(np.array([1 if use_sh else 0], np.uint32).tobytes())},
                        sampler={0: "shadow", 1: "linear"}, sample=make_sampler(o))
        sh_ = shader("PSBasePass")
        out = np.zeros((H, W, 4), np.float16); vel = np.zeros((H, W, 2), np.float16)
        M = svp.astype(np.float32)
        for y in range(H):
            for x in range(W):
                p = np.array([F32((x + 0.5) / W * 2 - 1), F32(1 - (y + 0.5) / H * 2), F32(0.25)], np.float32)
                ls = [F32(F32(F32(F32(p[0] * M[0, k]) + F32(p[1] * M[1, k])) + F32(p[2] * M[2, k])) + M[3, k]) for k in range(4)]
                inputs = {1: {0: p[0], 1: p[1], 2: p[2]}, 2: {0: F32(0), 1: F32(0), 2: F32(-1)}, 3: {0: ls[0], 1: ls[1], 2: ls[2], 3: ls[3]},
                          4: {0: p[0], 1: p[1], 2: p[2], 3: F32(1)}, 5: {0: p[0], 1: p[1], 2: p[2], 3: F32(1)}}
                o_ = {}
                sh_.run_wave([sh_.lane(res, {}, inputs=inputs, outputs=o_)])
                out[y, x] = [np.float16(o_[0][k]) for k in range(4)]
                vel[y, x] = [np.float16(o_[1][k]) for k in range(2)]
        ulps = np.abs(got.view(np.int16).astype(np.int32) - out.view(np.int16).astype(np.int32))
        print(f"base pass (SH {'on' if use_sh else 'off'}): oracle vs PSBasePass.cso: max {int(ulps.max())} binary16 steps, {float((ulps > 0).mean()) * 100:.2f} % of halves differ; "
              f"max abs {np.abs(got.astype(np.float32) - out.astype(np.float32)).max():.3e}")
        flat[f"sh{use_sh}/rgba"] = out.view(np.uint16); flat[f"sh{use_sh}/velocity"] = vel.view(np.uint16)
    flat["mesh"] = pos; flat["eye"] = np.asarray(eye, np.float32); flat["light"] = np.asarray(light, np.float32)
    flat["light_rgbi"] = np.asarray(lrgbi, np.float32); flat["ambient_rgbi"] = np.asarray(argbi, np.float32)
    np.savez_compressed(os.path.join(OUT, "dxil_base_pass.npz"), **flat)


# ------------------------------------------------------------------------------------------------ PSCube (CubeCast / RayCast) + PSResolveOIT
class CubeTex:
    """TextureCube view of one mip of a volume's cube map (colour RGBA16F or depth R32F); faces [6][S][S][C]"""
    def __init__(self, arr):
        self.a = np.asarray(arr)
        self.mips = [self.a]

    def dims(self, mip=0):
        return (self.a.shape[2], self.a.shape[1])


def _fma(a, b, c):
    return np.float32(np.float64(np.float32(a)) * np.float64(np.float32(b)) + np.float64(np.float32(c)))


def make_cube_callbacks(o, frag):
    """sampleLevel / textureGather on the cube maps for the fragment in `frag` (dict: face): the four texels around the
    direction's face (u, v), resolved across cube edges as a seamless TextureCube does (mvo_cube_resolve_texel), in Gather
    order (-,+) (+,+) (+,-) (-,-). The texture unit is the caller's; PSCube's arithmetic is what is under test."""
    import ctypes as C
    base = make_sampler(o)
    out3 = (C.c_int * 3)()

    def footprint(tex, coords):
        p = [F32(coords[0]), F32(coords[1]), F32(coords[2])]
        f = frag["face"]
        h = F32(0.5)
        u, v = {0: (F32(F32(-p[2] * h) + h), F32(F32(-p[1] * h) + h)), 1: (F32(F32(p[2] * h) + h), F32(F32(-p[1] * h) + h)),
                2: (F32(F32(p[0] * h) + h), F32(F32(p[2] * h) + h)), 3: (F32(F32(p[0] * h) + h), F32(F32(-p[2] * h) + h)),
                4: (F32(F32(p[0] * h) + h), F32(F32(-p[1] * h) + h)), 5: (F32(F32(-p[0] * h) + h), F32(F32(-p[1] * h) + h))}[f]
        S = tex.a.shape[1]
        fx, fy = _fma(u, S, -0.5), _fma(v, S, -0.5)
        i0, j0 = int(np.floor(fx)), int(np.floor(fy))
        tex_ = []
        for ti, tj in ((i0, j0 + 1), (i0 + 1, j0 + 1), (i0 + 1, j0), (i0, j0)):
            o.b.cube_resolve_texel(S, f, ti, tj, out3)
            tex_.append(tex.a[out3[0], out3[2], out3[1]])
        return tex_, F32(fx - np.floor(fx)), F32(fy - np.floor(fy))

    def gather(tex, smp, coords, offs, channel):
        t, _, _ = footprint(tex, coords)
        return [F32(t[k][channel]) for k in range(4)]

    def sample(tex, smp, coords, offs, lod, cmp_):
        if not isinstance(tex, CubeTex):
            return base(tex, smp, coords, offs, lod, cmp_)
        t, bx, by = footprint(tex, coords)
        bw = [F32(F32(F32(1) - bx) * by), F32(bx * by), F32(bx * F32(F32(1) - by)), F32(F32(F32(1) - bx) * F32(F32(1) - by))]
        col = [F32(0)] * 4
        for k in range(4):
            col = [_fma(F32(t[k][c]) if c < len(t[k]) else F32(0), bw[k], col[c]) for c in range(4)]
        return col
    return sample, gather


def oit_case(o, eye, depth, stride=3, f32=False):
    """PSCube.cso per fragment and PSResolveOIT.cso per pixel, on every stride-th pixel, from the fragments the oracle's
    analytic rasteriser produced (depth key, exit point on the cube, face uv); returns per-layer colours and the blend."""
    N, G, W, H = o.N, o.G, o.W, o.H
    cnt, info, data, result = o.DebugOIT()
    po = o.ReadPerObject()
    att = o.ReadAttribs().astype(np.uint32)
    cubes_c, cubes_d = [], []
    for v in range(N):
        for m in range(5):
            if m == int(att[v, 0]) and (att[v, 2] & 0x8000):
                rgba, dep = o.ReadCubeMap(v, m)
                cubes_c.append(CubeTex(rgba.view(np.float16))); cubes_d.append(CubeTex(dep[..., None]))
            else:
                cubes_c.append(None); cubes_d.append(None)
    frag = {}
    sample, gather = make_cube_callbacks(o, frag)
    kcolors = Texture(np.zeros((8, H, W, 4), np.float32)) if f32 else Texture(np.zeros((8, H, W, 4), np.float16), quantise=lambda k, v: np.float16(v))
    kdepths = Texture(np.full((8, H, W, 1), 0xffffffff, np.uint32))
    res = Resources(
        srv={0: ResArray(3, [OracleTex(o, "lightmap", v) for v in range(N)]), 1: ResArray(0, [OracleTex(o, "volume", s) for s in range(o.srcs)]),
             2: Texture(depth[..., None]), 3: StructuredBuffer(per_object_bytes(po), 224), 4: ResArray(0, cubes_c), 5: ResArray(0, cubes_d), 6: kdepths},
        uav={0: kcolors}, cbv={0: CBuffer(per_frame_bytes(eye, (W, H)))}, sampler={0: "linear"}, sample=sample, gather=gather)
    ps, rs = shader("PSCube"), shader("PSResolveOIT")
    layers = np.zeros((H, W, 8, 4), np.float32 if f32 else np.float16); stored = np.zeros((H, W, 8), np.uint8); blend = np.zeros((H, W, 4), np.float32); done = np.zeros((H, W), bool)
    for py in range(0, H, stride):
        for px in range(0, W, stride):
            n = int(cnt[py, px])
            if n == 0:
                continue
            kcolors.a[:, py, px] = 0
            for l in range(8):
                kdepths.a[l, py, px, 0] = info[py, px, l, 0] if l < n else 0xffffffff
            for l in range(n):
                key, vol, face, _ = [int(x) for x in info[py, px, l]]
                if l and key == int(info[py, px, l - 1, 0]):
                    continue                                       # equal depth keys: the shader writes every matching layer at once
                frag["face"] = face
                smp = 0 if (att[vol, 2] & 0x8000) else int(att[vol, 1])
                z = np.array(key, np.uint32).view(np.float32)[()]
                d = data[py, px, l]
                inputs = {0: {0: F32(px + 0.5), 1: F32(py + 0.5), 2: z, 3: F32(1)}, 1: {0: d[3], 1: d[4], 2: F32(0)}, 2: {0: d[0], 1: d[1], 2: d[2]},
                          3: {0: vol}, 4: {0: 5 * vol + int(att[vol, 0])}, 5: {0: int(att[vol, 3])}, 6: {0: smp}}
                ps.run_wave([ps.lane(res, {}, inputs=inputs, outputs={})])
            layers[py, px] = kcolors.a[:, py, px]
            out = {}
            rs.run_wave([rs.lane(Resources(srv={0: kcolors}), {}, inputs={0: {0: F32(px + 0.5), 1: F32(py + 0.5)}}, outputs=out)])
            blend[py, px] = [float(out[0][k]) for k in range(4)]
            done[py, px] = True
    return dict(count=cnt, info=info, data=data, layers=layers, blend=blend, done=done, oracle_result=result)


def make_oit():
    import oracle.dxil.interp as I
    from harness import DXIL_SCENES, dxil_scene
    from oracle_binding import OracleCaster
    I.PROMOTE_HALF = True
    flat = {}
    from harness import nested_scene
    for name in list(DXIL_SCENES) + ["nested"]:
        if name == "nested":                                  # 12 concentric volumes: up to 8 blended layers per pixel
            o, vp, eye = nested_scene(OracleCaster, filter_model=1)
            depth = np.ones((o.H, o.W), np.float32)
            o.Cull()
            for v in range(o.N): o.RayMarchL(v)
        else:
            o, vp, eye, depth, shadow = dxil_scene(OracleCaster, name, filter_model=1)
        o.RayMarchV()
        r = oit_case(o, eye, depth)
        r32 = oit_case(o, eye, depth, f32=True)
        m = r["done"]
        w32 = np.zeros_like(r32["layers"])
        for py, px in np.argwhere(m):
            for l in range(int(r["count"][py, px])):
                if r["info"][py, px, l, 3]:
                    w32[py, px, l] = r["data"][py, px, l, 5:9]
        print(f"oit scene {name}: fp32 fragment colours before the RGBA16F store: {int((w32.view(np.uint32)[m] != r32['layers'].view(np.uint32)[m]).sum())}/{int(m.sum()) * 32} values differ")
        # per-layer colours: the oracle's, rounded to the K-colour format
        want = np.zeros_like(r["layers"])
        cnt, info, data = r["count"], r["info"], r["data"]
        for py, px in np.argwhere(m):
            for l in range(int(cnt[py, px])):
                if info[py, px, l, 3]:
                    want[py, px, l] = data[py, px, l, 5:9].astype(np.float16)
        ul = np.abs(want.view(np.int16).astype(np.int32) - r["layers"].view(np.int16).astype(np.int32))[m]
        db = np.abs(r["blend"][m] - r["oracle_result"][m])
        print(f"oit scene {name}: {int(m.sum())} pixels, {int(cnt[m].sum())} fragments ({int((info[..., 1][m][:, 0] >= 0).sum())}); K-colours vs PSCube.cso: max {int(ul.max())} binary16 steps, "
              f"{float((ul > 0).mean()) * 100:.3f} % of halves differ; blend vs PSResolveOIT.cso max abs {db.max():.3e}")
        for k in ("layers", "blend", "done"):
            flat[f"{name}/{k}"] = r[k].view(np.uint16) if r[k].dtype == np.float16 else r[k]
    np.savez_compressed(os.path.join(OUT, "dxil_oit.npz"), **flat)


def make_peel():
    """PSDepthPeel.cso (the K-buffer insertion: a chain of InterlockedMin over the 8 layers) on pixels where up to 12 nested
    volumes overlap, fragments fed in draw order; against the layers the oracle keeps."""
    from harness import nested_scene
    from oracle_binding import OracleCaster
    o, vp, eye = nested_scene(OracleCaster, filter_model=1)
    o.Cull(); o.RayMarchV()
    cnt, info, data, result = o.DebugOIT()
    keys = o.all_keys
    H, W = cnt.shape
    kd = Texture(np.full((8, H, W, 1), 0xffffffff, np.uint32))
    ps = shader("PSDepthPeel")
    res = Resources(uav={0: kd})
    n_frag = 0
    for py in range(0, H, 2):
        for px in range(0, W, 2):
            for k in keys[py, px]:
                if k != 0xffffffff:
                    z = np.array(k, np.uint32).view(np.float32)[()]
                    ps.run_wave([ps.lane(res, {}, inputs={0: {0: F32(px + 0.5), 1: F32(py + 0.5), 2: z, 3: F32(1)}}, outputs={})])
                    n_frag += 1
    m = np.zeros((H, W), bool); m[0::2, 0::2] = True
    got = kd.a[..., 0].transpose(1, 2, 0)                     # (H, W, 8)
    want = np.where(np.arange(8)[None, None, :] < cnt[..., None], info[..., 0], 0xffffffff)
    most = int((keys != 0xffffffff).sum(-1)[m].max())
    print(f"peel: {n_frag} fragments on {int(m.sum())} pixels, up to {most} per pixel; K-depth layers equal the oracle's: {np.array_equal(got[m], want[m])}")
    np.savez_compressed(os.path.join(OUT, "dxil_peel.npz"), layers=got, done=m)


# ------------------------------------------------------------------------------------------------ PSEnvironment
def make_env():
    """PSEnvironment.cso per pixel (the view ray through screenToWorld, one cube fetch). The cube fetch is the texture unit's:
    face selection, seamless bilinear footprint and fp32 blend as the oracle's environment pass has them."""
    import ctypes as C
    import oracle.dxil.interp as I
    from multivolumes_b200 import scene
    from oracle_binding import OracleCaster
    I.PROMOTE_HALF = True
    W, H, S = 64, 36, 16
    o = OracleCaster(filter_model=1, grid_size=16, light_grid_size=8, num_volumes=1, width=W, height=H)
    sky = scene.procedural_sky(S).astype(np.float32)
    o.SetEnvironment(sky)
    o.SetRenderTargets()
    vp, eye = scene.default_camera(W, H, eye=(14.0, 9.0, -30.0))
    o.UpdateFrame(vp, None, eye)
    o.RenderEnvironment()
    got = o.ReadFrame()
    cube = sky.astype(np.float16).astype(np.float32)             # the pass reads an RGBA16F cube
    out3 = (C.c_int * 3)()

    def sample(tex, smp, coords, offs, lod, cmp_):
        d = [F32(coords[0]), F32(coords[1]), F32(coords[2])]
        ax, ay, az = abs(d[0]), abs(d[1]), abs(d[2])
        if ax >= ay and ax >= az: face, ma = (0 if d[0] > 0 else 1), ax
        elif ay >= az: face, ma = (2 if d[1] > 0 else 3), ay
        else: face, ma = (4 if d[2] > 0 else 5), az
        im = F32(1) / ma
        p = [F32(d[k] * im) for k in range(3)]
        h = F32(0.5)
        u, v = {0: (F32(F32(-p[2] * h) + h), F32(F32(-p[1] * h) + h)), 1: (F32(F32(p[2] * h) + h), F32(F32(-p[1] * h) + h)),
                2: (F32(F32(p[0] * h) + h), F32(F32(p[2] * h) + h)), 3: (F32(F32(p[0] * h) + h), F32(F32(-p[2] * h) + h)),
                4: (F32(F32(p[0] * h) + h), F32(F32(-p[1] * h) + h)), 5: (F32(F32(-p[0] * h) + h), F32(F32(-p[1] * h) + h))}[face]
        fx, fy = _fma(u, S, -0.5), _fma(v, S, -0.5)
        i0, j0 = int(np.floor(fx)), int(np.floor(fy))
        wx, wy = F32(fx - np.floor(fx)), F32(fy - np.floor(fy))
        t = []
        for k in range(4):
            o.b.cube_resolve_texel(S, face, i0 + (k & 1), j0 + (k >> 1), out3)
            t.append(cube[out3[0], out3[2], out3[1]])
        lerp = lambda a, b, w: _fma(F32(b - a), w, a)
        return [lerp(lerp(t[0][c], t[1][c], wx), lerp(t[2][c], t[3][c], wx), wy) for c in range(3)] + [F32(0)]

    # cbPerFrame of PSEnvironment: g_eyePt (row 0), g_screenToWorld (rows 1-4, column-major)
    cb = np.zeros(20, np.float32); cb[0:3] = eye; cb[4:20] = o.ReadPerFrame()["screen_to_world"].T.reshape(16)
    res = Resources(srv={0: cube}, cbv={0: CBuffer(cb.tobytes())}, sampler={0: "linear"}, sample=sample)
    ps = shader("PSEnvironment")
    out = np.zeros((H, W, 4), np.float16)
    for y in range(H):
        for x in range(W):
            o_ = {}
            ps.run_wave([ps.lane(res, {}, inputs={0: {0: F32(x + 0.5), 1: F32(y + 0.5)}, 1: {0: F32((x + 0.5) / W), 1: F32((y + 0.5) / H)}}, outputs=o_)])
            out[y, x] = [np.float16(o_[0][k]) for k in range(4)]
    ulps = np.abs(got.view(np.int16).astype(np.int32) - out.view(np.int16).astype(np.int32))
    print(f"environment {W}x{H}: oracle vs PSEnvironment.cso: max {int(ulps.max())} binary16 steps, {float((ulps > 0).mean()) * 100:.2f} % of halves differ")
    np.savez_compressed(os.path.join(OUT, "dxil_env.npz"), rgba=out.view(np.uint16), sky=sky, view_proj=vp, eye=np.asarray(eye, np.float32))


MAKERS = {"env": make_env, "peel": make_peel, "oit": make_oit, "base_pass": make_base_pass, "init": make_init, "cull": make_cull, "post": make_post, "march_v": make_march_v, "march_l": make_march_l, "sh": make_sh}

if __name__ == "__main__":
    for n in (sys.argv[1:] or MAKERS):
        MAKERS[n]()
