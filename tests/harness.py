"""Shared scene set-up for the parity tests: drives the CPU oracle and the CUDA product through the
same MultiRayCaster calls with the same seeded inputs (the oracle is the checker, never the product)."""
import numpy as np

from multivolumes_b200 import scene

TOL_MAX_ABS = 2e-3      # BASELINE.json north_star: max-abs 2e-3 on RGBA16F
TOL_PSNR_DB = 50.0      # and PSNR >= 50 dB


def psnr(a, b, peak=None):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    mse = np.mean((a - b) ** 2)
    if mse == 0:
        return np.inf
    if peak is None:
        peak = max(float(np.abs(b).max()), 1.0)
    return 10.0 * np.log10(peak * peak / mse)


def assert_image_close(got_h, want_h, what=""):
    """got/want: float16 arrays. Bar: max-abs <= 2e-3 (relative to max(1, |want|)) and PSNR >= 50 dB."""
    g = got_h.astype(np.float32); w = want_h.astype(np.float32)
    assert np.isfinite(g).all() == np.isfinite(w).all(), what
    fin = np.isfinite(w)
    err = np.abs(g[fin] - w[fin]) / np.maximum(1.0, np.abs(w[fin]))
    assert err.size == 0 or err.max() <= TOL_MAX_ABS, f"{what}: max-abs {err.max():.3e}"
    assert psnr(g[fin], w[fin]) >= TOL_PSNR_DB, f"{what}: PSNR {psnr(g[fin], w[fin]):.1f} dB"


def rotation_y(angle):
    c, s = np.cos(angle), np.sin(angle)
    return np.array([[c, 0, -s], [0, 1, 0], [s, 0, c]], np.float64)


def rotation_xyz(rs):
    a, b, c = rs.uniform(0, 2 * np.pi, 3)
    rx = np.array([[1, 0, 0], [0, np.cos(a), np.sin(a)], [0, -np.sin(a), np.cos(a)]])
    rz = np.array([[np.cos(c), np.sin(c), 0], [-np.sin(c), np.cos(c), 0], [0, 0, 1]])
    return rx @ rotation_y(b) @ rz


def world43(scale, rot3, pos):
    m = np.zeros((4, 3), np.float32)
    m[:3, :3] = (np.eye(3) * scale) @ rot3
    m[3] = pos
    return m


def configure(c, *, mode=1, sh=False, depth=None, shadow=None, background=None, eye=(4.0, 16.0, -80.0), focus=(0.0, 0.0, 0.0),
              random_transforms=0, light_intensity=scene.LIGHT_INTENSITY, velocity=None):
    """Same calls on either backend. Returns (view_proj, eye)."""
    for i in range(c.srcs):
        c.InitVolumeData(i, mode, (0x9E3779B9 * (i + 1)) & 0xffffffff)
    c.SetLight(scene.LIGHT_PT, scene.LIGHT_COLOR, light_intensity)
    c.SetAmbient(scene.AMBIENT_COLOR, scene.AMBIENT_INTENSITY)
    c.SetVolumesWorld(20.0, (0, 0, 0))
    if random_transforms:
        rs = np.random.RandomState(random_transforms)
        for i in range(c.N):
            pos = rs.uniform(-60, 60, 3) * np.array([1, 0.3, 1])
            c.SetVolumeWorldMatrix(i, world43(rs.uniform(4, 14), rotation_xyz(rs), pos))
    vp, eye = scene.default_camera(c.W, c.H, eye=eye, focus=focus)
    if sh:
        # coefficients are an INPUT here (fixed numbers), so both backends light with identical SH
        c.SetSH(sh_coeffs())
    else:
        c.SetSH(None)
    c.SetRenderTargets(depth=depth, shadow=shadow, color=background, velocity=velocity)
    svp = scene.shadow_view_proj() if shadow is not None else None
    c.UpdateFrame(vp, svp, eye)
    return vp, eye


def sh_coeffs():
    rs = np.random.RandomState(7)
    k = rs.uniform(-0.3, 0.3, (9, 3)).astype(np.float32)
    k[0] = [2.9, 3.1, 3.6]
    return k


def checker_background(W, H):
    y, x = np.mgrid[0:H, 0:W]
    bg = np.zeros((H, W, 4), np.float16)
    bg[..., 0] = 0.2 + 0.3 * ((x // 8 + y // 8) & 1)
    bg[..., 1] = 0.35
    bg[..., 2] = 0.1 + 0.5 * (y / H)
    bg[..., 3] = 1.0
    return bg


def blob_shadow(size=64):
    """D16 shadow map: a disc of near depth (an occluder) in a far field."""
    y, x = np.mgrid[0:size, 0:size]
    d = np.full((size, size), 65535, np.uint16)
    d[(x - size * 0.45) ** 2 + (y - size * 0.5) ** 2 < (size * 0.2) ** 2] = 9000
    return d


def uv_sphere(radius=1.0, center=(0.0, 0.0, 0.0), rings=24, sectors=48):
    """Closed triangle mesh (positions (V, 3) float32, indices (3 T,) uint32) standing in for the reference's bunny.obj."""
    pos, idx = [], []
    for r in range(rings + 1):
        th = np.pi * r / rings
        for s in range(sectors):
            ph = 2.0 * np.pi * s / sectors
            pos.append((center[0] + radius * np.sin(th) * np.cos(ph), center[1] + radius * np.cos(th), center[2] + radius * np.sin(th) * np.sin(ph)))
    for r in range(rings):
        for s in range(sectors):
            a, b = r * sectors + s, r * sectors + (s + 1) % sectors
            c, d = a + sectors, b + sectors
            idx += [a, c, b, b, c, d]
    return np.asarray(pos, np.float32), np.asarray(idx, np.uint32)


def triangle_soup(count, seed, extent=30.0, size=6.0):
    """Random triangles, some of them large or crossing the camera's near plane."""
    rs = np.random.RandomState(seed)
    c = rs.uniform(-extent, extent, (count, 1, 3))
    c[:, 0, 2] = rs.uniform(-110.0, 60.0, count)       # the default camera sits at z = -80
    v = c + rs.uniform(-size, size, (count, 3, 3)) * rs.choice([0.2, 1.0, 6.0], (count, 1, 1))
    return v.reshape(-1, 3).astype(np.float32), np.arange(3 * count, dtype=np.uint32)


# ---------------------------------------------------------------- scenes of the DXIL golden vectors (oracle/dxil/make_golden.py)
DXIL_SCENES = {
    "a": dict(seed=2, grid=16, light_grid=8, n=3, W=96, H=54, ray=40, light=12),
    "b": dict(seed=6, grid=16, light_grid=8, n=2, W=96, H=54, ray=40, light=12),
    "c": dict(seed=11, grid=32, light_grid=12, n=4, W=128, H=72, ray=64, light=16),
    "d": dict(seed=4, grid=16, light_grid=8, n=3, W=96, H=54, ray=48, light=12, inside=True),     # the eye inside volume 0
}


def dxil_scene(cls, name, light_maps=True, **kw):
    """A tiny scene (depth map with occluder patches, shadow map, SH lighting, random transforms) set up on a caster of class
    `cls` up to the cull and, optionally, every light map. Returns (caster, view_proj, eye, depth, shadow)."""
    cfg = DXIL_SCENES[name]
    W, H = cfg["W"], cfg["H"]
    c = cls(grid_size=cfg["grid"], light_grid_size=cfg["light_grid"], num_volumes=cfg["n"], width=W, height=H,
            max_ray_samples=cfg["ray"], max_light_samples=cfg["light"], **kw)
    y, x = np.mgrid[0:H, 0:W]
    depth = np.where((x // 12 + y // 9) % 3 == 0, 0.9985, 1.0).astype(np.float32)
    shadow = blob_shadow(64)
    vp, eye = configure(c, sh=True, depth=depth, shadow=shadow, random_transforms=cfg["seed"], eye=cfg.get("eye", (6.0, 18.0, -62.0)))
    if cfg.get("inside"):
        # a 26-unit box around the eye whose far corner lies straight ahead (so the cull keeps it: it tests corners only)
        c.SetVolumeWorld(0, 26.0, (eye[0] + 11.14, eye[1] + 7.44, eye[2] + 6.12))
        c.UpdateFrame(vp, scene.shadow_view_proj(), eye)
    c.Cull()
    if light_maps:
        for v in range(cfg["n"]):
            c.RayMarchL(v)
    return c, vp, eye, depth, shadow


def nested_scene(cls, **kw):
    """12 concentric volumes of growing size (every pixel through the centre crosses 12 back faces: more than the 8 OIT layers)"""
    c = cls(grid_size=16, light_grid_size=8, num_volumes=12, num_volume_srcs=2, width=64, height=36, max_ray_samples=24, max_light_samples=8, **kw)
    for i in range(c.srcs):
        c.InitVolumeData(i, 1, 77 + i)
    c.SetLight(scene.LIGHT_PT, scene.LIGHT_COLOR, scene.LIGHT_INTENSITY)
    c.SetAmbient(scene.AMBIENT_COLOR, scene.AMBIENT_INTENSITY)
    for i in range(c.N):
        c.SetVolumeWorld(i, 6.0 + 1.5 * i, (0.3 * i, 0.0, 0.0))
    c.SetSH(None)
    c.SetRenderTargets()
    vp, eye = scene.default_camera(c.W, c.H, eye=(3.0, 9.0, -46.0))
    c.UpdateFrame(vp, None, eye)
    return c, vp, eye
