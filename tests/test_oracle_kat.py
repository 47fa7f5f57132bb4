"""Known-answer tests that pin the CPU oracle (the reference ships no tests, golden vectors or CPU
path — SURVEY.md §4 — so the oracle is pinned by closed forms derived from the shaders themselves
and by the B200 texture-unit dump in tests/golden/b200_tex_probe.npz)."""
import os

import numpy as np
import pytest

from oracle_binding import OracleCaster, oracle_binding
from multivolumes_b200 import scene

GOLD = os.path.join(os.path.dirname(__file__), "golden", "b200_tex_probe.npz")


@pytest.fixture(scope="module", autouse=True)
def _built(oracle_lib):
    return oracle_lib


def _mk(**kw):
    d = dict(grid_size=32, light_grid_size=16, num_volumes=1, width=64, height=36)
    d.update(kw)
    return OracleCaster(**d)


def test_fp16_conversion_matches_numpy():
    b = oracle_binding()
    rs = np.random.RandomState(1)
    vals = np.concatenate([rs.uniform(-70000, 70000, 2000), rs.uniform(-1, 1, 2000) * 1e-4, rs.uniform(-1, 1, 2000) * 1e-7,
                           [0.0, 65504.0, 65519.9, 65520.0, 1e9, 5.96e-8, 2.98e-8, 2.9802322e-8, 2.99e-8, 6.1e-5]]).astype(np.float32)
    with np.errstate(over="ignore"):
        want = vals.astype(np.float16).view(np.uint16)
    got = np.array([b.f32_to_f16(float(v)) for v in vals], np.uint16)
    assert np.array_equal(got, want)
    allh = np.arange(65536, dtype=np.uint16)
    back = np.array([b.f16_to_f32(int(h)) for h in allh[::7]], np.float32)
    assert np.array_equal(np.nan_to_num(back, nan=-1), np.nan_to_num(allh[::7].view(np.float16).astype(np.float32), nan=-1))


def test_r11g11b10_quantisation():
    b = oracle_binding()
    # 6-bit mantissa: 1 + k/64 exactly representable; halfway cases round to even
    assert b.quantize_r11(1.0) == 1.0
    assert b.quantize_r11(1.0 + 1 / 64) == np.float32(1.0 + 1 / 64)
    assert b.quantize_r11(1.0 + 1 / 128) == 1.0                      # tie -> even (mantissa 0)
    assert b.quantize_r11(1.0 + 3 / 128) == np.float32(1.0 + 2 / 64)  # tie -> even (mantissa 2)
    assert b.quantize_b10(1.0 + 1 / 32) == np.float32(1.0 + 1 / 32)
    assert b.quantize_b10(1.0 + 1 / 64) == 1.0
    assert b.quantize_r11(-3.0) == 0.0
    assert b.quantize_r11(1e9) == 65024.0 and b.quantize_b10(1e9) == 64512.0
    # every quantised value survives a trip through fp16 (the CUDA light map stores it in RGBA16F)
    for v in np.random.RandomState(0).uniform(0, 40, 200):
        q = b.quantize_r11(float(v))
        assert np.float16(q) == q


def test_sampler_matches_b200_texture_unit_dump():
    """MODEL_SM100 reproduces the hardware trilinear filter bit for bit on the committed probe dump."""
    g = np.load(GOLD)
    c = _mk(grid_size=32, filter_model=1)
    c.LoadVolumeData(0, g["rand32_tex"])
    got = c.SampleVolume(0, g["rand32_coords"])
    assert np.array_equal(got, g["rand32_out"])
    # exact-weights model differs (this is the intrinsic hardware filter error, not a bug)
    c0 = _mk(grid_size=32, filter_model=0)
    c0.LoadVolumeData(0, g["rand32_tex"])
    err = np.abs(c0.SampleVolume(0, g["rand32_coords"][:2000]) - g["rand32_out"][:2000])
    assert 1e-4 < err.max() < 0.05


@pytest.mark.parametrize("n", [96, 100])
def test_sampler_non_power_of_two_sizes(n):
    g = np.load(GOLD)
    z, y, x = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    tex = np.stack([x & 1, y & 1, z & 1, x], -1).astype(np.float16)
    c = _mk(grid_size=n, filter_model=1)
    c.LoadVolumeData(0, tex)
    got = c.SampleVolume(0, g[f"np2_{n}_coords"][:4096])
    assert np.array_equal(got, g[f"np2_{n}_out"][:4096])


def test_sampler_onehot_weights():
    g = np.load(GOLD)
    n = 4
    coords, want = g["onehot_coords"][:4096], g["onehot_w"][:4096]
    got = np.zeros((len(coords), 8), np.float32)
    for layer in range(2):
        tex = np.zeros((n, n, n, 4), np.float16)
        z = 1 + layer
        tex[z, 1, 1, 0] = 1; tex[z, 1, 2, 1] = 1; tex[z, 2, 1, 2] = 1; tex[z, 2, 2, 3] = 1
        c = _mk(grid_size=16, filter_model=1)   # grid must allow 5 mips; emulate the 4^3 probe inside a 16^3? no: use direct 4^3 below
        c.close()
    # the 4^3 probe texture cannot be a caster volume (needs >= 16 for 5 cube mips), so embed it: a 16^3
    # texture sampled at u' = u * 4 / 16 sees the same texel neighbourhood and fractions.
    for layer in range(2):
        tex = np.zeros((16, 16, 16, 4), np.float16)
        z = 1 + layer
        tex[z, 1, 1, 0] = 1; tex[z, 1, 2, 1] = 1; tex[z, 2, 1, 2] = 1; tex[z, 2, 2, 3] = 1
        c = _mk(grid_size=16, filter_model=1)
        c.LoadVolumeData(0, tex)
        got[:, layer * 4:(layer + 1) * 4] = c.SampleVolume(0, coords * np.float32(0.25))
    assert np.array_equal(np.round(got * 256).astype(np.uint16), want)


def _setup_single(c, eye=(0.0, 0.0, -60.0)):
    vp, eye = scene.default_camera(c.W, c.H, eye=eye)
    c.SetLight(scene.LIGHT_PT, scene.LIGHT_COLOR, 1.0)
    c.SetAmbient((1.0, 1.0, 1.0), 1.0)
    c.SetVolumeWorld(0, 20.0, (0, 0, 0))
    c.UpdateFrame(vp, None, eye)
    return vp, eye


def test_empty_volume_gives_zero_cube_map_and_untouched_frame():
    c = _mk(width=96, height=54)
    c.LoadVolumeData(0, np.zeros((32, 32, 32, 4), np.float16))
    _setup_single(c)
    bg = np.full((54, 96, 4), 0.25, np.float16)
    c.SetRenderTargets(color=bg)
    c.Render()
    att = c.ReadAttribs()[0]
    assert len(c.ReadVisible()) == 1
    rgba, _ = c.ReadCubeMap(0, int(att[0]))
    assert not rgba.view(np.uint16).any()
    assert np.array_equal(c.ReadFrame().view(np.uint16), bg.view(np.uint16))


def test_uniform_density_closed_form():
    """Uniform density rho, white colour: every step has dDensity = 0 after the first sample, so the
    adaptive factor is a closed form and the accumulated alpha follows the discrete recurrence
    A_{k+1} = A_k + 0.8 rho (1 - A_k) per sample (CSRayMarch.hlsl:117-153)."""
    rho = 0.05
    c = _mk(grid_size=32, width=640, height=360, filter_model=0)
    tex = np.zeros((32, 32, 32, 4), np.float16); tex[..., :3] = 1.0; tex[..., 3] = rho
    c.LoadVolumeData(0, tex)
    _setup_single(c, eye=(0.0, 0.0, -40.0))
    c.Cull()
    c.RayMarchL(0)
    # light map must be the constant lightColor*1 + ambient (no SH, shadow map absent)
    lm = c.ReadLightMap(0).astype(np.float32)
    c.RayMarchV()
    att = c.ReadAttribs()[0]
    mip, smp, mask = int(att[0]), int(att[1]), int(att[2])
    assert mask & 0x8000
    rgba, depth = c.ReadCubeMap(0, mip)
    s = 32 >> mip
    # centre texel of the +Z face (face 4): ray straight through the box along +z, chord length 2
    a = rgba[4, s // 2, s // 2].astype(np.float32)
    rho16 = float(np.float16(rho))
    step = np.float32(2 * np.sqrt(np.float32(3.0))) / np.float32(smp)
    # replay the recurrence in fp32
    A = np.float32(0); t = np.float32(0); prev = np.float32(0); n = 0
    while n < smp:
        if t > 2.0 + 1e-3: break
        transm = np.float32(1) - A
        d = np.float32(rho16) - prev
        fe = min(np.float32(1 / 256) / abs(d), 2.0) if d != 0 else 2.0
        new = step * max(np.float32(1.5) * np.float32(fe) * min(1 - rho16, 1.0) * (1 - transm), 1.0)
        prev = np.float32(rho16)
        A = A + np.float32(rho16) * np.float32(0.8) * transm
        t = t + np.float32(new); n += 1
    assert abs(a[3] - A) < 2e-3 * max(1.0, A) + 1e-3
    assert np.all(depth[4] == 1.0)


def test_sphere_symmetry_between_mirrored_faces():
    """The procedural density is symmetric in x, so with the eye on the z axis the -X and +X
    interior faces hold mirror-image alpha."""
    c = _mk(grid_size=32, width=640, height=360, filter_model=0)
    c.InitVolumeData(0, 0, 0)
    _setup_single(c, eye=(0.0, 0.0, -30.0))
    c.SetAmbient((1.0, 1.0, 1.0), 1.0)
    c.Cull(); c.RayMarchL(0); c.RayMarchV()
    att = c.ReadAttribs()[0]
    rgba, _ = c.ReadCubeMap(0, int(att[0]))
    a = rgba[..., 3].astype(np.float32)
    # +X face (0): u = -z ; -X face (1): u = +z  -> mirror in u
    assert np.abs(a[0] - a[1][:, ::-1]).max() < 2e-3


def test_eye_inside_box_sees_six_faces():
    c = _mk()
    c.InitVolumeData(0, 0, 0)
    vp, eye = scene.default_camera(c.W, c.H, eye=(0.5, 0.2, -1.0), focus=(0, 0, 10))
    c.SetVolumeWorld(0, 20.0, (0, 0, 0))
    c.UpdateFrame(vp, None, eye)
    c.Cull()
    vis = c.ReadVisible()
    if len(vis):
        assert int(c.ReadAttribs()[0][2]) & 0x3f == 0x3f


def test_volume_behind_depth_gives_zero_samples():
    c = _mk(width=640, height=360)
    c.InitVolumeData(0, 0, 0)
    _setup_single(c, eye=(0.0, 0.0, -40.0))
    c.SetRenderTargets(depth=np.zeros((360, 640), np.float32))   # everything occluded at the near plane
    c.Cull(); c.RayMarchV()
    st = c.GetStats()
    # tMax <= 0: exactly one sample per ray (the loop tests t > tMax after the first step)
    assert st["view_samples"] <= st["view_rays"]


def test_sh_constant_radiance():
    """Constant radiance c projects to L00 = 2 sqrt(pi) c, all other coefficients ~0, and the
    irradiance evaluates to pi c (SHIrradianceTypeless.hlsli:16-37)."""
    c = _mk()
    cube = np.full((6, 32, 32, 3), 0.75, np.float32)
    sh = c.TransformSH(cube)
    assert np.allclose(sh[0], 2 * np.sqrt(np.pi) * 0.75, rtol=2e-3)
    assert np.abs(sh[1:]).max() < 5e-3
    b = oracle_binding()
    out = np.zeros(4, np.float32)
    import ctypes as C
    n = np.array([0.3, -0.5, 0.81], np.float32); n /= np.linalg.norm(n)
    b.eval_sh_irradiance(sh.ctypes.data, n.ctypes.data_as(C.POINTER(C.c_float)), out.ctypes.data_as(C.POINTER(C.c_float)))
    assert np.allclose(out[:3], np.pi * 0.75, rtol=5e-3)


def test_cull_default_scene_counts():
    """Default 2x2 grid, start-up camera: all four volumes visible, order ascending (SURVEY.md §8d cfg 1)."""
    c = OracleCaster(grid_size=128, num_volumes=4, width=1280, height=720)
    vp, eye = scene.default_camera(1280, 720)
    c.UpdateFrame(vp, None, eye)
    c.Cull()
    assert list(c.ReadVisible()) == [0, 1, 2, 3]
    att = c.ReadAttribs()
    assert np.all(att[:, 1] <= 256) and np.all(att[:, 0] < 5)
    assert list(att[:, 3]) == [0, 1, 2, 3]
    # cube-map list is the subset with bit 15 set
    assert list(c.ReadCubeVolumes()) == [i for i in range(4) if att[i, 2] & 0x8000]


def test_set_volumes_world_grid_rule():
    """MultiRayCaster.cpp:277-295: spacing 1.5 * size, row-major in x then z, scale = size / 2."""
    c = OracleCaster(grid_size=16, num_volumes=16, width=64, height=36)
    c.SetVolumesWorld(20.0, (0, 0, 0))
    vp, eye = scene.default_camera(64, 36)
    c.UpdateFrame(vp, None, eye)
    world = c.ReadPerObject()[:, 44:56].reshape(16, 4, 3)
    assert np.allclose(world[:, 0, 0], 10.0) and np.allclose(world[:, 1, 1], 10.0)
    assert np.allclose(world[0, 3], [-45, 0, -45]) and np.allclose(world[5, 3], [-15, 0, -15]) and np.allclose(world[15, 3], [45, 0, 45])


def test_work_graph_order_lights_from_previous_visible_list():
    """Render(..., useWorkGraph = true), MultiRayCaster.cpp:358-362: rayMarchL runs before the graph that culls, so the
    light volume of frame f is visible_{f-1}[f % |visible_{f-1}|] (CSRayMarchL.hlsl:29-33), f % N on the first frame;
    everything downstream of the cull (lists, cube maps) is that of the plain order."""
    kw = dict(grid_size=16, light_grid_size=8, num_volumes=9, num_volume_srcs=1, width=160, height=90)
    wg, plain = OracleCaster(**kw), OracleCaster(**kw)
    for c in (wg, plain):
        c.InitVolumeData(0, 1, 5)
        c.SetVolumesWorld(20.0, (0, 0, 0))
        c.SetRenderTargets()
    prev_visible = np.zeros(0, np.uint32)
    for f in range(5):
        vp, eye = scene.default_camera(160, 90, eye=(4.0 + 25 * f, 16.0, -80.0 + 20 * f), focus=(12.0 * f, 0, 0))
        for c in (wg, plain):
            c.UpdateFrame(vp, None, eye)
        wg.Render(use_work_graph=True)
        plain.Render()
        want = int(prev_visible[f % len(prev_visible)]) if len(prev_visible) else f % 9
        assert wg.GetStats()["light_volume"] == want, (f, wg.GetStats()["light_volume"], want)
        vis = plain.ReadVisible()
        assert np.array_equal(wg.ReadVisible(), vis)
        assert plain.GetStats()["light_volume"] == int(vis[f % len(vis)])
        prev_visible = vis
    assert len(set(map(len, [prev_visible]))) == 1 and len(prev_visible) > 0
