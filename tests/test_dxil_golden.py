"""The oracle — and, on a GPU, the product — against outputs of the reference's OWN compiled shaders.

The reference ships its shaders as DXIL (`/root/reference/Bin/*.cso`). `oracle/dxil/` disassembles them (llvmlite) and executes
them (a small LLVM-IR / DXIL interpreter with wave intrinsics) on seeded inputs; `python -m oracle.dxil.make_golden` wrote
the vectors under tests/golden/dxil_*.npz, here, once, and nothing below touches /root/reference. What each vector pins:

  dxil_cull      CSVolumeCull.cso ......... visible / cube-map lists and VolumeInfo (u32 x 4): EXACT, at BASELINE shapes
  dxil_march_v   CSRayMarchV.cso .......... cube-map texels (RGBA16F) and cube depths: BIT-EXACT (the oracle: also the fp32 values before the store)
  dxil_march_l   CSRayMarchL.cso .......... light-map voxels (R11G11B10_FLOAT values): BIT-EXACT
  dxil_oit       PSCube.cso (CubeCast and RayCast) + PSResolveOIT.cso ... K-buffer colours and the blended pixel: BIT-EXACT
                 (fragments — depth key, exit point, face uv — from the oracle's analytic rasteriser; every 3rd pixel)
  dxil_peel      PSDepthPeel.cso ... the 8 K-buffer depth layers of pixels crossed by 12 nested volumes: EXACT
  dxil_post      CSTemporalAA.cso + PSToneMap.cso ... TAA output within one binary16 step on isolated texels, RGBA8 EXACT
  dxil_init      CSInitGridData.cso, CSR32FToRGBA16F.cso ... volume texels (RGBA16F): BIT-EXACT
  dxil_env       PSEnvironment.cso ... the sky behind the volumes (RGBA16F): BIT-EXACT
  dxil_base_pass PSBasePass.cso (mesh under the volumes) ... colour (RGBA16F) and velocity: BIT-EXACT on a clip-space quad
  dxil_sh        CSSHCubeMap / CSSHSum / CSSHNormalize.cso (no HLSL in the reference) ... 9 x 3 coefficients to 1e-6 relative
                 (the summation order of a wave reduction is the hardware's)

The texture unit is not shader code: the interpreter's sampler callbacks use the oracle's filter (model 0 = exact fp32
trilinear, model 1 = the sm_100a unit the product's tex3D runs on); vectors exist for both. `min16float` arithmetic is
evaluated in fp32 with the binary16 LITERALS the DXIL holds (what a driver without fp16 ALUs executes)."""
import os

import numpy as np
import pytest

from harness import DXIL_SCENES, dxil_scene
from oracle_binding import OracleCaster

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module", autouse=True)
def _built(oracle_lib):
    return oracle_lib


def _load(name):
    return np.load(os.path.join(GOLD, name))


def _product(**kw):
    from multivolumes_b200 import MultiRayCaster
    return MultiRayCaster(**kw)


def _casters():
    """(id, factory(filter_model, **kw), filter model of its texture unit or None = any)"""
    return [pytest.param(lambda model, **kw: OracleCaster(filter_model=model, **kw), None, id="oracle"),
            pytest.param(lambda model, **kw: _product(**kw), 1, id="product", marks=pytest.mark.gpu)]


# ---------------------------------------------------------------------------------------------------------------- cull
def _cull_cases():
    g = _load("dxil_cull.npz")
    return sorted({k.split("/")[0] for k in g.files})


@pytest.mark.parametrize("make,unit", _casters())
@pytest.mark.parametrize("case", _cull_cases())
def test_cull_equals_the_reference_shader(make, unit, case):
    g = _load("dxil_cull.npz")
    G, L, N, W, H = [int(x) for x in g[f"{case}/shape"]]
    c = make(1, grid_size=G, light_grid_size=L, num_volumes=N, num_volume_srcs=int(g[f"{case}/srcs"]), width=W, height=H,
             max_ray_samples=int(g[f"{case}/max_ray_samples"]))
    po = g[f"{case}/per_object"]
    c.SetVolumeWorldMatrices(po[:, 44:56].reshape(N, 4, 3))
    c.UpdateFrame(g[f"{case}/view_proj"], None, g[f"{case}/eye"])
    assert np.array_equal(c.ReadPerObject().view(np.uint32), po.view(np.uint32))          # same inputs as the shader saw
    c.Cull()
    vis, cub, att = c.ReadVisible(), c.ReadCubeVolumes(), c.ReadAttribs()
    assert np.array_equal(np.sort(vis), np.sort(g[f"{case}/visible"]))                      # Append order is the hardware's
    assert np.array_equal(np.sort(cub), np.sort(g[f"{case}/cube_volumes"]))
    info = g[f"{case}/volume_info"]
    for v in vis:
        assert np.array_equal(att[v], info[v].astype(np.uint16)), (v, att[v], info[v])


# ---------------------------------------------------------------------------------------------------------------- marches
@pytest.mark.parametrize("make,unit", _casters())
@pytest.mark.parametrize("model", [0, 1])
@pytest.mark.parametrize("name", sorted(DXIL_SCENES))
def test_view_march_equals_the_reference_shader(make, unit, model, name):
    if unit is not None and unit != model:
        pytest.skip("the product's texture unit is the hardware's (model 1)")
    g = _load("dxil_march_v.npz")
    c, vp, eye, depth, shadow = dxil_scene(lambda **kw: make(model, **kw), name)
    if unit is None:
        c.DebugF32(True)
    c.RayMarchV()
    cubes = g[f"f{model}/{name}/cubes"]
    assert np.array_equal(np.sort(c.ReadCubeVolumes()), cubes) and len(cubes) > 0
    rays = 0
    for v in cubes:
        k = f"f{model}/{name}/v{int(v)}"
        mip = int(g[k + "/mip"])
        rgba, dep = c.ReadCubeMap(int(v), mip)
        want_d = g[k + "/depth"]
        mask = want_d >= 0                                        # texels the shader wrote (visible faces, rays that hit)
        assert np.array_equal(np.asarray(dep)[mask].view(np.uint32), want_d[mask].view(np.uint32))
        assert np.array_equal(np.asarray(rgba).view(np.uint16)[mask], g[k + "/rgba"][mask]), (name, int(v))
        if unit is None:                                          # the oracle also keeps the fp32 scatter: equal before the RGBA16F store too
            s = c.G >> mip
            assert np.array_equal(c.DebugF32(True)[0][int(v), :, :s, :s].view(np.uint32)[mask], g[k + "/rgba_f32"].view(np.uint32)[mask])
        rays += int(mask.sum())
    assert rays >= 48


@pytest.mark.parametrize("make,unit", _casters())
@pytest.mark.parametrize("model", [0, 1])
@pytest.mark.parametrize("name", sorted(DXIL_SCENES))
def test_light_march_equals_the_reference_shader(make, unit, model, name):
    if unit is not None and unit != model:
        pytest.skip("the product's texture unit is the hardware's (model 1)")
    g = _load("dxil_march_l.npz")
    c, vp, eye, depth, shadow = dxil_scene(lambda **kw: make(model, **kw), name, light_maps=False)
    checked = 0
    for v in c.ReadVisible():
        c.RayMarchL(int(v))
        got = np.asarray(c.ReadLightMap(int(v))).view(np.float16)[..., :3].astype(np.float32)
        want = g[f"f{model}/{name}/v{int(v)}"]
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (name, int(v), int((got != want).sum()))
        checked += 1
    assert checked >= 2


# ---------------------------------------------------------------------------------------------------------------- TAA + tone map
@pytest.mark.parametrize("make,unit", _casters())
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_postprocess_against_the_reference_shaders(make, unit, seed):
    g = _load("dxil_post.npz")
    cur, hist, vel = [g[f"s{seed}/{k}"].view(np.float16) for k in ("current", "history", "velocity")]
    H, W = cur.shape[:2]
    c = make(0, grid_size=32, light_grid_size=16, num_volumes=1, width=W, height=H)
    c.SetRenderTargets(color=hist); c.Postprocess(False)                      # history := the given image (TAA off copies)
    c.SetRenderTargets(color=cur, velocity=vel if vel.any() else None); c.Postprocess(True)
    taa, rgba8 = c.ReadPost()
    want = g[f"s{seed}/taa_f32"].view(np.float16)
    got = np.asarray(taa).view(np.float16)
    ulps = np.abs(got.view(np.int16).astype(np.int32) - want.view(np.int16).astype(np.int32))
    assert ulps.max() <= 1 and (ulps > 0).mean() < 0.01, (int(ulps.max()), float((ulps > 0).mean()))
    assert np.array_equal(np.asarray(rgba8), g[f"s{seed}/rgba8_f32"])
    # evaluated in binary16 instead (a driver WITH fp16 ALUs) the reference itself moves by far more than that
    f16 = g[f"s{seed}/taa_f16"].view(np.float16).astype(np.float32)
    assert np.abs(f16 - want.astype(np.float32)).max() > 20 * np.abs(got.astype(np.float32) - want.astype(np.float32)).max()


# ---------------------------------------------------------------------------------------------------------------- SH projection
@pytest.mark.parametrize("make,unit", _casters())
@pytest.mark.parametrize("name", ["noise8", "sky16"])
def test_sh_projection_against_the_reference_kernels(make, unit, name):
    g = _load("dxil_sh.npz")
    c = make(1, grid_size=32, light_grid_size=16, num_volumes=1, width=64, height=48)
    got = c.TransformSH(g[name + "/cube"])
    want = g[name + "/coeffs"]
    assert np.abs(got - want).max() <= 2e-6 * np.abs(want).max()


# ---------------------------------------------------------------------------------------------------------------- ingest
@pytest.mark.parametrize("make,unit", _casters())
@pytest.mark.parametrize("G", [16, 24])
def test_volume_init_and_conversion_equal_the_reference_shaders(make, unit, G):
    g = _load("dxil_init.npz")
    c = make(1, grid_size=G, light_grid_size=8, num_volumes=1, width=64, height=48)
    c.InitVolumeData(0, 0, 0)
    assert np.array_equal(np.asarray(c.ReadVolume(0)).view(np.uint16), g[f"g{G}/rgba"])
    c.LoadVolumeData(0, g[f"g{G}/density"])
    assert np.array_equal(np.asarray(c.ReadVolume(0)).view(np.uint16), g[f"g{G}/converted"])


# ---------------------------------------------------------------------------------------------------------------- mesh base pass
@pytest.mark.parametrize("make,unit", _casters())
@pytest.mark.parametrize("use_sh", [0, 1])
def test_base_pass_equals_the_reference_pixel_shader(make, unit, use_sh):
    from harness import sh_coeffs
    g = _load("dxil_base_pass.npz")
    want = g[f"sh{use_sh}/rgba"]
    H, W = want.shape[:2]
    c = make(1, grid_size=32, light_grid_size=16, num_volumes=1, width=W, height=H)
    c.SetLight(g["light"], g["light_rgbi"][:3], float(g["light_rgbi"][3])); c.SetAmbient(g["ambient_rgbi"][:3], float(g["ambient_rgbi"][3]))
    c.SetSH(sh_coeffs() if use_sh else None)
    c.SetMesh(g["mesh"], np.arange(6, dtype=np.uint32)); c.SetMeshWorld(1.0, (0, 0, 0))
    c.RenderMesh(np.eye(4, dtype=np.float32), g["eye"])
    assert np.array_equal(np.asarray(c.ReadFrame()).view(np.uint16), want)
    assert np.array_equal(np.asarray(c.ReadVelocity()).view(np.uint16) & 0x7fff, g[f"sh{use_sh}/velocity"] & 0x7fff)   # +-0


# ---------------------------------------------------------------------------------------------------------------- OIT
def _oit_scene(make, name):
    from harness import nested_scene
    if name != "nested":
        return dxil_scene(lambda **kw: make(1, **kw), name)[0]
    c, vp, eye = nested_scene(lambda **kw: make(1, **kw))
    c.Cull()
    for v in range(c.N):
        c.RayMarchL(v)
    return c


@pytest.mark.parametrize("name", sorted(DXIL_SCENES) + ["nested"])
def test_oit_layers_and_blend_equal_the_reference_pixel_shaders(name):
    """oracle only (the product keeps its K-buffer in registers): per-layer colours as PSCube.cso stores them, the blend as
    PSResolveOIT.cso returns it"""
    g = _load("dxil_oit.npz")
    c = _oit_scene(lambda model, **kw: OracleCaster(filter_model=model, **kw), name)
    c.RayMarchV()
    cnt, info, data, result = c.DebugOIT()
    if name == "nested":                                      # PSDepthPeel.cso: the 8 nearest of up to 12 fragments, ascending
        p = _load("dxil_peel.npz")
        want_keys = np.where(np.arange(8)[None, None, :] < cnt[..., None], info[..., 0], 0xffffffff)
        assert np.array_equal(want_keys[p["done"]], p["layers"][p["done"]]) and cnt.max() == 8 and (c.all_keys != 0xffffffff).sum(-1).max() == 12
    m = g[f"{name}/done"]
    assert m.sum() >= 30 and np.array_equal(cnt[m] > 0, np.ones(int(m.sum()), bool))
    want = g[f"{name}/layers"]
    got = np.zeros_like(want)
    for py, px in np.argwhere(m):
        for l in range(int(cnt[py, px])):
            if info[py, px, l, 3]:
                got[py, px, l] = data[py, px, l, 5:9].astype(np.float16).view(np.uint16)
    assert np.array_equal(got[m], want[m])
    assert np.array_equal(result[m].view(np.uint32), g[f"{name}/blend"][m].view(np.uint32))


@pytest.mark.parametrize("make,unit", _casters())
@pytest.mark.parametrize("name", sorted(DXIL_SCENES) + ["nested"])
def test_resolved_frame_equals_the_reference_pixel_shaders(make, unit, name):
    """the frame after the resolve (no background: the render-target blend adds nothing) against PSResolveOIT.cso's output"""
    g = _load("dxil_oit.npz")
    c = _oit_scene(make, name)
    c.RayMarchV(); c.ResolveOIT()
    m = g[f"{name}/done"]
    frame = np.asarray(c.ReadFrame()).view(np.uint16)
    assert np.array_equal(frame[m], g[f"{name}/blend"][m].astype(np.float16).view(np.uint16))


# ---------------------------------------------------------------------------------------------------------------- environment
@pytest.mark.parametrize("make,unit", _casters())
def test_environment_equals_the_reference_pixel_shader(make, unit):
    g = _load("dxil_env.npz")
    want = g["rgba"]
    H, W = want.shape[:2]
    c = make(1, grid_size=16, light_grid_size=8, num_volumes=1, width=W, height=H)
    c.SetEnvironment(g["sky"])
    c.SetRenderTargets()
    c.UpdateFrame(g["view_proj"], None, g["eye"])
    c.RenderEnvironment()
    got = np.asarray(c.ReadFrame()).view(np.uint16)
    assert (want[..., :3] != 0).mean() > 0.9 and np.array_equal(got, want)
