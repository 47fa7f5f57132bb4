"""CPU checks of the depth-input producer: the OBJ import of the product library (host only, no
device call) against a plain Python restatement of XUSGObjLoader::Import(forDX = true), and known-answer tests of the
oracle's rasteriser (Direct3D rules: pixel centres, top-left fill rule, affine z, near-plane clip, LESS against 1.0)."""
import os

import numpy as np
import pytest

from harness import triangle_soup, uv_sphere
from multivolumes_b200 import scene
from oracle_binding import OracleCaster

REF_BUNNY = "/root/reference/Bin/Assets/bunny.obj"


@pytest.fixture(scope="module", autouse=True)
def _built(oracle_lib, product_lib):
    return oracle_lib


def _python_obj(path):
    pos, idx = [], []
    for line in open(path):
        t = line.split("#")[0].split()
        if not t:
            continue
        if t[0] == "v":
            x, y, z = map(float, t[1:4])
            pos.append((x, y, -z))                       # XUSGObjLoader.cpp:198
        elif t[0] == "f":
            c = []
            for w in t[1:]:
                v = int(w.split("/")[0])
                c.append(v - 1 if v > 0 else len(pos) + v)
            for k in range(1, len(c) - 1):
                idx += [c[0], c[k], c[k + 1]]
    return np.asarray(pos, np.float32).reshape(-1, 3), np.asarray(idx[::-1], np.uint32)   # :227 reverse(m_indices)


def test_obj_import_matches_python_restatement(tmp_path):
    from multivolumes_b200 import parse_obj
    p = tmp_path / "m.obj"
    p.write_text("# comment\nv 0 0 0\nv 1 0 0.5\nv 1 1 -2\nv 0 1 3.25\nvn 0 0 1\nvt 0 0\n"
                 "f 1 2 3\nf 1/1/1 3/1/1 4/1/1\nf -4 -3 -2 -1\nv 2 2 2\nf 5//1 1//1 2//1\n")
    pos, idx = parse_obj(str(p))
    wp, wi = _python_obj(str(p))
    assert np.array_equal(pos, wp) and np.array_equal(idx, wi)
    assert idx.shape == (15,) and pos.shape == (5, 3) and pos[1, 2] == -0.5


def test_obj_import_rejects_bad_files(tmp_path):
    from multivolumes_b200 import parse_obj
    bad = tmp_path / "bad.obj"
    bad.write_text("v 0 0 0\nv 1 0 0\nf 1 2 7\n")
    with pytest.raises(RuntimeError):
        parse_obj(str(bad))
    with pytest.raises(RuntimeError):
        parse_obj(str(tmp_path / "missing.obj"))


@pytest.mark.skipif(not os.path.exists(REF_BUNNY), reason="the reference tree is not present on this machine")
def test_reference_bunny_imports():
    from multivolumes_b200 import parse_obj
    pos, idx = parse_obj(REF_BUNNY)
    assert pos.shape == (34835, 3) and idx.shape == (69666 * 3,)
    wp, wi = _python_obj(REF_BUNNY)
    assert np.array_equal(pos, wp) and np.array_equal(idx, wi)


def _oracle(w=64, h=48):
    return OracleCaster(filter_model=1, grid_size=32, light_grid_size=16, num_volumes=1, width=w, height=h)


def _ndc_mesh(o, tris, w=64, h=48):
    """Rasterise triangles given directly in clip space (w = 1): view_proj = identity, mesh world = identity."""
    pos = np.asarray(tris, np.float32).reshape(-1, 3)
    o.SetMesh(pos, np.arange(pos.shape[0], dtype=np.uint32))
    o.SetMeshWorld(1.0, (0, 0, 0))
    o.RenderMeshDepth(np.eye(4, dtype=np.float32))
    return o.ReadDepth()[0]


def test_raster_full_screen_quad_constant_depth():
    o = _oracle()
    d = _ndc_mesh(o, [(-1, -1, .25), (1, -1, .25), (1, 1, .25), (-1, -1, .25), (1, 1, .25), (-1, 1, .25)])
    assert np.all(d == np.float32(0.25))


def test_raster_shared_edge_is_covered_exactly_once():
    # two triangles sharing the diagonal of a quad whose corners sit on pixel centres: with the top-left rule every pixel
    # of the quad belongs to exactly one of them, in either winding
    W, H = 64, 48
    def px(x, y, z):   # pixel-centre (x + 0.5, y + 0.5) -> NDC
        return ((x + 0.5) / W * 2 - 1, 1 - (y + 0.5) / H * 2, z)
    a, b, c, d4 = px(8, 6, .5), px(40, 6, .5), px(40, 30, .5), px(8, 30, .5)
    for t1, t2 in (((a, b, c), (a, c, d4)), ((a, c, b), (a, d4, c))):
        d1 = _ndc_mesh(_oracle(), t1) < 1.0
        d2 = _ndc_mesh(_oracle(), t2) < 1.0
        assert not np.any(d1 & d2)
        both = d1 | d2
        ys, xs = np.nonzero(both)
        # left and top edges are in, right and bottom edges are out
        assert xs.min() == 8 and xs.max() == 39 and ys.min() == 6 and ys.max() == 29
        assert both[6:30, 8:40].all()


def test_raster_affine_depth_and_less_test():
    W, H = 64, 48
    o = _oracle()
    near = [(-1, -1, .2), (1, -1, .6), (1, 1, .6), (-1, -1, .2), (1, 1, .6), (-1, 1, .2)]   # z = 0.4 + 0.2 x
    far = [(-1, -1, .5), (1, -1, .5), (1, 1, .5), (-1, -1, .5), (1, 1, .5), (-1, 1, .5)]
    d = _ndc_mesh(o, near + far)
    x = ((np.arange(W) + 0.5) / W * 2 - 1)[None, :].repeat(H, 0)
    want = np.minimum(0.4 + 0.2 * x, 0.5)
    assert np.abs(d - want).max() < 2e-7
    d2 = _ndc_mesh(_oracle(), far + near)                        # order of the triangles does not matter
    assert np.array_equal(d, d2)


def test_raster_near_plane_clip_and_depth_clip():
    # a triangle reaching behind the near plane (z < 0) is cut at z = 0; one beyond the far plane is rejected per pixel
    d = _ndc_mesh(_oracle(), [(-1, -1, -0.5), (1, -1, -0.5), (0, 1, 0.5)])
    assert (d < 1.0).any() and d.min() >= 0.0
    covered_rows = np.nonzero((d < 1.0).any(axis=1))[0]
    assert covered_rows.max() <= 24                               # the part with z < 0 (lower half) is gone
    d = _ndc_mesh(_oracle(), [(-1, -1, 1.5), (1, -1, 1.5), (0, 1, 1.5)])
    assert np.all(d == 1.0)


def test_mesh_depth_and_light_view_projection_of_a_sphere():
    W, H = 320, 180
    o = OracleCaster(filter_model=1, grid_size=32, light_grid_size=16, num_volumes=1, width=W, height=H)
    pos, idx = uv_sphere(radius=5.0)
    o.SetMesh(pos, idx)
    o.SetMeshWorld(1.8, (0.0, -9.0, 0.0))                         # m_meshPosScale, MultiVolumes.cpp:46
    o.SetLight(scene.LIGHT_PT, scene.LIGHT_COLOR, scene.LIGHT_INTENSITY)
    vp, eye = scene.default_camera(W, H)
    svp = o.RenderMeshDepth(vp)
    want_svp = scene.shadow_view_proj(scene.LIGHT_PT, scene_size=10.0 * 1.8)
    assert np.abs(svp - want_svp).max() < 1e-6
    depth, shadow = o.ReadDepth()
    analytic = scene.sphere_depth(W, H, vp, center=(0.0, -9.0, 0.0), radius=9.0)
    inside = (depth < 1.0) & (analytic < 1.0)
    assert inside.sum() > 500 and (depth < 1.0).sum() <= (analytic < 1.0).sum()     # the faceted sphere lies inside the true one
    assert np.abs(depth[inside] - analytic[inside]).max() < 2e-3
    assert shadow.shape == (1024, 1024) and shadow.min() < 65535 and shadow.max() == 65535


def test_raster_order_independence_on_a_soup():
    pos, idx = triangle_soup(300, seed=3)
    vp, _ = scene.default_camera(160, 90)
    out = []
    for perm_seed in (None, 1):
        o = OracleCaster(filter_model=1, grid_size=32, light_grid_size=16, num_volumes=1, width=160, height=90)
        tri = idx.reshape(-1, 3)
        if perm_seed is not None:
            tri = tri[np.random.RandomState(perm_seed).permutation(tri.shape[0])]
        o.SetMesh(pos, tri.reshape(-1))
        o.RenderMeshDepth(vp)
        out.append(o.ReadDepth())
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])
    assert (out[0][0] < 1.0).sum() > 1000


# ---------------------------------------------------------------- the shaded base pass (ObjectRenderer::Render)
def _expected_shade(N, ws, eye, light_pt, light_rgbi, ambient_rgbi, sh=None):
    """PSBasePass.hlsl:94-153 in float64 for unit normal(s) N (..., 3) at world positions ws (..., 3); unshadowed."""
    N = np.asarray(N, np.float64); ws = np.asarray(ws, np.float64)
    L = np.asarray(light_pt, np.float64); L = L / np.linalg.norm(L)
    V = np.asarray(eye, np.float64) - ws; V = V / np.linalg.norm(V, axis=-1, keepdims=True)
    H = V + L; H = H / np.linalg.norm(H, axis=-1, keepdims=True)
    sat = lambda x: np.clip(x, 0.0, 1.0)
    NoL, NoH, NoV = sat((N * L).sum(-1)), sat((N * H).sum(-1)), sat((N * V).sum(-1))
    brdf = np.array([1.0, 0.6, 0.2]) / np.pi
    light = np.asarray(light_rgbi[:3], np.float64) * light_rgbi[3]
    if sh is None:
        amb = np.asarray(ambient_rgbi[:3], np.float64) * ambient_rgbi[3] * (0.5 + 0.5 * (N[..., 1:2] * 0.5 + 0.5))
    else:                                                    # SHIrradianceTypeless.hlsli:16-37
        c1, c2, c3, c4 = 0.429043, 0.511664, 0.247708, 0.886227
        x, y, z = -N[..., 0:1], -N[..., 1:2], N[..., 2:3]
        s = np.asarray(sh, np.float64).reshape(9, 3)
        amb = np.maximum(0.0, c1 * (x * x - y * y) * s[8] + c3 * (3 * z * z - 1) * s[6] + c4 * s[0]
                         + 2 * c1 * (s[4] * x * y + s[7] * x * z + s[5] * y * z) + 2 * c2 * (s[3] * x + s[1] * y + s[2] * z))
    fres = (1.0 - NoV) ** 5; fres = fres + (1.0 - fres) * 0.08
    spec = (NoH ** 64 * fres)[..., None]
    return (brdf * NoL[..., None] + spec) * light + brdf * amb


@pytest.mark.parametrize("facing,use_sh", [(True, False), (False, False), (True, True)])
def test_base_pass_of_a_clip_space_quad_matches_closed_form(facing, use_sh):
    """A screen-filling quad given in clip space (identity matrices): constant normal (0, 0, -1 or +1), WSPos = the
    pixel centre's NDC position, nothing in shadow. Colour against the float64 evaluation of the pixel shader."""
    W, H = 64, 48
    o = _oracle(W, H)
    a, b, c, d = (-1, -1, .25), (1, -1, .25), (1, 1, .25), (-1, 1, .25)
    tris = [a, c, b, a, d, c] if facing else [a, b, c, a, c, d]          # cross(e1, e2) = -z when facing the eye at z < 0
    pos = np.asarray(tris, np.float32)
    eye = (0.3, 0.2, -3.0)
    light, amb = (0.6, 0.9, -1.0), (0.4, 0.6, 1.0)
    o.SetLight(light, (1.0, 0.7, 0.3), 2.0)
    o.SetAmbient(amb, 1.5)
    sh = None
    if use_sh:
        sh = np.random.RandomState(3).uniform(0.0, 0.5, (9, 3)).astype(np.float32); sh[0] += 1.0
        o.SetSH(sh)
    o.SetMesh(pos, np.arange(6, dtype=np.uint32))
    o.SetMeshWorld(1.0, (0, 0, 0))
    o.RenderMesh(np.eye(4, dtype=np.float32), eye, clear_rgba=(0.1, 0.2, 0.3, 0.0))
    rgba = o.ReadFrame().astype(np.float64)
    depth, _ = o.ReadDepth()
    assert np.all(depth == np.float32(0.25)) and np.all(rgba[..., 3] == 1.0)
    assert np.all(o.ReadVelocity().view(np.float16) == 0)                # first frame: previous = current (0 * -0.5 = -0)
    xs = (np.arange(W) + 0.5) / W * 2 - 1; ys = 1 - (np.arange(H) + 0.5) / H * 2
    ws = np.stack([np.broadcast_to(xs, (H, W)), np.broadcast_to(ys[:, None], (H, W)), np.full((H, W), 0.25)], -1)
    n = np.array([0.0, 0.0, -1.0 if facing else 1.0])
    want = _expected_shade(np.broadcast_to(n, ws.shape), ws, eye, light, (1.0, 0.7, 0.3, 2.0), amb + (1.5,), sh)
    assert np.abs(rgba[..., :3] - want).max() <= 2e-3 * max(1.0, want.max())      # fp16 storage
    if facing:
        assert want[..., 0].max() > want[..., 0].min() * 1.01            # the specular / Fresnel terms vary across the quad


def test_base_pass_clear_colour_and_uncovered_pixels():
    W, H = 64, 48
    o = _oracle(W, H)
    o.SetMesh(np.asarray([(-1, -1, .5), (0, 1, .5), (1, -1, .5)], np.float32), np.arange(3, dtype=np.uint32))
    o.RenderMesh(np.eye(4, dtype=np.float32), (0, 0, -3), clear_rgba=(0.25, 0.5, 0.75, 0.0))
    rgba, (depth, _) = o.ReadFrame(), o.ReadDepth()
    out = depth == 1.0
    assert 0 < out.sum() < W * H
    assert np.all(rgba[out] == np.asarray([0.25, 0.5, 0.75, 0.0], np.float16)) and np.all(rgba[~out][:, 3] == 1.0)
    d2 = _ndc_mesh(_oracle(W, H), [(-1, -1, .5), (0, 1, .5), (1, -1, .5)], W, H)
    assert np.array_equal(d2, depth)                                    # same coverage and depth as the depth-only pass


def test_base_pass_is_perspective_correct_and_writes_velocity():
    """One large triangle under the perspective camera, rendered from two eye points in turn. Depth, the view-dependent
    colour and the velocity (current minus previous clip position of the same surface point) are compared with a float64
    ray / plane intersection per pixel: a screen-space-linear interpolation would fail all three."""
    W, H = 160, 90
    o = _oracle(W, H)
    tri = np.asarray([(-30, -12, -20), (0, 25, 10), (34, -10, 40)], np.float32)       # strongly slanted in depth
    nrm = np.cross(tri[1] - tri[0], tri[2] - tri[1]).astype(np.float64); nrm /= np.linalg.norm(nrm)
    o.SetMesh(tri, np.arange(3, dtype=np.uint32))
    light, lrgbi, argbi = scene.LIGHT_PT, (1.0, 0.7, 0.3, 2.0), (0.4, 0.6, 1.0, 1.0)
    o.SetLight(light, lrgbi[:3], lrgbi[3]); o.SetAmbient(argbi[:3], argbi[3])
    cams = [scene.default_camera(W, H, eye=(4.0, 16.0, -80.0)), scene.default_camera(W, H, eye=(9.0, 14.0, -76.0))]
    for vp, eye in cams:
        o.RenderMesh(vp, eye)
    vp, eye = [np.asarray(x, np.float64) for x in cams[1]]
    vp0 = np.asarray(cams[0][0], np.float64)
    rgba, (depth, _), vel = o.ReadFrame().astype(np.float64), o.ReadDepth(), o.ReadVelocity().view(np.float16).astype(np.float64)
    cov = depth < 1.0
    assert cov.sum() > 1500
    inv = np.linalg.inv(vp)
    xs = (np.arange(W) + 0.5) / W * 2 - 1; ys = 1 - (np.arange(H) + 0.5) / H * 2
    ndc = np.stack([np.broadcast_to(xs, (H, W)), np.broadcast_to(ys[:, None], (H, W))], -1)
    def unproject(z):
        p = np.concatenate([ndc, np.full((H, W, 1), z), np.ones((H, W, 1))], -1) @ inv
        return p[..., :3] / p[..., 3:]
    p0, p1 = unproject(0.0), unproject(0.5)
    dirs = p1 - p0
    t = ((tri[0].astype(np.float64) - p0) @ nrm) / (dirs @ nrm)
    ws = p0 + dirs * t[..., None]
    clip = np.concatenate([ws, np.ones((H, W, 1))], -1) @ vp
    clip0 = np.concatenate([ws, np.ones((H, W, 1))], -1) @ vp0
    # interior pixels only (the snapped edges move the boundary by up to 1/256 pixel)
    inner = cov & np.roll(cov, 1, 0) & np.roll(cov, -1, 0) & np.roll(cov, 1, 1) & np.roll(cov, -1, 1)
    assert np.abs(depth[inner] - (clip[..., 2] / clip[..., 3])[inner]).max() < 2e-5
    want_v = (clip[..., :2] / clip[..., 3:] - clip0[..., :2] / clip0[..., 3:]) * np.array([0.5, -0.5])
    assert np.abs(want_v[inner]).max() > 0.01
    assert np.abs(vel[inner] - want_v[inner]).max() < 2e-3 * max(1.0, np.abs(want_v[inner]).max()) + 1e-4
    n = nrm if (nrm @ (eye - ws[inner][0])) != 0 else nrm
    want = _expected_shade(np.broadcast_to(n, ws.shape), ws, eye, light, lrgbi, argbi)
    # the triangle may be shadowed by nothing but itself: lit everywhere
    assert np.abs(rgba[..., :3][inner] - want[inner]).max() <= 3e-3 * max(1.0, want[inner].max())


def test_recomputed_normals_of_a_sphere_point_outward_or_inward_consistently():
    """ObjLoader::recomputeNormals on a UV sphere: the shaded disc is brightest toward the light and the ambient-only
    limb (NoL = 0) carries the hemisphere term of the normal there."""
    W, H = 160, 90
    o = _oracle(W, H)
    pos, idx = uv_sphere(radius=5.0, rings=48, sectors=96)
    o.SetMesh(pos, idx)
    o.SetMeshWorld(2.0, (0.0, 0.0, 0.0))
    o.SetLight(scene.LIGHT_PT, (1.0, 1.0, 1.0), 1.0); o.SetAmbient((1.0, 1.0, 1.0), 0.0)
    vp, eye = scene.default_camera(W, H)
    o.RenderMesh(vp, eye)
    rgba, (depth, _) = o.ReadFrame().astype(np.float64), o.ReadDepth()
    cov = depth < 1.0
    # analytic sphere normal per covered pixel
    inv = np.linalg.inv(np.asarray(vp, np.float64))
    xs = (np.arange(W) + 0.5) / W * 2 - 1; ys = 1 - (np.arange(H) + 0.5) / H * 2
    ndc = np.stack([np.broadcast_to(xs, (H, W)), np.broadcast_to(ys[:, None], (H, W)), depth.astype(np.float64), np.ones((H, W))], -1)
    p = ndc @ inv; ws = p[..., :3] / p[..., 3:]
    n = ws / np.maximum(np.linalg.norm(ws, axis=-1, keepdims=True), 1e-9)
    lit_out = _expected_shade(n, ws, eye, scene.LIGHT_PT, (1, 1, 1, 1.0), (1, 1, 1, 0.0))
    lit_in = _expected_shade(-n, ws, eye, scene.LIGHT_PT, (1, 1, 1, 1.0), (1, 1, 1, 0.0))
    inner = cov & np.roll(cov, 2, 0) & np.roll(cov, -2, 0) & np.roll(cov, 2, 1) & np.roll(cov, -2, 1)
    err_out = np.abs(rgba[..., :3][inner] - lit_out[inner]).mean(); err_in = np.abs(rgba[..., :3][inner] - lit_in[inner]).mean()
    assert min(err_out, err_in) < 0.02 and max(err_out, err_in) > 5 * min(err_out, err_in)
