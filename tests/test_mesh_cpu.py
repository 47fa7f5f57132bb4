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
