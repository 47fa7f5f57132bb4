"""The environment pass (LightProbe::RenderEnvironment + PSEnvironment.hlsl) of the oracle pinned by closed forms, and the PNG
writer of the product library (host-only entry point, no device needed) decoded with zlib."""
import struct
import zlib

import numpy as np
import pytest

from multivolumes_b200 import scene


@pytest.fixture(scope="module", autouse=True)
def _built(oracle_lib):
    return oracle_lib


def _oracle(**kw):
    from oracle_binding import OracleCaster
    return OracleCaster(filter_model=0, **dict(dict(grid_size=16, light_grid_size=8, num_volumes=1, width=96, height=54), **kw))


def test_environment_constant_radiance_and_depth_test():
    """A constant cube map filters to the constant whatever the direction; alpha is 0 (PSEnvironment.hlsl:68); pixels in
    front of which the mesh pass wrote a depth < 1 keep the colour RT (DEPTH_READ_LESS_EQUAL against the quad's z = 1)."""
    o = _oracle()
    cube = np.empty((6, 8, 8, 3), np.float32); cube[...] = (0.25, 0.5, 2.0)
    depth = np.ones((54, 96), np.float32); depth[10:20, 30:50] = 0.7
    bg = np.zeros((54, 96, 4), np.float16); bg[..., 0] = 0.125; bg[..., 3] = 1.0
    o.SetEnvironment(cube)
    o.SetRenderTargets(depth=depth, color=bg)
    vp, eye = scene.default_camera(96, 54)
    o.UpdateFrame(vp, None, eye)
    o.RenderEnvironment()
    f = o.ReadFrame().astype(np.float32)
    sky = depth >= 1.0
    assert np.all(f[sky] == np.array([0.25, 0.5, 2.0, 0.0], np.float32))
    assert np.all(f[~sky] == np.array([0.125, 0.0, 0.0, 1.0], np.float32))
    o.SetEnvironment(None)
    o.RenderEnvironment()
    assert np.array_equal(o.ReadFrame().view(np.uint16), bg.view(np.uint16))


@pytest.mark.parametrize("focus,face", [((0, 0, 100), 4), ((0, 0, -100), 5), ((100, 0, 0), 0), ((-100, 0, 0), 1), ((0, 100, 1), 2), ((0, -100, 1), 3)])
def test_environment_face_selection(focus, face):
    """Every face of the cube map in its own colour: the pixel at the centre of the screen, whose ray is the view direction,
    must show the colour of the face that direction points into (D3D order +X, -X, +Y, -Y, +Z, -Z)."""
    o = _oracle(width=65, height=65)
    cube = np.zeros((6, 4, 4, 3), np.float32)
    for f in range(6):
        cube[f] = (f + 1, 10 * (f + 1), 0.5)
    o.SetEnvironment(cube)
    o.SetRenderTargets()
    vp, eye = scene.default_camera(65, 65, eye=(0.0, 0.0, 0.0), focus=focus)
    o.UpdateFrame(vp, None, eye)
    o.RenderEnvironment()
    c = o.ReadFrame().astype(np.float32)[32, 32]
    assert tuple(c[:3]) == (face + 1, 10 * (face + 1), 0.5), (face, c)


def test_environment_bilinear_between_texels():
    """Looking along +Z at a 2x2 face: the centre pixel's direction hits (u, v) = (0.5, 0.5), the common corner of the four
    texels, so the result is their mean (fp32 weights 0.5 / 0.5)."""
    o = _oracle(width=65, height=65)
    cube = np.zeros((6, 2, 2, 3), np.float32)
    cube[4, :, :, 0] = [[1.0, 3.0], [5.0, 7.0]]
    o.SetEnvironment(cube)
    o.SetRenderTargets()
    vp, eye = scene.default_camera(65, 65, eye=(0.0, 0.0, 0.0), focus=(0, 0, 100))
    o.UpdateFrame(vp, None, eye)
    o.RenderEnvironment()
    assert abs(float(o.ReadFrame()[32, 32, 0]) - 4.0) < 2e-2


@pytest.mark.parametrize("shape", [(37, 53), (200, 120)])
def test_png_writer_round_trip(tmp_path, product_lib, shape):
    """mv_write_png (MultiVolumes::SaveImage's job): signature, chunk CRCs, IHDR, and the zlib stream (stored blocks, the
    second shape needs more than one) inflate back to the image with filter type 0 on every row."""
    from multivolumes_b200 import write_png
    h, w = shape
    img = (np.random.RandomState(h).rand(h, w, 4) * 255).astype(np.uint8)
    path = str(tmp_path / "shot.png")
    write_png(path, img)
    d = open(path, "rb").read()
    assert d[:8] == b"\x89PNG\r\n\x1a\n"
    pos, idat, seen = 8, b"", []
    while pos < len(d):
        n, = struct.unpack(">I", d[pos:pos + 4]); typ = d[pos + 4:pos + 8]; data = d[pos + 8:pos + 8 + n]
        crc, = struct.unpack(">I", d[pos + 8 + n:pos + 12 + n])
        assert zlib.crc32(typ + data) & 0xffffffff == crc, typ
        seen.append(typ)
        if typ == b"IHDR":
            assert struct.unpack(">IIBBBBB", data) == (w, h, 8, 6, 0, 0, 0)
        if typ == b"IDAT":
            idat += data
        pos += 12 + n
    assert seen[0] == b"IHDR" and seen[-1] == b"IEND"
    rows = np.frombuffer(zlib.decompress(idat), np.uint8).reshape(h, 1 + w * 4)
    assert (rows[:, 0] == 0).all() and np.array_equal(rows[:, 1:].reshape(h, w, 4), img)
