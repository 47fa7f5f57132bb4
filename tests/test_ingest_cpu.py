"""CPU checks of the DDS ingest: the header parser of the product library (host only) on files
written by tests/dds_util.py, and the oracle's CSR32FToRGBA16F resampling on closed-form inputs."""
import numpy as np
import pytest

from dds_util import write_dds
from oracle_binding import OracleCaster


@pytest.fixture(scope="module", autouse=True)
def _built(oracle_lib, product_lib):
    return oracle_lib


@pytest.mark.parametrize("kind,dx10,fmt,bpt,off", [("r32f", True, 1, 4, 148), ("r16f", True, 2, 2, 148), ("r16un", True, 3, 2, 148),
                                                   ("r8un", True, 4, 1, 148), ("r32f", False, 1, 4, 128), ("r16f", False, 2, 2, 128),
                                                   ("r8un", False, 4, 1, 128)])
def test_dds_header_parse(tmp_path, kind, dx10, fmt, bpt, off):
    from multivolumes_b200 import parse_dds
    vol = np.random.RandomState(1).uniform(0, 1, (5, 6, 7)).astype(np.float32)
    p = tmp_path / "v.dds"
    write_dds(str(p), vol, kind, dx10, mips=2)
    info = parse_dds(str(p))
    assert info == dict(width=7, height=6, depth=5, format=fmt, bytes_per_texel=bpt, data_offset=off)


def test_dds_parse_rejects(tmp_path):
    from multivolumes_b200 import parse_dds
    vol = np.zeros((4, 4, 4), np.float32)
    good = tmp_path / "g.dds"
    write_dds(str(good), vol)
    data = good.read_bytes()
    for name, blob in (("magic", b"XXXX" + data[4:]), ("short", data[:200]), ("tiny", data[:60]),
                       ("2d", data[:128] + data[128:132] + b"\x03\0\0\0" + data[136:]),          # resourceDimension = TEXTURE2D
                       ("rgba", data[:128] + b"\x02\0\0\0" + data[132:])):                       # DXGI R32G32B32A32_FLOAT
        p = tmp_path / (name + ".dds")
        p.write_bytes(blob)
        with pytest.raises(RuntimeError):
            parse_dds(str(p))
    with pytest.raises(RuntimeError):
        parse_dds(str(tmp_path / "missing.dds"))


def _o(g=16):
    return OracleCaster(filter_model=1, grid_size=g, light_grid_size=8, num_volumes=1, width=32, height=32)


def test_oracle_resample_same_size_is_identity():
    a, b = _o(), _o()
    vol = np.random.RandomState(2).uniform(0, 1, (16, 16, 16)).astype(np.float32)
    a.LoadVolumeData(0, vol.reshape(-1))            # the direct path (texel == voxel)
    assert a.b.volume_upload_r32f_sized(b.h, 0, vol.ctypes.data, 16, 16, 16) == 0
    assert np.array_equal(a.ReadVolume(0), b.ReadVolume(0))


def test_oracle_resample_constant_and_ramp():
    o = _o(16)
    o.LoadVolumeData(0, np.full((24, 24, 24), 0.5, np.float32))
    v = o.ReadVolume(0).view(np.float16)
    assert np.all(v[..., 3] == np.float16(0.125)) and np.all(v[..., :3] == np.float16(1.0))
    # a ramp along x, twice the grid's resolution: a voxel centre falls on the boundary of two source texels
    ramp = np.broadcast_to((np.arange(32, dtype=np.float32) / 32.0)[None, None, :], (32, 32, 32)).copy()
    o.LoadVolumeData(0, ramp)
    a = o.ReadVolume(0).view(np.float16)[..., 3].astype(np.float32) * 4.0
    want = (np.arange(16) * 2 + 0.5) / 32.0
    assert np.abs(a[3, 5, :] - want).max() < 1e-3
