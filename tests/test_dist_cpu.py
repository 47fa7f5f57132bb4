"""N > 1 host logic on CPU: two gloo ranks drive the ORACLE through multivolumes_b200.dist (the same
ShardedRenderer the CUDA product uses on GPUs) and must reproduce the single-rank frame bit for bit —
sharding by volume / light slab / row band changes no per-texel arithmetic."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
KW = dict(grid_size=32, light_grid_size=16, num_volumes=6, num_volume_srcs=3, width=96, height=54)
FRAMES = 3


def _scene(c):
    from harness import checker_background, configure
    configure(c, sh=True, background=checker_background(96, 54), eye=(6.0, 14.0, -70.0))


def _camera(i):
    from multivolumes_b200 import scene
    return scene.default_camera(96, 54, eye=(6.0 + 3 * i, 14.0, -70.0))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, HERE); sys.path.insert(0, os.path.dirname(HERE))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle_binding import OracleCaster
    from multivolumes_b200.dist import HostExchange, ShardedRenderer
    c = OracleCaster(filter_model=1, threads=2, **KW)
    _scene(c)
    r = ShardedRenderer(c, rank, world, mode="collective", exchange=HostExchange(c, rank, world))
    for i in range(FRAMES):
        vp, eye = _camera(i)
        r.render(vp, None, eye, taa=True)
    if rank == 0:
        taa, rgba8 = c.ReadPost()
        np.savez(os.path.join(out_dir, "sharded.npz"), frame=c.ReadFrame().view(np.uint16), taa=taa.view(np.uint16), rgba8=rgba8)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_frame_equals_single_rank(world, oracle_lib, tmp_path):
    sys.path.insert(0, HERE)
    from oracle_binding import OracleCaster
    port = 29500 + (os.getpid() % 2000) + world
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "sharded.npz")
    c = OracleCaster(filter_model=1, threads=2, **KW)
    _scene(c)
    for i in range(FRAMES):
        vp, eye = _camera(i)
        c.UpdateFrame(vp, None, eye); c.ResetColor(); c.Render(); c.Postprocess(True)
    taa, rgba8 = c.ReadPost()
    assert np.array_equal(got["frame"], c.ReadFrame().view(np.uint16))
    assert np.array_equal(got["taa"], taa.view(np.uint16))
    assert np.array_equal(got["rgba8"], rgba8)


def test_partition_helpers():
    from multivolumes_b200.dist import light_slab, owner_of, row_band
    for H, w in ((1080, 8), (2160, 3), (54, 5), (7, 8)):
        bands = [row_band(H, r, w) for r in range(w)]
        assert bands[0][0] == 0 and bands[-1][1] == H and all(bands[i][1] == bands[i + 1][0] for i in range(w - 1))
    for L, w in ((96, 8), (96, 5), (16, 3), (5, 8)):
        slabs = [light_slab(L, r, w) for r in range(w)]
        assert slabs[0][0] == 0 and max(s[1] for s in slabs) == L
        assert sum(s[1] - s[0] for s in slabs) == L
    assert [owner_of(v, 4) for v in range(6)] == [0, 1, 2, 3, 0, 1]


def test_view_march_tile_ranges_tile_every_volume_exactly():
    """Fused-mode split of the view march (multivolumes_b200.dist.view_march_tile_range, mirrored by the cull kernel): for
    any tile count, volume position and world size the ranks' ranges are disjoint, cover [0, tiles), differ by at most one
    tile, and rotate with the volume's position in the march order."""
    sys.path.insert(0, os.path.dirname(HERE))
    from multivolumes_b200.dist import view_march_tile_range
    rs = np.random.RandomState(3)
    cases = [(0, 0, 2), (1, 5, 8), (7, 3, 8), (6 * 32 * 64, 11, 8), (3 * 16 * 32, 0, 3)] + \
            [(int(rs.randint(0, 60000)), int(rs.randint(0, 600)), int(rs.randint(1, 9))) for _ in range(300)]
    for tiles, k, world in cases:
        ranges = [view_march_tile_range(tiles, k, r, world) for r in range(world)]
        covered = np.zeros(tiles, np.int32)
        for b, e in ranges:
            assert 0 <= b <= e <= tiles
            covered[b:e] += 1
        assert (covered == 1).all(), (tiles, k, world)
        sizes = [e - b for b, e in ranges]
        assert max(sizes) - min(sizes) <= 1
        # rank r of volume k holds the part rank r + 1 holds of volume k + 1
        assert ranges[0] == view_march_tile_range(tiles, k + 1, 1 % world, world)
