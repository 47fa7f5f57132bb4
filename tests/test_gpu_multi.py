"""Multi-GPU parity on real devices (needs >= 2 GPUs; skipped otherwise): the sharded frame — by
volume for the march, by z-slab for the light map, by row band for OIT + post-process — must equal the
single-GPU frame bit for bit, in both exchange modes (fused peer stores / NCCL collectives). The scene has a non-zero
velocity field and TAA on over four frames, so the history fetch (uv - velocity, bilinear) of one rank's rows lands on
rows another rank wrote: the fused mode must have exchanged the TAA output as well as the cube maps."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

WORKER = r'''
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, %(root)r); sys.path.insert(0, %(here)r)
from multivolumes_b200 import MultiRayCaster, scene
from multivolumes_b200.dist import CudaExchange, ShardedRenderer
from harness import checker_background, configure
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
mode, out = sys.argv[1], sys.argv[2]
# fewer GPUs than ranks: the ranks share devices (CUDA IPC and the device-side barriers work between processes on one GPU,
# which time-slices them); NCCL refuses that, so the few host-side exchanges of the fused mode go over gloo
ngpu = torch.cuda.device_count()
dev = rank %% ngpu
torch.cuda.set_device(dev)
if ngpu >= world:
    dist.init_process_group("nccl", device_id=torch.device("cuda", dev))
else:
    dist.init_process_group("gloo")
kw = dict(grid_size=64, light_grid_size=24, num_volumes=9, num_volume_srcs=3, width=640, height=360)
# uninstrumented casters take the pipelined paths (frames in flight on two streams); "fused-serial" keeps the counters on
c = MultiRayCaster(device=dev, count_samples=(os.environ.get("MV_TEST_VARIANT") == "fused-serial"), **kw)
stream = torch.cuda.Stream()
c.SetStream(stream.cuda_stream)
vel = (np.random.RandomState(7).uniform(-1, 1, (360, 640, 2)) * (0.02 if mode != "collective" else 0.0)).astype(np.float16)
configure(c, sh=True, background=checker_background(640, 360), velocity=vel)
with torch.cuda.stream(stream):
    x = CudaExchange(c, rank, world) if mode == "collective" else None
    r = ShardedRenderer(c, rank, world, mode=mode, exchange=x)
    for i in range(4):
        vp, eye = scene.default_camera(640, 360, eye=(4.0 + 6 * i, 16.0 + 8 * i, -80.0 - 30 * i))
        r.render(vp, None, eye, taa=True)
    c.Sync()
    dist.barrier()
    if rank == 0:
        taa, rgba8 = c.ReadPost()
        np.savez(out, rgba8=rgba8)
    else:
        # the last rank's arena must hold every cube map of the frame: whole maps from their owners (collective) or tile
        # ranges stored by whichever rank marched them (fused, tile-balanced split)
        att = c.ReadAttribs()
        cubes = {f"cube{v}": c.ReadCubeMap(int(v), int(att[v][0]))[0].view(np.uint16) for v in c.ReadCubeVolumes()}
        np.savez(out + f".r{rank}.npz", **cubes)
    dist.barrier()
dist.destroy_process_group()
'''


def _single(velocity_scale):
    sys.path.insert(0, HERE)
    from harness import checker_background, configure
    from multivolumes_b200 import MultiRayCaster, scene
    kw = dict(grid_size=64, light_grid_size=24, num_volumes=9, num_volume_srcs=3, width=640, height=360)
    c = MultiRayCaster(**kw)
    vel = (np.random.RandomState(7).uniform(-1, 1, (360, 640, 2)) * velocity_scale).astype(np.float16)
    configure(c, sh=True, background=checker_background(640, 360), velocity=vel)
    for i in range(4):
        vp, eye = scene.default_camera(640, 360, eye=(4.0 + 6 * i, 16.0 + 8 * i, -80.0 - 30 * i))
        c.UpdateFrame(vp, None, eye); c.ResetColor(); c.Render(); c.Postprocess(True)
    att = c.ReadAttribs()
    cubes = {f"cube{v}": c.ReadCubeMap(int(v), int(att[v][0]))[0].view(np.uint16) for v in c.ReadCubeVolumes()}
    return c.ReadPost()[1], cubes


@pytest.mark.parametrize("mode,world", [("fused", 2), ("fused-overlap", 2), ("fused-serial", 2), ("collective", 2), ("fused", 4), ("fused-serial", 4)])
def test_multi_gpu_frame_equals_single_gpu(mode, world, tmp_path):
    import torch
    if torch.cuda.device_count() < world and (mode == "collective" or world > 2):
        pytest.skip(f"needs {world} GPUs")        # (the fused modes at world 2 also run with both ranks on ONE GPU)
    script = tmp_path / "worker.py"
    script.write_text(WORKER % dict(root=ROOT, here=HERE))
    out = str(tmp_path / "out.npz")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29600 + os.getpid() % 300), str(script), mode, out]
    # "fused": frames pipelined across the ranks (light march of frame i + 1 beside frame i, two barrier channels);
    # "fused-overlap": not pipelined, light march of the frame's light volume beside the view march of the other volumes;
    # "fused-serial": every pass in order on one stream
    env = dict(os.environ, MV_TEST_VARIANT=mode)
    if mode == "fused-overlap":
        env["MV_SHARD_V_BLOCKS"] = "3"; env["MV_SHARD_PIPELINE"] = "0"
    if mode.startswith("fused-"):
        cmd[cmd.index(mode)] = "fused"
    subprocess.run(cmd, check=True, timeout=600, env=env)
    got = np.load(out)["rgba8"]
    # the collective mode gathers finished bands only (no history exchange): it is the baseline, run with a static field
    want, cubes = _single(0.02 if mode != "collective" else 0.0)
    if not np.array_equal(got, want):
        rows = np.nonzero((got != want).any(axis=(1, 2)))[0]
        raise AssertionError(f"{int((got != want).any(axis=2).sum())} pixels differ, rows {rows[:40].tolist()}, max diff {int(np.abs(got.astype(int) - want.astype(int)).max())}")
    peer = np.load(out + ".r1.npz")
    assert len(cubes) > 0 and sorted(peer.files) == sorted(cubes)
    for k, v in cubes.items():
        assert np.array_equal(peer[k], v), k
