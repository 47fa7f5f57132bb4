#!/usr/bin/env python
"""Sharded frame vs one-GPU frame on a bench workload (diagnostic): under torch.distributed.run, every rank renders K frames
of the workload sharded; rank 0 then renders the same frames with an unsharded caster and compares the RGBA8 frames."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch, torch.distributed as dist
import bench
from multivolumes_b200 import MultiRayCaster, scene
from multivolumes_b200.dist import ShardedRenderer

wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg4"]
K = int(sys.argv[2]) if len(sys.argv) > 2 else 6
instr = len(sys.argv) > 3 and sys.argv[3] == "serial"
STEP = int(os.environ.get("MV_STEP", "7"))
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
kw = dict(grid_size=wl["g"], light_grid_size=wl["l"], num_volumes=wl["n"], num_volume_srcs=wl.get("srcs"), width=wl["w"], height=wl["h"])


def make(device):
    c = MultiRayCaster(device=device, count_samples=instr, **kw)
    bench.build_scene(c, wl, scene, c.TransformSH(scene.procedural_sky(64)))
    return c


c = make(rank)
stream = torch.cuda.Stream()
c.SetStream(stream.cuda_stream)
with torch.cuda.stream(stream):
    r = ShardedRenderer(c, rank, world, mode="fused")
    for i in range(K):
        if not instr and i == K // 3:
            c.SetInstrumentation(True, False)        # as bench.py does between its passes: pipelined -> serial -> pipelined
        if not instr and i == K // 3 + 4:
            c.SetInstrumentation(False, False)
        bench.step_frame(c, wl, scene, STEP * i, lambda vp, svp, eye: r.render(vp, svp, eye, taa=wl["taa"]))
    c.Sync(); dist.barrier()
    if rank == 0:
        taa, got = c.ReadPost()
        frame = c.ReadFrame()
        lv = c.GetStats()["light_volume"]
        lm = c.ReadLightMap(lv)
    dist.barrier()
if rank == 0:
    s = make(0)
    for i in range(K):
        bench.step_frame(s, wl, scene, STEP * i, lambda vp, svp, eye: (s.UpdateFrame(vp, svp, eye), s.RenderEnvironment(), s.Render(), s.Postprocess(wl["taa"])))
    taa1, want = s.ReadPost()
    d = (got != want).any(axis=2)
    print(f"world {world} {'serial' if instr else 'pipelined'}: rgba8 pixels differing {int(d.sum())}, rows {np.nonzero(d.any(axis=1))[0][:30].tolist()}, cols {np.nonzero(d.any(axis=0))[0][:20].tolist()}")
    dt = (taa.view(np.uint16) != taa1.view(np.uint16)).any(axis=2)
    print(f"  taa texels differing {int(dt.sum())}, rows {np.nonzero(dt.any(axis=1))[0][:30].tolist()}")
    print(f"  light volume {lv} vs {s.GetStats()['light_volume']}: light map texels differing {int((lm.view(np.uint16) != s.ReadLightMap(lv).view(np.uint16)).any(axis=3).sum())}")
    att = s.ReadAttribs()
    for v in s.ReadCubeVolumes():
        a, b = c.ReadCubeMap(int(v), int(att[v][0])), s.ReadCubeMap(int(v), int(att[v][0]))
        n = int((a[0].view(np.uint16) != b[0].view(np.uint16)).any(axis=3).sum())
        if n:
            print(f"  cube map {v}: {n} texels differ")
dist.barrier()
dist.destroy_process_group()
