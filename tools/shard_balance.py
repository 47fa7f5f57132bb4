#!/usr/bin/env python
"""Per-rank share of the view march under the sharded frame (run under torchrun): rays, samples and the march's own
device time (CUDA events around the launch, peer barrier excluded) on every rank.
python -m torch.distributed.run --nproc-per-node N tools/shard_balance.py cfg4"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import bench
from multivolumes_b200 import MultiRayCaster, scene
from multivolumes_b200.dist import ShardedRenderer
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg2"]
c = MultiRayCaster(device=rank, count_samples=True, time_passes=False, density_only=bool(wl.get("density_only")), grid_size=wl["g"], light_grid_size=wl["l"],
                   num_volumes=wl["n"], num_volume_srcs=wl.get("srcs"), width=wl["w"], height=wl["h"])
stream = torch.cuda.Stream(); c.SetStream(stream.cuda_stream)
bench.build_scene(c, wl, scene, c.TransformSH(scene.procedural_sky(64)))
with torch.cuda.stream(stream):
    r = ShardedRenderer(c, rank, world, mode="fused")
    acc = np.zeros(4)
    frames = 20
    for i in range(10 + frames):
        vp, eye = bench.camera(scene, wl, i)
        r.render(vp, None, eye, taa=wl["taa"])
        if i >= 10:
            st = c.GetStats()
            acc += [st["view_rays"], st["view_samples"], st["direct_samples"], st["oit_fragments"]]
    c.Sync()
    t = torch.tensor(acc / frames, device="cuda", dtype=torch.float64)
    g = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(g, t)
    if rank == 0:
        for q, v in enumerate(g):
            print(f"{sys.argv[1]} rank {q}: view rays {v[0].item():.0f} samples {v[1].item():.0f} direct samples {v[2].item():.0f} fragments {v[3].item():.0f}")
dist.barrier(); dist.destroy_process_group()
