#!/usr/bin/env python
"""Small frames through every pass, for compute-sanitizer: compute-sanitizer --tool memcheck python tools/sanitize.py
Covers: instrumented (serial) and uninstrumented (frames pipelined on two streams) casters, RGBA16F and density-only
storage, the plain and the work-graph order, cube-map and direct-scheme volumes, the mesh depth producer, 40 volumes
(three clusters of the light march's pre-cull); round 2: empty-space bricks forced on (MV_OCC_BRICKS), the environment
pass, the screenshot, the mesh base pass, and volume-sharded storage on two virtual ranks of one device (proxies, owner-only marches, peer
stores into the other caster's exchange block, device-side barriers)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")
os.environ["MV_OCC_BRICKS"] = "8"
import numpy as np
from multivolumes_b200 import MultiRayCaster, scene
from harness import checker_background, configure, triangle_soup, uv_sphere
pos, idx = uv_sphere(radius=5.0, rings=16, sectors=32)
for variant in (dict(count_samples=True), dict(count_samples=False), dict(count_samples=False, density_only=True)):
    c = MultiRayCaster(grid_size=32, light_grid_size=16, num_volumes=40, num_volume_srcs=4, width=320, height=180, **variant)
    for eye in ((4.0, 16.0, -80.0), (10.0, 40.0, -160.0)):       # near: cube-map volumes; far: direct-scheme volumes
        configure(c, sh=True, background=checker_background(320, 180), eye=eye)
        c.SetMesh(pos, idx); c.SetMeshWorld(1.8, (0.0, -9.0, 0.0))
        vp, e = scene.default_camera(320, 180, eye=eye)
        svp = c.RenderMeshDepth(vp)
        c.SetEnvironment(scene.procedural_sky(16))
        for i in range(4):
            if i == 3:     # the shaded base pass as producer of colour / depth / shadow map / velocity (soup: clipped triangles)
                c.SetMesh(*triangle_soup(300, seed=2))
                svp = c.RenderMesh(vp, e, clear_rgba=(0.1, 0.1, 0.1, 0.0))
            c.UpdateFrame(vp, svp, e); c.RenderEnvironment(); c.Render(use_work_graph=(i == 2)); c.Postprocess(True)
        c.Screenshot("/tmp/mv_sanitize.png")
        print(variant, eye, {k: v for k, v in c.GetStats().items() if k in ("visible_count", "cubemap_count", "light_volume")})
    c.Sync()
    del c
# volume-sharded storage, two virtual ranks
kw = dict(grid_size=32, light_grid_size=12, num_volumes=9, num_volume_srcs=9, width=320, height=180)
ranks = [MultiRayCaster(shard_volumes=(r, 2, 8), count_samples=(r == 0), **kw) for r in range(2)]
for r in range(2):
    ranks[r].SetPeerBlock(1 - r, ranks[1 - r].ExchangeBlock()[0])
    configure(ranks[r], sh=True, background=checker_background(320, 180), random_transforms=5, eye=(8.0, 34.0, -120.0))
    ranks[r].SetRowBand(90 * r, 90 * (r + 1))
for f in range(3):
    vp, e = scene.default_camera(320, 180, eye=(8.0 + 3 * f, 34.0, -120.0 + 9 * f))
    for c in ranks:
        c.UpdateFrame(vp, None, e); c.ResetColor()
    for c in ranks:
        c.Render()
    for c in ranks:
        c.Postprocess(True)
    for c in ranks:
        c.Sync()
print("sharded volumes", ranks[0].GetStats()["visible_count"])
del ranks
print("done")
