#!/usr/bin/env python
"""Small frames through every pass, for compute-sanitizer: compute-sanitizer --tool memcheck python tools/sanitize.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from multivolumes_b200 import MultiRayCaster, scene
from harness import checker_background, configure, uv_sphere
c = MultiRayCaster(grid_size=32, light_grid_size=16, num_volumes=16, num_volume_srcs=4, width=320, height=180, count_samples=True)
pos, idx = uv_sphere(radius=5.0, rings=16, sectors=32)
for eye in ((4.0, 16.0, -80.0), (10.0, 40.0, -160.0)):       # near: cube-map volumes; far: direct-scheme volumes
    configure(c, sh=True, background=checker_background(320, 180), eye=eye)
    c.SetMesh(pos, idx); c.SetMeshWorld(1.8, (0.0, -9.0, 0.0))
    vp, e = scene.default_camera(320, 180, eye=eye)
    svp = c.RenderMeshDepth(vp)
    for i in range(3):
        c.UpdateFrame(vp, svp, e); c.ResetColor(); c.Render(); c.Postprocess(True)
    print(eye, c.GetStats())
c.Sync()
print("done")
