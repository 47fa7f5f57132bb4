#!/usr/bin/env python
"""Work counters of one steady-state frame per workload: python tools/stats.py cfg2 cfg3 cfg4"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from multivolumes_b200 import MultiRayCaster, scene
for name in sys.argv[1:] or ["cfg2"]:
    wl = bench.WORKLOADS[name]
    c = MultiRayCaster(count_samples=True, time_passes=True, grid_size=wl["g"], light_grid_size=wl["l"], num_volumes=wl["n"], width=wl["w"], height=wl["h"])
    bench.build_scene(c, wl, scene, c.TransformSH(scene.procedural_sky(64)))
    for i in range(wl["n"] + 4):
        vp, eye = bench.camera(scene, wl, i)
        c.UpdateFrame(vp, None, eye); c.ResetColor(); c.Render(); c.Postprocess(wl["taa"])
    print(name, json.dumps(c.GetStats()), json.dumps({k: round(v, 4) for k, v in c.GetTimings().items()}), flush=True)
    del c
