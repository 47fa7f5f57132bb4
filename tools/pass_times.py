#!/usr/bin/env python
"""Per-pass device time of the frame loop (tuning aid): python tools/pass_times.py [workload] [frames]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from multivolumes_b200 import MultiRayCaster, scene

wl = dict(bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg2"])
wl.update(json.loads(os.environ.get("MV_WL", "{}")))
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 50
c = MultiRayCaster(count_samples=not os.environ.get("MV_NOSTATS"), time_passes=True, density_only=bool(wl.get("density_only")), grid_size=wl["g"], light_grid_size=wl["l"], num_volumes=wl["n"], num_volume_srcs=wl.get("srcs"), width=wl["w"], height=wl["h"])
bench.build_scene(c, wl, scene, c.TransformSH(scene.procedural_sky(64)))
acc = {}
for i in range(30 + frames):
    vp, eye = bench.camera(scene, wl, i)
    c.UpdateFrame(vp, None, eye); c.RenderEnvironment(); c.Render(); c.Postprocess(wl["taa"])
    if i >= 30:
        for k, v in c.GetTimings().items():
            acc[k] = acc.get(k, 0.0) + v / frames
st = c.GetStats()
print(json.dumps({"lib": os.path.basename(os.environ.get("MV_B200_LIB", "default")), **{k: round(v, 4) for k, v in acc.items()},
                  "view_samples": st["view_samples"], "view_skipped": st["view_skipped"], "light_samples": st["light_samples"], "direct_samples": st["direct_samples"], "direct_skipped": st["direct_skipped"]}))
