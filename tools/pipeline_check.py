#!/usr/bin/env python
"""Pipelined (two-stream) frames against the serial path at full size: python tools/pipeline_check.py cfg2 [frames]
Both casters render the same orbit; the RGBA8 frame, the TAA image and the visible list are compared every few frames."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from multivolumes_b200 import MultiRayCaster, scene
wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg2"]
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 48
kw = dict(count_samples=False, density_only=bool(wl.get("density_only")), grid_size=wl["g"], light_grid_size=wl["l"], num_volumes=wl["n"], num_volume_srcs=wl.get("srcs"), width=wl["w"], height=wl["h"])
os.environ["MV_OVERLAP"] = "0"
serial = MultiRayCaster(**kw)
os.environ["MV_OVERLAP"] = "1"
piped = MultiRayCaster(**kw)
for c in (serial, piped):
    bench.build_scene(c, wl, scene, c.TransformSH(scene.procedural_sky(64)))
bad = 0
for f in range(frames):
    vp, eye = bench.camera(scene, wl, 7 * f)
    for c in (serial, piped):
        c.UpdateFrame(vp, None, eye); c.ResetColor(); c.Render(); c.Postprocess(wl["taa"])
    if f % 6 == 5 or f == frames - 1:
        (ta, ba), (tb, bb) = serial.ReadPost(), piped.ReadPost()
        same = np.array_equal(ba, bb) and np.array_equal(ta.view(np.uint16), tb.view(np.uint16)) and np.array_equal(serial.ReadVisible(), piped.ReadVisible())
        bad += 0 if same else 1
        print(f"frame {f}: {'identical' if same else 'DIFFERENT'} (mean level {ba[..., :3].mean():.2f})", flush=True)
print("PIPELINE_CHECK", "ok" if bad == 0 else f"{bad} mismatching read-backs")
sys.exit(1 if bad else 0)
