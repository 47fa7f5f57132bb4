#!/usr/bin/env python
"""Can the ALU-bound passes (resolve, TAA) of one frame run in the issue slots the texture-bound marches of another leave free?
Two independent casters of the same workload on ONE GPU, each on its own streams, frames interleaved without host syncs:
combined frames/s against one caster alone. (Tuning probe for a three-stage frame pipeline.)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from multivolumes_b200 import MultiRayCaster, scene

wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg4"]
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 150
kw = dict(grid_size=wl["g"], light_grid_size=wl["l"], num_volumes=wl["n"], width=wl["w"], height=wl["h"])


def make():
    c = MultiRayCaster(count_samples=False, **kw)
    bench.build_scene(c, wl, scene, c.TransformSH(scene.procedural_sky(64)))
    return c


def frame(c, i):
    bench.step_frame(c, wl, scene, i, lambda vp, svp, eye: (c.UpdateFrame(vp, svp, eye), c.RenderEnvironment(), c.Render(), c.Postprocess(wl["taa"])))


for i in range(frames + 40):
    bench.camera(scene, wl, i)
casters = [make(), make()]
for n in (1, 2):
    cs = casters[:n]
    for i in range(40):
        for c in cs:
            frame(c, i)
    for c in cs:
        c.Sync()
    t0 = time.perf_counter()
    for i in range(frames):
        for c in cs:
            frame(c, 40 + i)
    for c in cs:
        c.Sync()
    dt = time.perf_counter() - t0
    print(f"{n} caster(s): {n * frames / dt:.1f} frames/s combined ({1000 * dt / frames:.3f} ms per round)", flush=True)
