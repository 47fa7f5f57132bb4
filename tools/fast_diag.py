#!/usr/bin/env python
"""Which pass's fast build moves the outputs by how much: renders small scenes with MV_FAST_MASK variations and compares
every output with the oracle (diagnostic; run on the GPU box)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from harness import blob_shadow, checker_background, configure, psnr
from oracle_binding import OracleCaster
from multivolumes_b200 import MultiRayCaster, scene

def cmp(name, got, want):
    g = got.astype(np.float32); w = want.astype(np.float32)
    err = np.abs(g - w) / np.maximum(1.0, np.abs(w))
    print(f"    {name:14s} max {err.max():.3e}  >2e-3: {int((err > 2e-3).sum()):7d} / {err.size}  >2e-4: {int((err > 2e-4).sum()):7d}  psnr {psnr(g, w):.1f}")

def run(mask, kw, cfg, frames=3, taa=True):
    os.environ["MV_FAST_MASK"] = str(mask)
    o = OracleCaster(filter_model=1, **kw); p = MultiRayCaster(**kw)
    for c in (o, p):
        configure(c, background=checker_background(kw["width"], kw["height"]), **cfg)
        for _ in range(frames):
            c.Render(); c.Postprocess(taa)
    so, sp = o.GetStats(), p.GetStats()
    print(f"  mask {mask:2d}: samples view {so['view_samples']}/{sp['view_samples']} light {so['light_samples']}/{sp['light_samples']} direct {so['direct_samples']}/{sp['direct_samples']} frags {so['oit_fragments']}/{sp['oit_fragments']}")
    lv = so["light_volume"]
    cmp("light map", p.ReadLightMap(lv), o.ReadLightMap(lv))
    att = o.ReadAttribs()
    for v in o.ReadCubeVolumes()[:2]:
        cmp(f"cube {v}", p.ReadCubeMap(int(v), int(att[v][0]))[0], o.ReadCubeMap(int(v), int(att[v][0]))[0])
    cmp("frame", p.ReadFrame(), o.ReadFrame())
    (to, bo), (tp, bp) = o.ReadPost(), p.ReadPost()
    cmp("taa", tp, to)
    d = np.abs(bo.astype(int) - bp.astype(int)); print(f"    rgba8 max {d.max()}  >1: {int((d > 1).sum())}  >0: {int((d > 0).sum())}")

scenes = [("small cube+sh", dict(grid_size=32, light_grid_size=16, num_volumes=9, num_volume_srcs=3, width=320, height=180), dict(sh=True, random_transforms=9, eye=(0, 40, -90))),
          ("direct+depth", dict(grid_size=64, light_grid_size=16, num_volumes=16, num_volume_srcs=4, width=320, height=180), dict(sh=True, eye=(10.0, 40.0, -160.0), shadow=blob_shadow())),
          ("mid 128", dict(grid_size=128, light_grid_size=48, num_volumes=4, width=960, height=540), dict(sh=True))]
for name, kw, cfg in scenes:
    print(name)
    for mask in (0, 1, 2, 4, 8, 16, 30, 31):
        run(mask, kw, cfg)
