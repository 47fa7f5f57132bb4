#!/usr/bin/env python
"""How far the frame with the HLSL text's decimal `min16float` literals sits from the frame with the binary16 literals the shipped
DXIL holds (the default of oracle and product):
renders BASELINE.json configs[0] (4 x 128^3, 1280x720) on the CPU oracle both ways and writes profiles/r02_min16_delta.json.
Also: the exact-fp32 trilinear sampler against the sm_100 texture-unit model (filter_model 0 vs 1)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import bench
from harness import psnr
from oracle_binding import OracleCaster, oracle_binding
from multivolumes_b200 import scene

wl = bench.WORKLOADS["cfg1"]
kw = dict(grid_size=wl["g"], light_grid_size=wl["l"], num_volumes=wl["n"], width=wl["w"], height=wl["h"])


def render(half, model):
    oracle_binding().set_min16_consts_as_half(half)
    o = OracleCaster(filter_model=model, **kw)
    bench.build_scene(o, wl, scene, None)
    for i in range(4):
        bench.step_frame(o, wl, scene, 20 * i, lambda vp, svp, eye: (o.UpdateFrame(vp, svp, eye), o.ResetColor(), o.Render(), o.Postprocess(False)))
    oracle_binding().set_min16_consts_as_half(1)
    return o.ReadFrame().astype(np.float32), o.ReadPost()[1], o.GetStats()


def delta(a, b):
    d = np.abs(a[0] - b[0]) / np.maximum(1.0, np.abs(b[0]))
    return {"frame_max_abs": float(d.max()), "frame_psnr_db": float(psnr(a[0], b[0])), "beyond_2e-3": int((d > 2e-3).sum()), "values": int(d.size),
            "rgba8_max_diff": int(np.abs(a[1].astype(int) - b[1].astype(int)).max()), "view_samples": [a[2]["view_samples"], b[2]["view_samples"]]}


base = render(1, 1)
out = {"workload": "cfg1 (4 x 128^3, 1280x720), 4 frames, TAA off",
       "hlsl_text_fp32_literals_vs_shipped_dxil_binary16_literals": delta(render(0, 1), base),
       "exact_fp32_trilinear_vs_sm100_texture_unit_model": delta(render(1, 0), base)}
with open(os.path.join(ROOT, "profiles", "r02_min16_delta.json"), "w") as f:
    json.dump(out, f, indent=1)
print(json.dumps(out, indent=1))
