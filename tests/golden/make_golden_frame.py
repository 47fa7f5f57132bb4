#!/usr/bin/env python
"""Writes tests/golden/oracle_frame_tiny.npz: every output of the path for one tiny animated scene, rendered by the CPU
oracle (texture-unit model of sm_100a, oracle/mvo_sampler.h). The reference itself cannot run here (HLSL / D3D12), so these
are the oracle's own outputs, frozen: the CPU suite checks that the oracle still reproduces them (a change of its arithmetic
must be deliberate and regenerate this file), the GPU suite checks the CUDA path against the committed bytes.

    python tests/golden/make_golden_frame.py        (needs oracle/_build/libmv_oracle.so: make -C oracle)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

KW = dict(grid_size=16, light_grid_size=8, num_volumes=6, num_volume_srcs=2, width=96, height=54)
FRAMES = 3


def render(caster):
    """The same calls on either backend; returns the dict of outputs the fixture holds."""
    from harness import blob_shadow, checker_background, configure
    from multivolumes_b200 import scene
    rs = np.random.RandomState(2)
    vel = (rs.uniform(-1, 1, (54, 96, 2)) * 0.003).astype(np.float16)
    vp0, _ = scene.default_camera(96, 54)
    depth = scene.sphere_depth(96, 54, vp0, center=(0, 0, 0), radius=7.0)
    configure(caster, sh=True, depth=depth, shadow=blob_shadow(), background=checker_background(96, 54), velocity=vel, random_transforms=4,
              eye=(5.0, 30.0, -75.0))
    svp = scene.shadow_view_proj()
    for f in range(FRAMES):
        vp, eye = scene.default_camera(96, 54, eye=(5.0 + 7 * f, 30.0 - 4 * f, -75.0 + 9 * f))
        caster.UpdateFrame(vp, svp, eye)
        caster.ResetColor(); caster.Render(use_work_graph=(f == 1)); caster.Postprocess(True)
    out = {"visible": caster.ReadVisible(), "cube_volumes": caster.ReadCubeVolumes()}
    att = caster.ReadAttribs()
    out["attribs_visible"] = att[out["visible"]]
    st = caster.GetStats()
    out["light_volume"] = np.uint32(st["light_volume"])
    out["counters"] = np.array([st[k] for k in ("view_rays", "view_samples", "view_light_fetches", "light_samples", "direct_rays", "direct_samples", "oit_fragments")], np.uint64)
    out["light_map"] = caster.ReadLightMap(st["light_volume"]).view(np.uint16)
    for v in out["cube_volumes"][:2]:
        c, d = caster.ReadCubeMap(int(v), int(att[v][0]))
        out[f"cube{v}_rgba"] = c.view(np.uint16)
        out[f"cube{v}_depth"] = d.view(np.uint32)
    out["frame"] = caster.ReadFrame().view(np.uint16)
    taa, rgba8 = caster.ReadPost()
    out["taa"] = taa.view(np.uint16)
    out["rgba8"] = rgba8
    return out


if __name__ == "__main__":
    from oracle_binding import OracleCaster
    data = render(OracleCaster(filter_model=1, **KW))
    assert len(data["visible"]) > 0 and len(data["cube_volumes"]) > 0 and data["counters"][1] > 0
    path = os.path.join(HERE, "oracle_frame_tiny.npz")
    np.savez_compressed(path, **data)
    print(path, os.path.getsize(path), "bytes;", {k: (v.shape if hasattr(v, "shape") else v) for k, v in data.items()})
