#!/usr/bin/env python
"""ORACLE TOOLING: a randomised sweep of the oracle against the reference's compiled shaders (cull, light march, view march,
PSCube + PSResolveOIT) beyond the committed vectors.   python -m oracle.dxil.stress [first_seed] [count]     (run HERE)
Prints one line per scene and the totals; profiles/r02_dxil_stress.txt is such a run."""
import sys
import time

import numpy as np

from oracle.dxil import make_golden as G
import oracle.dxil.interp as I


def main(first=100, count=10):
    sys.path.insert(0, G.ROOT + "/tests")
    import harness
    from harness import dxil_scene, sh_coeffs
    from multivolumes_b200 import scene
    from oracle_binding import OracleCaster
    I.PROMOTE_HALF = True
    tot = dict(cull=[0, 0], light=[0, 0], view=[0, 0], view32=[0, 0], kcol=[0, 0], blend=[0, 0])
    t0 = time.time()
    for seed in range(first, first + count):
        rs = np.random.RandomState(seed)
        grid = int(rs.choice([16, 32])); n = int(rs.randint(2, 6))
        eye = tuple(float(v) for v in (rs.uniform(-12, 12), rs.uniform(3, 22), -rs.uniform(26, 70)))
        harness.DXIL_SCENES["stress"] = dict(seed=seed, grid=grid, light_grid=int(rs.choice([8, 12])), n=n, W=96, H=54,
                                             ray=int(rs.choice([32, 48, 96])), light=int(rs.choice([8, 16])), eye=eye)
        cfg = harness.DXIL_SCENES["stress"]
        model = int(rs.randint(0, 2))
        o, vp, e, depth, shadow = dxil_scene(OracleCaster, "stress", filter_model=model, light_maps=False)
        # cull
        c = G.cull_case(o, e, cfg["ray"])
        vis, cub, att = o.ReadVisible(), o.ReadCubeVolumes(), o.ReadAttribs()
        ok = (np.array_equal(np.sort(vis), np.sort(c["visible"])) and np.array_equal(np.sort(cub), np.sort(c["cube_volumes"]))
              and all(np.array_equal(att[v], c["volume_info"][v].astype(np.uint16)) for v in c["visible"]))
        tot["cull"][0] += 0 if ok else 1; tot["cull"][1] += 1
        line = [f"seed {seed}: G {grid} N {n} filter {model} visible {len(vis)} cube {len(cub)} cull {'ok' if ok else 'DIFFERS'}"]
        if len(vis):
            light = dict(eye=e, pos=tuple(scene.LIGHT_PT) + (1.0,), color=tuple(scene.LIGHT_COLOR) + (scene.LIGHT_INTENSITY,),
                         ambient=tuple(scene.AMBIENT_COLOR) + (scene.AMBIENT_INTENSITY,))
            for v in vis:
                v = int(v)
                want = G.march_l_case(o, v, shadow, scene.shadow_view_proj(), light, sh_coeffs(), cfg["light"])
                o.RayMarchL(v)
                got = o.ReadLightMap(v).view(np.float16)[..., :3].astype(np.float32)
                tot["light"][0] += int((got != want).sum()); tot["light"][1] += got.size
            o.DebugF32(True)
            dx = G.march_v_case(o, e, depth, cfg["ray"])
            dx32 = G.march_v_case(o, e, depth, cfg["ray"], f32=True)
            o.RayMarchV()
            cube32, _ = o.DebugF32(True)
            for v, (mip, rgba, dep) in dx.items():
                orgba, odep = o.ReadCubeMap(v, mip)
                m = dep >= 0
                tot["view"][0] += int((rgba.view(np.uint16)[m] != orgba.view(np.uint16)[m]).sum()) + int((dep[m] != odep[m]).sum()); tot["view"][1] += int(m.sum()) * 5
                s_ = o.G >> mip
                tot["view32"][0] += int((cube32[v, :, :s_, :s_].view(np.uint32)[m] != dx32[v][1].view(np.uint32)[m]).sum()); tot["view32"][1] += int(m.sum()) * 4
            r = G.oit_case(o, e, depth, stride=4)
            m = r["done"]
            if m.any():
                want = np.zeros_like(r["layers"])
                for py, px in np.argwhere(m):
                    for l in range(int(r["count"][py, px])):
                        if r["info"][py, px, l, 3]:
                            want[py, px, l] = r["data"][py, px, l, 5:9].astype(np.float16)
                tot["kcol"][0] += int((want.view(np.uint16)[m] != r["layers"].view(np.uint16)[m]).sum()); tot["kcol"][1] += int(r["count"][m].sum()) * 4
                tot["blend"][0] += int((r["blend"][m].view(np.uint32) != r["oracle_result"][m].view(np.uint32)).sum()); tot["blend"][1] += int(m.sum()) * 4
        line.append(" | running totals (differing / compared): " + ", ".join(f"{k} {a}/{b}" for k, (a, b) in tot.items()))
        print("".join(line), flush=True)
    print(f"TOTAL after {count} scenes in {time.time() - t0:.0f} s: " + ", ".join(f"{k} {a}/{b}" for k, (a, b) in tot.items()))


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 100, int(sys.argv[2]) if len(sys.argv) > 2 else 10)
