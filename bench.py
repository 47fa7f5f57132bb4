#!/usr/bin/env python
"""bench.py — frames/s and ray-march samples/s of the cube-map-space volume rendering path.

  python bench.py --gpus 1 --steps K --warmup W                 (one process; N = 1)
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference ...                          (the CPU oracle on the host cores)

A step is one frame: UpdateFrame (camera on the reference's orbit, MultiVolumes.cpp:328-337) ->
colour-RT reset -> Render (cull, light march of one volume, view march, OIT resolve) -> Postprocess
(TAA + tone map). Workload at every N: BASELINE.json configs[3], the configuration the north_star's target is quoted
on — 64 volumes of 256^3 at 3840x2160, animated transforms + TAA, SH environment lighting (9.5 GB, fits one GPU) —
unless --workload says otherwise. Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: N, srcs, G, L, W, H, sh, taa
    "cfg1": dict(n=4, g=128, l=96, w=1280, h=720, sh=False, taa=False, note="BASELINE.json configs[0]"),
    "cfg2": dict(n=16, g=128, l=96, w=1920, h=1080, sh=True, taa=True, note="BASELINE.json configs[1]"),
    "cfg3": dict(n=64, g=256, l=96, w=1920, h=1080, sh=True, taa=True, mesh=True,
                 note="BASELINE.json configs[2]: scene depth + shadow map rasterised every frame from the occluder mesh (mv_mesh_*)"),
    "cfg4": dict(n=64, g=256, l=96, w=3840, h=2160, sh=True, taa=True, animate=True,
                 note="BASELINE.json configs[3]: per-volume rotation about y at seeded rates (SetVolumeWorldMatrix every frame)"),
    "cfg5i": dict(n=512, srcs=8, g=512, l=96, w=3840, h=2160, sh=True, taa=True,
                  note="BASELINE.json configs[4] with 8 source volumes instanced 64x (VolTexId = i % srcs): 8.6 GB instead of 512 GB of RGBA16F"),
    "cfg5": dict(n=512, g=512, l=96, w=3840, h=2160, sh=True, taa=True, density_only=True,
                 note="BASELINE.json configs[4]: 512 distinct source volumes in the density-only R16F storage SURVEY.md 8(d) names for it "
                      "(137 GB resident on every GPU; colour (1, 1, 1) as for the reference's file assets)"),
    "cfg5s": dict(n=512, g=512, l=96, w=3840, h=2160, sh=True, taa=True, shard_volumes=True, proxy=128,
                  note="BASELINE.json configs[4] as specified: 512 distinct RGBA16F sources of 512^3 (550 GB) sharded by volume — rank r holds the "
                       "sources s % world == r (69 GB at 8 ranks) and a 128^3 R16F density proxy of every other one (mv_create_sharded); needs >= 4 GPUs"),
    "cfg5s-mini": dict(n=64, g=256, l=96, w=1920, h=1080, sh=True, taa=True, shard_volumes=True, proxy=64,
                       note="volume-sharded storage at cfg3's size (for 2-GPU runs)"),
    "cfg2d": dict(n=16, g=128, l=96, w=1920, h=1080, sh=True, taa=True, density_only=True, note="cfg2's densities in the density-only R16F storage"),
    "cfg4d": dict(n=64, g=256, l=96, w=3840, h=2160, sh=True, taa=True, density_only=True, note="cfg4's densities in the density-only R16F storage"),
    "tiny": dict(n=4, g=32, l=16, w=320, h=180, sh=True, taa=True, note="CI-sized"),
}
def ncu_summary(workload):
    """Numbers that only a profiler can give (DRAM traffic of the dominant kernel, texture-pipe utilisation), read from the
    tracked summary of the committed ncu capture of this workload — never constants in this file. None when there is none."""
    p = os.path.join(ROOT, "profiles", "ncu_summary.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f).get(workload)
    return None


def tex_peak():
    """Trilinear RGBA16F fetch rate measured on this pool's B200 by tools/tex_probe.cu (L1-resident, coherent quads)."""
    with open(os.path.join(ROOT, "profiles", "r01_tex_probe.json")) as f:
        d = json.load(f)
    return float(d["tex_rate"]["rgba16f_32_l1"]) / 1e9, float(d["l2_read_gbs"]["64MB"])


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def build_scene(c, wl, scene, sky_coeffs):
    for i in range(c.srcs):
        c.InitVolumeData(i, 1, (0x9E3779B9 * (i + 1)) & 0xffffffff)
        if wl.get("density_only") and not getattr(c, "density_only", False):
            # the oracle has one storage format: give it the same volume, (1, 1, 1, a) as RGBA16F
            v = c.ReadVolume(i).copy()
            v[..., :3] = 1.0
            c.LoadVolumeData(i, v)
    c.SetLight(scene.LIGHT_PT, scene.LIGHT_COLOR, scene.LIGHT_INTENSITY)
    c.SetAmbient(scene.AMBIENT_COLOR, scene.AMBIENT_INTENSITY)
    c.SetVolumesWorld(20.0, (0, 0, 0))
    c.SetSH(sky_coeffs if wl["sh"] else None)
    if wl["sh"]:
        c.SetEnvironment(scene.procedural_sky(64))     # the radiance map the SH coefficients come from (LightProbe): drawn behind the volumes
    c.SetRenderTargets(depth=None)
    if wl.get("mesh"):
        # the occluder of Bin/all64.bat: bunny.obj where the reference tree is at hand (this container), its stand-in of the
        # same extent and triangle count elsewhere (the GPU box has no /root/reference). Depth and shadow map are produced
        # from it every frame by the caster's own rasteriser (mv_mesh_render_depth).
        pos, idx = occluder(scene)
        c.SetMesh(pos, idx)
        c.SetMeshWorld(*scene.MESH_WORLD)


_OCCLUDER = None


def occluder(scene):
    global _OCCLUDER
    if _OCCLUDER is None:
        path = os.environ.get("MV_OCCLUDER_OBJ", "/root/reference/Bin/Assets/bunny.obj")
        if os.path.exists(path):
            from multivolumes_b200 import parse_obj
            _OCCLUDER = parse_obj(path) + (os.path.basename(path),)
        else:
            _OCCLUDER = scene.occluder_mesh() + ("procedural stand-in (69 696 triangles)",)
    return _OCCLUDER[0], _OCCLUDER[1]


_WORLDS = {}


def animate(c, wl, frame):
    """cfg4 'animated transforms': every volume of the reference's grid turns about its own y axis at a seeded rate
    (SetVolumeWorld only places axis-aligned boxes, so the matrices go in through SetVolumeWorldMatrix). The matrices of a
    frame are a pure function of its index: they are kept, so that a frame's host inputs are computed once."""
    if not wl.get("animate"):
        return
    key = (wl["n"], frame)
    if key in _WORLDS:
        c.SetVolumeWorldMatrices(_WORLDS[key])
        return
    t = frame / 60.0
    n = wl["n"]
    row = int(np.ceil(np.sqrt(n))); col = n // row
    size = 20.0
    i = np.arange(row * col)
    rate = ((i * 2654435761) % 1000) / 1000.0 * 1.6 - 0.8            # rad/s, seeded by the index
    ang = rate * t
    cs, sn = np.cos(ang) * (size * 0.5), np.sin(ang) * (size * 0.5)
    px = -((row / 2.0 - 0.5) * size * 1.5) + (i % row) * size * 1.5
    pz = -((col / 2.0 - 0.5) * size * 1.5) + (i // row) * size * 1.5
    m = np.zeros((row * col, 4, 3), np.float32)
    m[:, 0, 0] = cs; m[:, 0, 2] = -sn; m[:, 1, 1] = size * 0.5; m[:, 2, 0] = sn; m[:, 2, 2] = cs
    m[:, 3, 0] = px; m[:, 3, 2] = pz
    _WORLDS[key] = m
    c.SetVolumeWorldMatrices(m)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def window(self, t0, t1):
        self.t0, self.t1 = t0, t1

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if self.t0 - 0.05 <= t <= self.t1 + 0.15] or [r for _, r in self.rows[-3:]]
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except Exception:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


_CAMERAS = {}


def camera(scene, wl, frame):
    """View-projection matrix and eye of frame `frame` of the orbit (host inputs of UpdateFrame; computed once per frame index)."""
    key = (wl["w"], wl["h"], frame)
    if key not in _CAMERAS:
        _CAMERAS[key] = scene.orbit_camera(wl["w"], wl["h"], frame)
    return _CAMERAS[key]


def host_threads():
    """Host cores this process may use. torch.distributed.run exports OMP_NUM_THREADS=1 to its workers; the CPU arms size
    their OpenMP team explicitly from the affinity mask instead, so that N > 1 launches time the same thing as N = 1."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def make_oracle(wl, threads):
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_binding import OracleCaster
    return OracleCaster(filter_model=1, threads=threads, grid_size=wl["g"], light_grid_size=wl["l"], num_volumes=wl["n"],
                        num_volume_srcs=wl.get("srcs"), width=wl["w"], height=wl["h"])


def step_frame(c, wl, scene, i, render):
    """One frame's inputs on either backend: animated transforms, occluder depth (cfg3), camera matrices; `render` draws."""
    vp, eye = camera(scene, wl, i)
    animate(c, wl, i)
    svp = c.RenderMeshDepth(vp) if wl.get("mesh") else None
    render(vp, svp, eye)


# --------------------------------------------------------------------------------------------
def run_reference(args, wl, rank, world):
    """--impl reference: the reference publishes no CPU implementation (HLSL on D3D12 only), so the arm
    times the scalar C++/OpenMP transliteration (the oracle) on the host cores, all threads."""
    if rank != 0:
        return None
    from multivolumes_b200 import scene
    cores = host_threads()
    o = make_oracle(wl, cores)
    sky = o.TransformSH(scene.procedural_sky(64))
    build_scene(o, wl, scene, sky)

    def frame(i, shard=None):
        if shard:
            o.SetShard(*shard)
            o.SetRowBand(*shard_band(wl["h"], *shard))

        def render(vp, svp, eye):
            o.UpdateFrame(vp, svp, eye); o.RenderEnvironment(); o.Render(); o.Postprocess(wl["taa"])
        step_frame(o, wl, scene, i, render)
        st = o.GetStats()
        return st["view_samples"] + st["direct_samples"] + st["light_samples"]

    def shard_band(H, r, w):
        return (H * r) // w, (H * (r + 1)) // w

    # calibrate with one full frame, then size the per-step sample so the whole run stays bounded
    t = time.perf_counter(); frame(0); t_full = time.perf_counter() - t
    budget = float(args.ref_budget)
    S = int(min(wl["n"], max(1, np.ceil((args.steps + args.warmup) * t_full / budget))))
    for i in range(args.warmup):
        frame(1 + i, (i % S, S))
    t0 = time.perf_counter(); samples = 0
    for i in range(args.steps):
        samples += frame(1 + args.warmup + i, (i % S, S))
    dt = time.perf_counter() - t0
    fps = args.steps / (dt * S)            # S sample-steps make one frame's worth of work
    line = {"impl": "reference", "metric": "frames_per_s", "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000.0 * dt / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, wl, world), "samples_per_s": samples / dt,
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": o.GetStats()["threads"], "kind": "port",
                             "sample": f"each step = 1/{S} of a frame (volumes v % {S} == step % {S}, light-map slab, row band {S}-th); "
                                       f"frames/s = steps / (time x {S}); calibration full frame {t_full:.2f} s"},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    return line


def launches_per_frame(wl, world, exchange, work_graph=False):
    """Kernels of libmv_b200.so per frame: k_cull; the light march (k_light_classify, k_ray_march_l and, with a light probe,
    k_light_scan, k_light_emit, k_light_ao, k_light_finalize); k_ray_march_v; k_ray_cast_direct; k_resolve_oit; k_postprocess.
    One GPU, pipelined frames: + k_light_commit. Sharded, fused exchange: + k_light_commit and three peer barriers
    (k_peer_signal + k_peer_wait each); with the light / view overlap (8 ranks) the view march is two launches.
    With a light probe: + k_environment. cfg3: + the depth / shadow producer (k_mesh_setup + k_mesh_raster per pass, clears,
    D16 conversion). Volume-sharded storage: two barriers and a commit instead of the slab exchange."""
    n = 1 + (6 if wl["sh"] else 2) + 1 + 1 + 1 + 1   # --work-graph: the same count (k_pick_light_volume instead of k_cull)
    n += 1 if wl["sh"] else 0
    if world == 1 and not work_graph and os.environ.get("MV_OVERLAP", "1") != "0":
        n += 1                                        # pipelined frames: k_light_commit (light map through the staging buffer)
    if world > 1:      # fused: k_light_commit + light-channel signal / wait + two barriers on the main channel (signal + wait each)
        n += 1 + ((4 if wl.get("shard_volumes") else 6) if exchange == "fused" else 0)
    if wl.get("mesh"):
        n += 7
    return n


def workload_config(args, wl, world):
    fmt = "R16F density-only" if wl.get("density_only") else "RGBA16F"
    return {"workload": f"{args.workload}: {wl['n']} volumes x {wl['g']}^3 {fmt} (procedural density x seeded value noise), "
                        f"{wl['w']}x{wl['h']}, light map {wl['l']}^3, SH {'on' if wl['sh'] else 'off'}, TAA {'on' if wl['taa'] else 'off'}, "
                        f"orbit camera; {wl['note']}",
            "l2_policy": f"inputs larger than L2 ({((wl.get('srcs') or wl['n']) * wl['g'] ** 3 * (2 if wl.get('density_only') else 8) + wl['n'] * wl['l'] ** 3 * 8) / 1e6:.0f} MB "
                         f"of volume and light-map textures vs 126 MB)",
            "parallelism": "1 GPU" if world == 1 else (f"{world} GPUs, volume-sharded storage: light map / cube map / screen-space march of a volume by the rank that holds it, "
                                                       f"interleaved row stripes for the resolve; exchange = {args.exchange}" if wl.get("shard_volumes") else
                                                       f"{world} GPUs: cube-map tile ranges, light-map z-slabs, interleaved row stripes; exchange = {args.exchange}"),
            "e2e_inputs": "per-frame matrices (PerObject records) from pinned host memory; result = tone-mapped RGBA8 frame read back to pinned host "
                          "memory every step (Present, 3 frames in flight as in the reference's frame loop)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg4", choices=sorted(WORKLOADS))
    ap.add_argument("--exchange", default="fused", choices=["fused", "collective"])
    ap.add_argument("--work-graph", action="store_true", help="Render(..., useWorkGraph = true): cull inside the view-march launch (one GPU, not pipelined)")
    ap.add_argument("--cpu-baseline-frames", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-budget", type=float, default=150.0, help="seconds of CPU time the reference arm may use")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.warmup < 3:
        args.warmup = 3

    if args.impl == "reference":
        line = run_reference(args, wl, rank, world)
        if line:
            print(json.dumps(line), flush=True)
        return

    import torch
    import torch.distributed as dist
    from multivolumes_b200 import MultiRayCaster, PinnedBuffer, scene
    from multivolumes_b200.dist import CudaExchange, ShardedRenderer
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product has no CPU path (use --impl reference for the CPU oracle)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    hbm_peak, peak_src = peaks()
    tex_peak_gfetch, l2_peak_gbs = tex_peak()

    if wl.get("shard_volumes") and (world < 2 or (wl["n"] * wl["g"] ** 3 * 8 / world > 150e9)):
        raise SystemExit(f"bench.py: workload {args.workload} shards {wl['n'] * wl['g'] ** 3 * 8 / 1e9:.0f} GB of volumes over the ranks; {world} GPU(s) cannot hold it")

    def make_caster():
        c = MultiRayCaster(device=local_rank, count_samples=False, time_passes=False, density_only=bool(wl.get("density_only")),
                           shard_volumes=(rank, world, wl["proxy"]) if wl.get("shard_volumes") else None,
                           grid_size=wl["g"], light_grid_size=wl["l"], num_volumes=wl["n"], num_volume_srcs=wl.get("srcs"), width=wl["w"], height=wl["h"])
        build_scene(c, wl, scene, sky_coeffs(c))
        return c

    _sky = []

    def sky_coeffs(c):
        if not _sky:
            _sky.append(c.TransformSH(scene.procedural_sky(64)))
        return _sky[0]

    c = make_caster()
    stream = torch.cuda.Stream()
    c.SetStream(stream.cuda_stream)      # order the caster's kernels with torch's events / NCCL on one stream
    with torch.cuda.stream(stream):
        x = CudaExchange(c, rank, world) if world > 1 and args.exchange == "collective" else None
        r = ShardedRenderer(c, rank, world, mode=args.exchange, exchange=x, use_work_graph=args.work_graph)

        def barrier():
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()

        def frame(i):
            step_frame(c, wl, scene, i, lambda vp, svp, eye: r.render(vp, svp, eye, taa=wl["taa"]))

        # host inputs of every frame of the run (camera matrices, world matrices): a function of the frame index, computed here once
        # so that the timed loops measure the library (UpdateFrame's matrix work included), not numpy
        class _Sink:
            def SetVolumeWorldMatrices(self, m):
                pass
        for i in range(max(wl["n"], args.warmup + args.steps) + 1):
            camera(scene, wl, i); animate(_Sink(), wl, i)

        # ---- device-resident throughput ----
        # the light march fills ONE volume's light map per frame (round robin over the visible list, CSRayMarchL.hlsl:29-33):
        # N untimed frames first, so that the timed frames are steady-state images whatever --warmup says
        for i in range(wl["n"]):
            frame(i)
        for i in range(args.warmup):
            frame(i)
        barrier()
        sampler = ClockSampler(local_rank) if rank == 0 else None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        acc = {k: 0.0 for k in ("cull", "ray_march_light", "ray_march_view", "resolve_oit", "postprocess")}
        stats_acc = {}
        t_wall0 = time.time()
        e0.record(stream)
        for i in range(args.steps):
            frame(args.warmup + i)
        e1.record(stream)
        barrier()
        t_wall1 = time.time()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ms_total = float(ms.item())
        clocks = None
        if sampler:
            sampler.window(t_wall0, t_wall1)
            clocks = sampler.stop()

        # ---- per-pass device time and work counters: two more passes over the same frames ----
        # (a) CUDA events on the caster's stream around every pass, with the same kernels as the timed loop;
        # (b) the counting variants of the kernels (samples, fetches, rays), not timed.
        n_prof = min(args.steps, 50)
        for count, timed in ((False, True), (True, False)):
            c.SetInstrumentation(count, timed)
            for i in range(3):            # kernels are loaded on first launch (lazy module loading): keep that out of the averages
                frame(i)
            c.Sync()
            for i in range(n_prof):
                frame(args.warmup + i)
                if timed:
                    t = c.GetTimings()
                    for k in acc:
                        acc[k] += t[k]
                else:
                    st = c.GetStats()
                    for k, v in st.items():
                        stats_acc[k] = stats_acc.get(k, 0) + v
        barrier()
        c.SetInstrumentation(False, False)
        per_pass = {k: v / n_prof for k, v in acc.items()}
        samples_frame = (stats_acc["view_samples"] + stats_acc["direct_samples"] + stats_acc["light_samples"]) / n_prof
        if world > 1:   # whole-job samples: sum over ranks
            tsum = torch.tensor([samples_frame], device="cuda", dtype=torch.float64)
            dist.all_reduce(tsum)
            samples_frame = float(tsum.item())

        # ---- end to end: host matrices in, RGBA8 frame out, every step ----
        # The frame loop of the reference application: UpdateFrame from host matrices, Render, Postprocess, Present, with
        # FrameCount = 3 frames in flight (MultiRayCaster.h:52). Present = asynchronous read-back of the tone-mapped frame
        # into pinned host memory; the host blocks on the frame presented three steps earlier before reusing its buffer.
        # Every step's H2D and D2H copies complete inside the timed region (all slots are drained before the clock stops).
        # N > 1: every rank reads ITS rows back over its own PCIe link into one host frame shared by the ranks (POSIX shared
        # memory, page-locked in every process), so no rank carries the whole 33 MB frame.
        slots = 3
        outs = r.present_buffers(slots)
        n_e2e = min(args.steps, 100)

        no_present = os.environ.get("MV_BENCH_E2E") == "nopresent"     # diagnostic: the same wall-clock loop without the read-back

        def e2e_step(i):
            frame(i)
            if not no_present:
                r.present(outs, i % slots)

        def drain():
            for k in range(slots):
                c.PresentWait(k)

        for i in range(3):
            e2e_step(i)
        drain()
        barrier()
        t0 = time.perf_counter()
        for i in range(n_e2e):
            e2e_step(args.warmup + i)
        drain()
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e_fps = n_e2e / float(dt.item())
        checksum = int(sum(int(o.array[::16, ::16].astype(np.uint64).sum()) for o in outs)) if rank == 0 else 0
        # the same loop with a blocking read-back every step (no frames in flight), for comparison
        t0 = time.perf_counter()
        for i in range(n_e2e):
            frame(args.warmup + i)
            r.present(outs, 0)
            c.PresentWait(0)
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e_blocking_fps = n_e2e / float(dt.item())

    if rank == 0:
        fps = args.steps / (ms_total / 1000.0)
        view_ms = per_pass["ray_march_view"]
        vs, vl, vr, vk = (stats_acc[k] / n_prof for k in ("view_samples", "view_light_fetches", "view_rays", "view_skipped"))
        # SURVEY.md 8(d), per unit: a march sample = 1 trilinear density fetch (8 B texel; 2 B in the density-only storage)
        # + 1 trilinear light-map fetch (8 B) when alpha > 0.01; per ray a 4 B depth read and 8 B + 4 B written
        alg_bytes = vs * (2 if wl.get("density_only") else 8) + vl * 8 + vr * 16
        sec = view_ms / 1000.0 if view_ms > 0 else float("inf")
        fetches = (vs + vl) / sec / 1e9            # algorithmic: every sample of the reference's loop
        issued = (vs - vk + vl) / sec / 1e9        # what the texture unit was actually asked for (empty-space bricks)
        prof = ncu_summary(args.workload) if world == 1 else None
        line = {"metric": "frames_per_s", "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": workload_config(args, wl, world),
                "samples_per_s": samples_frame * fps, "samples_per_frame": samples_frame,
                "e2e": {"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": wl["n"] * 224, "d2h_bytes_per_step": wl["w"] * wl["h"] * 4,
                        "steps": n_e2e, "checksum": checksum,
                        "frames_in_flight": 3, "blocking_readback_value": e2e_blocking_fps},
                "gpu_launches": args.steps * launches_per_frame(wl, world, args.exchange, args.work_graph),
                "clocks": clocks,
                "per_pass_ms": per_pass, "per_pass_note": "rank 0, instrumented pass (CUDA events around each pass; at N > 1 the light and view marches include their peer barriers)",
                # the bound BASELINE.json's north_star names for the ray march: the texture unit's filtered-fetch rate
                "roofline": {"kernel": "k_ray_march_v", "bound": "tex", "achieved": fetches, "peak": tex_peak_gfetch, "unit": "Gfetch/s",
                             "frac": fetches / tex_peak_gfetch, "peak_source": "profiles/r01_tex_probe.json (tools/tex_probe.cu on this pool's B200: trilinear RGBA16F, L1-resident)",
                             "achieved_def": "(march samples + light-map fetches) of one launch, counted by the kernel's counting variant, / its CUDA-event duration; "
                                             "SURVEY.md 8(d): 1 trilinear volume fetch per sample + 1 trilinear light fetch when alpha > 0.01",
                             "issued": {"achieved": issued, "frac": issued / tex_peak_gfetch,
                                        "note": "fetches the texture unit actually served: samples inside bricks known to be empty are not fetched (same results)"},
                             "launch_ms": view_ms, "samples_per_launch": vs, "light_fetches_per_launch": vl, "rays_per_launch": vr, "skipped_per_launch": vk,
                             "traffic": prof.get("k_ray_march_v", {}).get("dram_bytes") if prof else None,
                             "ncu": prof.get("k_ray_march_v") if prof else None,
                             "l2": {"achieved": issued * 64.0, "peak": l2_peak_gbs, "unit": "GB/s", "frac": issued * 64.0 / l2_peak_gbs,
                                    "note": "64 B trilinear RGBA16F footprint per issued fetch against the measured L2 read bandwidth (an upper bound on L2 -> L1 traffic: most footprints hit L1)"},
                             "hbm": {"achieved": alg_bytes / sec / 1e9, "peak": hbm_peak, "unit": "GB/s", "frac": alg_bytes / sec / 1e9 / hbm_peak,
                                     "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes}}}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"], line["parity_checked"], line["parity"] = cpu_baseline(args, wl, make_caster)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
    del outs
    r.close()
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline(args, wl, make_caster):
    """The oracle on this box's host cores, a bounded sample of the same workload (whole frames) — and, since those frames
    are rendered anyway, the parity check of the product on the bench's own workload: a FRESH caster renders the same
    frames from the same start state and its composited frame / RGBA8 back buffer are compared with the oracle's."""
    from multivolumes_b200 import scene
    o = make_oracle(wl, host_threads())
    build_scene(o, wl, scene, o.TransformSH(scene.procedural_sky(64)))
    p = make_caster()
    sky = o.TransformSH(scene.procedural_sky(64)) if wl["sh"] else None
    for c in (o, p):
        c.SetSH(sky)                       # identical coefficients on both sides (the SH projection has its own parity test)
    n, t_cpu, samples = 0, 0.0, 0
    while n < args.cpu_baseline_frames and (n == 0 or t_cpu < 25.0):
        i = args.warmup + n
        t0 = time.perf_counter()
        step_frame(o, wl, scene, i, lambda vp, svp, eye: (o.UpdateFrame(vp, svp, eye), o.RenderEnvironment(), o.Render(), o.Postprocess(wl["taa"])))
        t_cpu += time.perf_counter() - t0
        step_frame(p, wl, scene, i, lambda vp, svp, eye: (p.UpdateFrame(vp, svp, eye), p.RenderEnvironment(), p.Render(), p.Postprocess(wl["taa"])))
        st = o.GetStats(); samples += st["view_samples"] + st["direct_samples"] + st["light_samples"]
        n += 1
    fo, fp = o.ReadFrame().astype(np.float32), p.ReadFrame().astype(np.float32)
    (_, bo), (_, bp) = o.ReadPost(), p.ReadPost()
    err = np.abs(fo - fp) / np.maximum(1.0, np.abs(fo))
    mse = float(np.mean((fo - fp) ** 2))
    parity = {"frames": n, "visible_lists_equal": bool(np.array_equal(o.ReadVisible(), p.ReadVisible()) and np.array_equal(o.ReadCubeVolumes(), p.ReadCubeVolumes())),
              "frame_bit_exact": bool(np.array_equal(o.ReadFrame().view(np.uint16), p.ReadFrame().view(np.uint16))),
              "frame_max_abs": float(err.max()), "frame_psnr_db": None if mse == 0 else float(10 * np.log10(max(float(np.abs(fo).max()), 1.0) ** 2 / mse)),
              "rgba8_max_diff": int(np.abs(bo.astype(int) - bp.astype(int)).max()),
              "bar": "visible lists bit-exact; frame max-abs <= 2e-3 and PSNR >= 50 dB (BASELINE.json north_star)"}
    ok = parity["visible_lists_equal"] and parity["frame_max_abs"] <= 2e-3 and (parity["frame_psnr_db"] is None or parity["frame_psnr_db"] >= 50.0)
    p.close()
    return ({"value": n / t_cpu, "unit": "frames/s", "cores": o.GetStats()["threads"], "kind": "port",
             "sample": f"{n} whole frames of the same workload (frames {args.warmup}..{args.warmup + n - 1} of the orbit, from a fresh caster)",
             "samples_per_s": samples / t_cpu}, bool(ok), parity)


if __name__ == "__main__":
    main()
