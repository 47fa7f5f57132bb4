"""Multi-GPU frame driver: one process per GPU over torch.distributed.

Partition (BASELINE.json north_star; the single-adapter reference has no counterpart):
  * the cull is replicated on every rank (N <= a few hundred volumes: cheaper than any collective);
  * by volume — rank r marches the cube maps of the volumes v with v % world == r (collective mode) or, with the peers
    mapped (fused mode: every texel is stored into all arenas by whoever marches it), one contiguous part of EVERY
    cube-map volume's tile list (``view_march_tile_range``; the cull kernel evaluates the same integers); rank r also
    fills z-slab r of the frame's light map;
  * by screen band — rank r resolves the OIT and post-processes rows [H r / world, H (r + 1) / world).
Exchange steps: the light-map slabs and the marched cube maps must reach every rank before the
resolve; the finished bands must reach rank 0.

Two ways to move them (``mode``):
  * ``"fused"``  — CUDA-IPC maps every rank's exchange block into every peer; the march / light /
    post-process kernels store their results straight into all peers' blocks over NVLink and a
    device-side flag barrier orders the phases. No collective is called in the frame.
  * ``"collective"`` — torch.distributed collectives (NCCL on GPUs; gloo in the CPU tests) between the
    passes: all-gather of the light slabs, broadcast of every visible cube map from its owner, gather
    of the bands. This is the baseline the fused mode is measured against, and the only mode the CPU
    tests can run (they drive the oracle through the same class).

The class is written against the MultiRayCaster surface (``CasterBase``) plus a small ``Exchange``
adapter, so the host logic is identical for the CUDA product and for the test oracle.
"""
import numpy as np


def row_band(height, rank, world):
    return (height * rank) // world, (height * (rank + 1)) // world


def light_slab(L, rank, world):
    d = (L + world - 1) // world
    z0 = min(L, rank * d)
    return z0, min(L, z0 + d)


def owner_of(volume, world):
    return volume % world


def view_march_tile_range(tiles, k, rank, world):
    """Fused mode: the 8x4-texel tiles [begin, end) of the k-th cube-map volume (in march order) that `rank` marches. The
    volume's tile list (face-major, row-major; `tiles` = tiles per face x visible faces) is cut into `world` contiguous
    parts and the parts are handed out rotated by k, so that no rank always gets the same face of every volume. Mirrors
    cull_body in csrc/k_cull.cuh (same integer arithmetic); the parts of the `world` ranks tile [0, tiles) exactly."""
    part = (rank + world - k % world) % world
    return tiles * part // world, tiles * (part + 1) // world


class CudaExchange:
    """Collectives on regions of the product's exchange block (device memory), NCCL through torch."""

    def __init__(self, caster, rank, world, group=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group = torch, dist, group
        self.c, self.rank, self.world = caster, rank, world
        ptr, nbytes = caster.ExchangeBlock()
        self.lay = caster.ExchangeLayout()

        class _Block:   # zero-copy view of the cudaMalloc'ed block as a torch tensor
            __cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3, "strides": None}
        self.block = torch.as_tensor(_Block(), device=f"cuda:{caster.device}")
        assert self.block.data_ptr() == ptr

    def region(self, off, nbytes):
        return self.block[off:off + nbytes]

    def all_gather_light(self):
        L = self.c.L
        d = (L + self.world - 1) // self.world
        slab_bytes = d * L * L * 8
        off = self.lay.light_staging_offset
        whole = self.region(off, slab_bytes * self.world)
        mine = self.region(off + slab_bytes * self.rank, slab_bytes)
        self.dist.all_gather_into_tensor(whole, mine, group=self.group)
        self.c.LightCommit()

    def broadcast_cubes(self, cube_volumes, attribs):
        ops = []
        for v in cube_volumes:
            co, cb, do, db = self.c.CubeRegion(int(v), int(attribs[v][0]))
            src = owner_of(int(v), self.world)
            for off, n in ((co, cb), (do, db)):
                t = self.region(off, n)
                ops.append(self.dist.broadcast(t, src=src, group=self.group, async_op=True))
        for w in ops:
            w.wait()

    def gather_back_buffer(self, row0, row1, bands):
        W = self.c.W
        off = self.lay.back_buffer_offset
        mine = self.region(off + row0 * W * 4, (row1 - row0) * W * 4)
        if self.rank == 0:
            outs = [self.region(off + a * W * 4, (b - a) * W * 4) for a, b in bands]
            self.dist.gather(mine, outs, dst=0, group=self.group)
        else:
            self.dist.gather(mine, None, dst=0, group=self.group)


class HostExchange:
    """Same steps through host read-backs and gloo (CPU tests with the oracle)."""

    def __init__(self, caster, rank, world, group=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group = torch, dist, group
        self.c, self.rank, self.world = caster, rank, world

    def all_gather_light(self):
        c, L = self.c, self.c.L
        vol = c.GetStats()["light_volume"]
        lm = np.ascontiguousarray(c.ReadLightMap(vol).view(np.uint16))
        d = (L + self.world - 1) // self.world
        for r in range(self.world):
            z0, z1 = light_slab(L, r, self.world)
            if z1 <= z0:
                continue
            t = self.torch.from_numpy(lm[z0:z1].copy().view(np.uint8))
            self.dist.broadcast(t, src=r, group=self.group)
            slab = np.ascontiguousarray(t.numpy())
            c._ck(c.b.write_lightmap_slab(c.h, vol, z0, z1, slab.ctypes.data), "write_lightmap_slab")

    def broadcast_cubes(self, cube_volumes, attribs):
        c = self.c
        for v in cube_volumes:
            mip = int(attribs[v][0])
            rgba, depth = c.ReadCubeMap(int(v), mip)
            t1 = self.torch.from_numpy(np.ascontiguousarray(rgba).view(np.uint8))
            t2 = self.torch.from_numpy(np.ascontiguousarray(depth).view(np.uint8))
            src = owner_of(int(v), self.world)
            self.dist.broadcast(t1, src=src, group=self.group)
            self.dist.broadcast(t2, src=src, group=self.group)
            a, b = np.ascontiguousarray(t1.numpy()), np.ascontiguousarray(t2.numpy())
            c._ck(c.b.write_cubemap(c.h, int(v), mip, a.ctypes.data, b.ctypes.data), "write_cubemap")

    def gather_back_buffer(self, row0, row1, bands):
        c = self.c
        taa, rgba8 = c.ReadPost()
        frame = c.ReadFrame()
        for what, img in ((0, frame), (1, taa), (2, rgba8)):
            mine = self.torch.from_numpy(np.ascontiguousarray(img[row0:row1]).view(np.uint8))
            outs = [self.torch.empty((b - a,) + tuple(mine.shape[1:]), dtype=mine.dtype) for a, b in bands] if self.rank == 0 else None
            self.dist.gather(mine, outs, dst=0, group=self.group)
            if self.rank == 0:
                for (a, b), t in zip(bands, outs):
                    if b > a:
                        rows = np.ascontiguousarray(t.numpy())
                        c._ck(c.b.write_rows(c.h, what, a, b, rows.ctypes.data), "write_rows")


class ShardedRenderer:
    """Renders one frame of a scene on `world` ranks; rank 0 ends up with the whole frame."""

    STRIPE_HEIGHT = 30   # + 2 halo rows = two 16-row OIT tile rows

    def __init__(self, caster, rank, world, mode="collective", exchange=None, group=None, stripes=True, use_work_graph=False):
        assert mode in ("fused", "collective")
        assert not (use_work_graph and world > 1), "the work-graph path (cull inside the march launch) is a one-GPU path"
        self.c, self.rank, self.world, self.mode, self.group = caster, rank, world, mode, group
        self.use_work_graph = use_work_graph
        self.bands = [row_band(caster.H, r, world) for r in range(world)]
        self.row0, self.row1 = self.bands[rank]
        caster.SetShard(rank, world)
        caster.SetRowBand(self.row0, self.row1)
        self.x = exchange
        if world > 1 and mode == "fused":
            if stripes:   # interleaved row stripes instead of one band: the expensive pixels (direct marches) cluster
                caster.SetRowBand(0, caster.H)
                caster.SetRowStripes(self.STRIPE_HEIGHT)
            self._map_peers()

    def _map_peers(self):
        import torch.distributed as dist
        handles = [None] * self.world
        dist.all_gather_object(handles, self.c.IpcExport(), group=self.group)
        for r, h in enumerate(handles):
            if r != self.rank:
                self.c.IpcImport(r, h)
        dist.barrier(group=self.group)

    # --- Present of the frame loop (end-to-end path) ---
    def present_buffers(self, slots):
        """`slots` whole-frame RGBA8 host buffers for Present. One GPU: pinned memory of this process. Sharded (fused): ONE
        POSIX shared-memory segment mapped and page-locked by every rank's process, so that each rank reads its own rows
        back over its own PCIe link (mv_present_rows_async) and rank 0 sees the assembled frame without carrying it."""
        from .caster import PinnedBuffer
        c = self.c
        if self.world == 1 or self.mode != "fused":
            return [PinnedBuffer((c.H, c.W, 4), np.uint8) for _ in range(slots)] if self.rank == 0 else [None] * slots
        import ctypes
        import torch.distributed as dist
        from multiprocessing import shared_memory
        nbytes = slots * c.H * c.W * 4
        name = [None]
        if self.rank == 0:
            self._shm = shared_memory.SharedMemory(create=True, size=nbytes)
            name[0] = self._shm.name
        dist.broadcast_object_list(name, src=0, group=self.group)
        if self.rank != 0:
            self._shm = shared_memory.SharedMemory(name=name[0])
            try:      # the creator unlinks the segment; attached processes must not report it as leaked
                from multiprocessing import resource_tracker
                resource_tracker.unregister(self._shm._name, "shared_memory")
            except Exception:
                pass
        base = ctypes.addressof(ctypes.c_char.from_buffer(self._shm.buf))
        c._ck(c.b.host_register(base, nbytes), "host_register")
        self._shm_base = base
        dist.barrier(group=self.group)

        class _Slot:
            pass
        outs = []
        for k in range(slots):
            o = _Slot()
            o.ptr = base + k * c.H * c.W * 4
            o.array = np.frombuffer(self._shm.buf, dtype=np.uint8, count=c.H * c.W * 4, offset=k * c.H * c.W * 4).reshape(c.H, c.W, 4)
            outs.append(o)
        return outs

    def present(self, outs, slot):
        c = self.c
        if self.world > 1 and self.mode == "fused":
            c.PresentRowsAsync(outs[slot].ptr, slot)
        else:
            c.PresentAsync(outs[slot].ptr if self.rank == 0 else None, slot)

    def close(self):
        if getattr(self, "_shm", None) is not None:
            try:
                self.c.b.host_unregister(self._shm_base)
            except Exception:
                pass
            shm, self._shm = self._shm, None
            try:
                shm.close()
                if self.rank == 0:
                    shm.unlink()
            except Exception:
                pass

    def render(self, view_proj, shadow_vp, eye, taa=True, reset_color=True):
        c = self.c
        c.UpdateFrame(view_proj, shadow_vp, eye)
        if reset_color:
            c.RenderEnvironment()
        if self.world == 1:
            c.Render(use_work_graph=self.use_work_graph)
            c.Postprocess(taa)
            return
        if self.mode == "fused":
            c.Render()              # cull -> light slab -> barrier + commit -> march (peer stores) -> barrier -> OIT band
            c.Postprocess(taa)      # band; RGBA8 rows also into rank 0's back buffer, TAA rows into every peer's history; closing barrier
            return
        c.Cull()
        c.RayMarchL(-1)
        self.x.all_gather_light()
        c.RayMarchV()
        cubes, att = c.ReadCubeVolumes(), c.ReadAttribs()      # replicated cull -> same lists on every rank
        self.x.broadcast_cubes(cubes, att)
        c.ResolveOIT()
        c.AdvanceFrame()
        c.Postprocess(taa)
        self.x.gather_back_buffer(self.row0, self.row1, self.bands)
