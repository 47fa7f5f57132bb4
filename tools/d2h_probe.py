#!/usr/bin/env python
"""Aggregate device-to-host bandwidth of N ranks copying concurrently (diagnostic for the end-to-end Present path):
each rank copies `mb` MB per iteration into (a) its own cudaHostAlloc'ed buffer, (b) its slice of ONE POSIX shared-memory
segment that every rank page-locked with cudaHostRegister (mv_host_register) — what ShardedRenderer.present_buffers uses."""
import ctypes, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch, torch.distributed as dist
from multiprocessing import shared_memory
from multivolumes_b200 import binding

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
mb = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
n = int(mb * 1e6)
src = torch.empty(n, dtype=torch.uint8, device="cuda")
own = torch.empty(n, dtype=torch.uint8).pin_memory()
name = [None]
if rank == 0:
    shm = shared_memory.SharedMemory(create=True, size=n * world)
    name[0] = shm.name
dist.broadcast_object_list(name, src=0)
if rank != 0:
    shm = shared_memory.SharedMemory(name=name[0])
base = ctypes.addressof(ctypes.c_char.from_buffer(shm.buf))
b = binding()
assert b.host_register(base, n * world) == 0
cudart = ctypes.CDLL("libcudart.so", mode=ctypes.RTLD_GLOBAL) if False else None
shared = torch.frombuffer(shm.buf, dtype=torch.uint8, count=n, offset=rank * n)


def run(dst, label, iters=200):
    s = torch.cuda.Stream()
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    with torch.cuda.stream(s):
        for _ in range(iters):
            dst.copy_(src, non_blocking=True)
    s.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"{label}: {world} ranks x {mb} MB x {iters}: per rank {n * iters / t.item() / 1e9:.1f} GB/s, aggregate {world * n * iters / t.item() / 1e9:.1f} GB/s", flush=True)


run(own, "own pinned buffer")
run(shared, "shared registered segment")
dist.barrier()
del shared
b.host_unregister(base)
shm.close()
if rank == 0:
    shm.unlink()
dist.destroy_process_group()
