"""Diagnostic under torchrun: cudaHostRegister of a node-shared memfd array, before / after NCCL init, whole vs 1-GiB chunks."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

local = int(os.environ.get("LOCAL_RANK", "0"))
rank = int(os.environ.get("RANK", "0"))
torch.cuda.set_device(local)
rt = torch.cuda.cudart()
dist.init_process_group("gloo")


def shared(n):
    name = [None]
    fd = None
    if rank == 0:
        fd = os.memfd_create("d")
        os.ftruncate(fd, n)
        name[0] = f"/proc/{os.getpid()}/fd/{fd}"
    dist.broadcast_object_list(name, src=0)
    a = np.memmap(name[0], dtype=np.uint8, mode="r+", shape=(n,))
    dist.barrier()
    if fd is not None:
        os.close(fd)
    return a


def try_reg(tag, a, chunk=None):
    t = torch.from_numpy(a)
    n = a.nbytes
    t0 = time.perf_counter()
    if chunk is None:
        rcs = [int(rt.cudaHostRegister(t.data_ptr(), n, 0))]
    else:
        rcs = [int(rt.cudaHostRegister(t.data_ptr() + o, min(chunk, n - o), 0)) for o in range(0, n, chunk)]
    dt = time.perf_counter() - t0
    print(f"[rank {rank}] {tag}: {n / 1e9:.1f} GB chunk={chunk} rc={sorted(set(rcs))} ({rcs.count(0)}/{len(rcs)} ok) {dt:.2f}s", flush=True)
    for i, rc in enumerate(rcs):
        if rc == 0:
            rt.cudaHostUnregister(t.data_ptr() + (0 if chunk is None else i * chunk))


x = torch.zeros(1 << 20, device="cuda")
for n in (3 * 10 ** 9, 5 * 10 ** 9):
    a = shared(n)
    if rank == 0:
        a[::4096] = 1
    dist.barrier()
    try_reg("gloo only, whole", a)
    try_reg("gloo only, 1 GiB chunks", a, 1 << 30)
    del a
# now NCCL
pg = dist.new_group(backend="nccl")
dist.all_reduce(x, group=pg)
torch.cuda.synchronize()
for n in (3 * 10 ** 9, 5 * 10 ** 9, 20 * 10 ** 9):
    a = shared(n)
    if rank == 0:
        a[::4096] = 1
    dist.barrier()
    try_reg("after NCCL, whole", a)
    try_reg("after NCCL, 1 GiB chunks", a, 1 << 30)
    try_reg("after NCCL, 2 GiB chunks", a, 2 << 30)
    del a
big = torch.empty(100 << 30, dtype=torch.uint8, device="cuda")
a = shared(5 * 10 ** 9)
try_reg("after NCCL + 100 GiB device alloc, whole", a)
del big
torch.cuda.empty_cache()
try_reg("after freeing it, whole", a)
dist.destroy_process_group()
