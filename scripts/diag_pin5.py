"""Diagnostic: memory zones of the box, and page-locking node-shared arrays in different orders (2 gloo ranks)."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from microaligner_b200 import ops, parallel  # noqa: E402

local = int(os.environ.get("LOCAL_RANK", "0"))
rank = int(os.environ.get("RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("gloo")
comm = parallel.init(dist.group.WORLD)
rt = torch.cuda.cudart()
torch.zeros(1, device="cuda")


def zones(tag):
    if rank:
        return
    out = {}
    node = zone = None
    for ln in open("/proc/zoneinfo"):
        p = ln.split()
        if ln.startswith("Node"):
            node, zone = p[1].rstrip(","), p[3]
        elif p and p[0] == "pages" and p[1] == "free":
            out[(node, zone)] = [int(p[2])]
        elif p and p[0] == "managed":
            out[(node, zone)].append(int(p[1]))
    s = ", ".join(f"n{k[0]}/{k[1]} free {v[0] * 4 >> 20}G of {v[1] * 4 >> 20}G" for k, v in out.items() if v[1] * 4 >> 20)
    mi = {ln.split(":")[0]: ln.split()[1] for ln in open("/proc/meminfo")}
    print(f"[zones] {tag}: {s}; Shmem {int(mi['Shmem']) >> 20}G Unevictable {int(mi['Unevictable']) >> 20}G", flush=True)


def reg(tag, arr, keep=False, who=(0, 1)):
    rc = None
    if rank in who:
        t = torch.from_numpy(arr)
        t0 = time.perf_counter()
        rc = int(rt.cudaHostRegister(t.data_ptr(), arr.nbytes, 0))
        dt = time.perf_counter() - t0
        if rc == 0 and not keep:
            rt.cudaHostUnregister(t.data_ptr())
        elif rc:
            ops._clear_cuda_error()
        print(f"[rank {rank}] {tag}: {arr.nbytes / 1e9:.1f} GB rc={rc} {dt:.1f}s", flush=True)
    comm.barrier()


S = 50000
zones("start")
e = comm.shared_host_empty((S, S, 2), np.float32)
reg("E flow-sized untouched, first thing", e, keep=True)
zones("after E")
c = comm.shared_host_empty((S, S), np.uint16)
d = comm.shared_host_empty((S, S), np.uint16)
reg("C untouched, kept registered", c, keep=True)
if rank == 0:
    from benchdata import synth_pair_large
    synth_pair_large(S, S, seed=0, out=(c, d))
comm.barrier()
zones("after filling C (registered) and D (not)")
reg("D filled by the thread pool, rank 0 first", d, keep=True, who=(0,))
reg("D then rank 1", d, keep=True, who=(1,))
zones("after D")
g = comm.shared_host_empty((S, S), np.uint16)
if rank == 0:
    g[:] = 3
comm.barrier()
reg("G memset, both ranks at once", g)
h = comm.shared_host_empty((S, S), np.uint16)
if rank == 0:
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(32) as ex:
        list(ex.map(lambda i: h[i * 500:(i + 1) * 500].fill(5), range(100)))
comm.barrier()
reg("H filled by 32 threads, both ranks at once", h)
zones("end")
dist.destroy_process_group()
