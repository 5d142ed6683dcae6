"""Diagnostic: cudaHostRegister of node-shared memory mapped by two processes (memfd via /proc, /dev/shm file), by size."""
import os
import sys
import time

import numpy as np
import torch
import torch.multiprocessing as mp


def worker(rank, q_in, q_out):
    torch.cuda.init()
    rt = torch.cuda.cudart()
    while True:
        job = q_in.get()
        if job is None:
            return
        path, n, touch = job
        a = np.memmap(path, dtype=np.uint8, mode="r+", shape=(n,))
        if touch:
            a[rank * (n // 2):(rank + 1) * (n // 2):4096] = 1
        t = torch.from_numpy(a)
        t0 = time.perf_counter()
        rc = int(rt.cudaHostRegister(t.data_ptr(), n, 0))
        dt = time.perf_counter() - t0
        rate = None
        if rc == 0:
            d = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
            torch.cuda.synchronize(); t1 = time.perf_counter()
            t[rank * (1 << 30):(rank + 1) * (1 << 30)].copy_(d, non_blocking=True)
            torch.cuda.synchronize(); rate = (1 << 30) / (time.perf_counter() - t1) / 1e9
            rt.cudaHostUnregister(t.data_ptr())
        q_out.put((rank, rc, round(dt, 2), rate))
        del t, a


if __name__ == "__main__":
    mp.set_start_method("spawn")
    qs = [mp.Queue() for _ in range(2)]
    out = mp.Queue()
    ps = [mp.Process(target=worker, args=(r, qs[r], out)) for r in range(2)]
    for p in ps:
        p.start()
    for kind in ("memfd", "shmfile"):
        for gb in (3, 5, 8, 20):
            for touch in (True, False):
                n = gb * 10 ** 9
                if kind == "memfd":
                    fd = os.memfd_create("d")
                    os.ftruncate(fd, n)
                    path = f"/proc/{os.getpid()}/fd/{fd}"
                else:
                    path = f"/dev/shm/diag_{gb}_{int(touch)}"
                    with open(path, "wb") as f:
                        f.truncate(n)
                for q in qs:
                    q.put((path, n, touch))
                res = sorted(out.get() for _ in range(2))
                print(kind, gb, "GB touched" if touch else "GB untouched", res, flush=True)
                if kind == "memfd":
                    os.close(fd)
                else:
                    os.unlink(path)
    for q in qs:
        q.put(None)
    for p in ps:
        p.join()
