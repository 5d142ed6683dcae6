"""Diagnostic: which host ranges can cudaHostRegister page-lock on this box (size, backing), and at what speed."""
import mmap
import os
import resource
import time

import numpy as np
import torch

torch.cuda.init()
rt = torch.cuda.cudart()
print("RLIMIT_MEMLOCK", resource.getrlimit(resource.RLIMIT_MEMLOCK))
os.system("df -h /dev/shm /tmp | cat; free -g | head -2")
for kind in ("anon", "memfd"):
    for gb in (1, 3, 5, 9, 12):
        n = gb << 30
        try:
            if kind == "anon":
                a = np.empty(n, np.uint8)
            else:
                fd = os.memfd_create("diag")
                os.ftruncate(fd, n)
                a = np.memmap(f"/proc/{os.getpid()}/fd/{fd}", dtype=np.uint8, mode="r+", shape=(n,))
            t = torch.from_numpy(a)
            t0 = time.perf_counter()
            a[::4096] = 1          # touch
            t1 = time.perf_counter()
            rc = int(rt.cudaHostRegister(t.data_ptr(), n, 0))
            t2 = time.perf_counter()
            msg = f"{kind} {gb} GiB: touch {t1 - t0:.2f}s register rc={rc} in {t2 - t1:.2f}s"
            if rc == 0:
                d = torch.empty(n, dtype=torch.uint8, device="cuda")
                torch.cuda.synchronize()
                t3 = time.perf_counter()
                d.copy_(t, non_blocking=True)
                torch.cuda.synchronize()
                t4 = time.perf_counter()
                t.copy_(d, non_blocking=True)
                torch.cuda.synchronize()
                t5 = time.perf_counter()
                msg += f"  H2D {n / (t4 - t3) / 1e9:.1f} GB/s  D2H {n / (t5 - t4) / 1e9:.1f} GB/s"
                rt.cudaHostUnregister(t.data_ptr())
                del d
            print(msg, flush=True)
            del t, a
        except Exception as e:  # noqa: BLE001
            print(kind, gb, "EXC", repr(e)[:200], flush=True)
            break
# pageable copy rate for comparison
a = np.ones(2 << 30, np.uint8)
t = torch.from_numpy(a)
d = torch.empty(2 << 30, dtype=torch.uint8, device="cuda")
for _ in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter(); d.copy_(t); torch.cuda.synchronize(); t1 = time.perf_counter()
    t.copy_(d); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"pageable 2 GiB: H2D {2.147 / (t1 - t0):.1f} GB/s  D2H {2.147 / (t2 - t1):.1f} GB/s")
