"""Host<->device copy rates on this box (pinned vs pageable), to budget the e2e path."""
import time

import torch

n = 800 * 1000 * 1000  # 3.2 GB of float32
dev = torch.device("cuda:0")
d = torch.empty(n, dtype=torch.float32, device=dev)
for pinned in (True, False):
    t0 = time.perf_counter()
    h = torch.empty(n, dtype=torch.float32, pin_memory=pinned)
    h.fill_(1.0)
    t_alloc = time.perf_counter() - t0
    for name, fn in (("H2D", lambda: d.copy_(h, non_blocking=True)), ("D2H", lambda: h.copy_(d, non_blocking=True))):
        fn(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 3
        print(f"pinned={pinned} {name}: {n * 4 / dt / 1e9:.1f} GB/s ({dt * 1e3:.1f} ms)   alloc+touch {t_alloc:.2f} s")
    del h
t0 = time.perf_counter(); h = torch.empty(n, dtype=torch.float32, pin_memory=True); print("2nd pinned alloc (cached?)", time.perf_counter() - t0)
