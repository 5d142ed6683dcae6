"""Diagnostic: ops.pin_rows (register, touch-and-retry) on node-shared arrays in the order the bench uses them."""
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
torch.zeros(1, device="cuda")


def pin(tag, arr):
    t0 = time.perf_counter()
    ok = ops.pin_rows(arr)
    print(f"[rank {rank}] {tag}: {arr.nbytes / 1e9:.1f} GB pinned={ok} {time.perf_counter() - t0:.1f}s "
          f"is_pinned={torch.from_numpy(arr.reshape(-1)[:1]).is_pinned()}", flush=True)
    comm.barrier()


S = 50000
c = comm.shared_host_empty((S, S), np.uint16)
d = comm.shared_host_empty((S, S), np.uint16)
if rank == 0:
    from benchdata import synth_pair_large
    synth_pair_large(S, S, seed=0, out=(c, d))
comm.barrier()
e = comm.shared_host_empty((S, S, 2), np.float32)
pin("E flow-sized untouched", e)
pin("C filled", c)
pin("D filled", d)
f = comm.shared_host_empty((S, S), np.uint16)
pin("F image-sized untouched", f)
dist.destroy_process_group()
