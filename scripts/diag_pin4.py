"""Diagnostic under torchrun + NCCL: page-locking arrays from parallel.Comm.shared_host_empty, by how they were filled."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from microaligner_b200 import ops, parallel  # noqa: E402

local = int(os.environ.get("LOCAL_RANK", "0"))
rank = int(os.environ.get("RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
comm = parallel.init(dist.group.WORLD)
if rank == 0:
    os.system("cat /sys/kernel/mm/transparent_hugepage/shmem_enabled /sys/kernel/mm/transparent_hugepage/enabled; ulimit -a | grep -i 'lock\\|virtual'")
rt = torch.cuda.cudart()
S = int(sys.argv[1]) if len(sys.argv) > 1 else 50000


def reg(tag, arr):
    t = torch.from_numpy(arr)
    rc = int(rt.cudaHostRegister(t.data_ptr(), arr.nbytes, 0))
    if rc == 0:
        rt.cudaHostUnregister(t.data_ptr())
    else:
        ops._clear_cuda_error()
    print(f"[rank {rank}] {tag}: {arr.nbytes / 1e9:.1f} GB rc={rc}", flush=True)


a = comm.shared_host_empty((S, S), np.uint16)
reg("A untouched", a)
comm.barrier()
b = comm.shared_host_empty((S, S), np.uint16)
if rank == 0:
    b[:] = 7
comm.barrier()
reg("B memset by rank 0", b)
reg("A again (B exists)", a)
comm.barrier()
c = comm.shared_host_empty((S, S), np.uint16)
d = comm.shared_host_empty((S, S), np.uint16)
if rank == 0:
    from benchdata import synth_pair_large
    synth_pair_large(S, S, seed=0, out=(c, d))
comm.barrier()
reg("C filled by synth_pair_large (thread pool)", c)
reg("D filled by synth_pair_large (thread pool)", d)
e = comm.shared_host_empty((S, S, 2), np.float32)
reg("E flow-sized untouched", e)
f = np.memmap(f"/dev/shm/x{rank}", dtype=np.uint16, mode="w+", shape=(S, S))
reg("F private /dev/shm file untouched", f)
dist.destroy_process_group()
