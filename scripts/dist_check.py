"""Multi-GPU parity check (NCCL): sharded register()+warp() must equal the single-GPU result bit for bit.
torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29512 scripts/dist_check.py [size]"""
import contextlib
import io
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from benchdata import synth_pair_large  # noqa: E402
from microaligner_b200 import OptFlowRegistrator, Warper, parallel  # noqa: E402


def run(ref, mov, kw):
    reg = OptFlowRegistrator()
    for k, v in kw.items():
        setattr(reg, k, v)
    reg.ref_img, reg.mov_img = ref, mov
    with contextlib.redirect_stdout(io.StringIO()):
        flow = reg.register()
    w = Warper()
    w.tile_size, w.overlap = kw.get("tile_size", 1000), kw.get("overlap", 100)
    w.image, w.flow = mov, flow
    return flow, w.warp(), [d["better"] for d in reg.decisions]


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
    ok = True
    for kw in (dict(use_full_res_img=True, use_dog=True, num_pyr_lvl=3), dict(use_full_res_img=False, num_pyr_lvl=2)):
        ref_h, mov_h = synth_pair_large(S, S + 300, seed=1)
        ref, mov = torch.from_numpy(ref_h).to(dev), torch.from_numpy(mov_h).to(dev)
        parallel.init(dist.group.WORLD)
        f1, i1, d1 = run(ref, mov, kw)
        parallel.init(None)
        f0, i0, d0 = run(ref, mov, kw)
        same = bool(torch.equal(f0, f1)) and bool(torch.equal(i0, i1)) and d0 == d1
        # the drop-in numpy API under the process group: rows up / rows down per rank, full node-shared arrays back
        parallel.init(dist.group.WORLD)
        f2, i2, d2 = run(ref_h, mov_h, kw)
        same_np = bool(np.array_equal(f2, f0.cpu().numpy())) and bool(np.array_equal(i2, i0.cpu().numpy())) and d2 == d0
        del f2, i2
        print(f"rank {dist.get_rank()}/{dist.get_world_size()} {kw}: decisions {d1} device path identical={same} "
              f"numpy API identical={same_np}", flush=True)
        ok = ok and same and same_np
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
