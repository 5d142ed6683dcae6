"""End to end through the numpy API with PAGEABLE input arrays (what a drop-in caller passes), one GPU, 20000^2 pair:
call 1 uploads staged, call 2 page-locks the caller's arrays (ops.pin_on_reuse), later calls run at DMA line rate.
Prints the wall time of every call and what happened to the arrays."""
import contextlib
import io
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from benchdata import synth_pair_large  # noqa: E402
from microaligner_b200 import OptFlowRegistrator, Warper  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
ref, mov = synth_pair_large(S, S, seed=1)           # plain np.empty arrays: pageable
pinned = lambda a: bool(torch.from_numpy(a.reshape(-1)[:1]).is_pinned())      # noqa: E731
for call in range(1, 6):
    torch.cuda.synchronize()
    t = time.perf_counter()
    reg = OptFlowRegistrator()
    reg.num_pyr_lvl, reg.num_iterations, reg.tile_size, reg.overlap, reg.use_full_res_img, reg.use_dog = 4, 3, 1000, 100, True, False
    reg.ref_img, reg.mov_img = ref, mov
    with contextlib.redirect_stdout(io.StringIO()):
        flow = reg.register()
    w = Warper()
    w.tile_size, w.overlap, w.image, w.flow = 1000, 100, mov, flow
    img = w.warp()
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t) * 1e3
    print(f"call {call}: {ms:8.1f} ms  {S * S / ms / 1e3:7.0f} Mpx/s   inputs page-locked afterwards: ref {pinned(ref)} mov {pinned(mov)}  "
          f"checksum {int(np.asarray(img[::97, ::89], dtype=np.int64).sum())}", flush=True)
    del flow, img
