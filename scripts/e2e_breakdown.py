"""Where the end-to-end (numpy API) time goes: per-phase wall clock with device synchronisation."""
import contextlib
import io
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from benchdata import synth_pair_large  # noqa: E402
from microaligner_b200 import OptFlowRegistrator, Warper, ops  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
ref = torch.empty((S, S), dtype=torch.uint16, pin_memory=True).numpy()
mov = torch.empty((S, S), dtype=torch.uint16, pin_memory=True).numpy()
synth_pair_large(S, S, 0, out=(ref, mov))


def T(label, fn):
    torch.cuda.synchronize(); t = time.perf_counter(); r = fn(); torch.cuda.synchronize()
    sys.stderr.write(f"  {label:34s} {1e3 * (time.perf_counter() - t):8.1f} ms\n"); return r


reg = OptFlowRegistrator(); reg.use_full_res_img = True
w = Warper()
for it in range(3):
    sys.stderr.write(f"iteration {it}\n")
    with contextlib.redirect_stdout(io.StringIO()):
        r_d = T("to_device(ref)", lambda: ops.to_device(ref))
        m_d = T("to_device(mov)", lambda: ops.to_device(mov))
        reg.ref_img, reg.mov_img = r_d, m_d
        f_d = T("register (device)", lambda: reg.register())
        f_h = T("to_host(flow) 3.2 GB", lambda: ops.to_host(f_d))
        f_d2 = T("to_device(flow)", lambda: ops.to_device(f_h))
        w.image, w.flow = m_d, f_d2
        o_d = T("warp (device)", lambda: w.warp())
        o_h = T("to_host(warped)", lambda: ops.to_host(o_d))
        reg.ref_img, reg.mov_img = ref, mov
        f2 = T("register (numpy in/out)", lambda: reg.register())
        w.image, w.flow = mov, f2
        o2 = T("warp (numpy in/out)", lambda: w.warp())
    sys.stderr.write("")
