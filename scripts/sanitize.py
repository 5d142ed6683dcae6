"""Tiny end-to-end run for compute-sanitizer (memcheck / racecheck): every kernel once, small shapes.
compute-sanitizer --tool memcheck python scripts/sanitize.py"""
import contextlib
import io
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from benchdata import synth_pair  # noqa: E402
from microaligner_b200 import OptFlowRegistrator, Warper, ops  # noqa: E402

ref, mov = synth_pair(330, 410, 1, np.uint16)
reg = OptFlowRegistrator()
reg.num_pyr_lvl, reg.num_iterations, reg.tile_size, reg.overlap = 1, 2, 120, 16
reg.use_full_res_img, reg.use_dog = True, True
reg.ref_img, reg.mov_img = ref, mov
with contextlib.redirect_stdout(io.StringIO()):
    flow = reg.register()
w = Warper()
w.tile_size, w.overlap = 120, 16
w.image, w.flow = mov, flow
out = w.warp()
pages = [ops.to_device(ref), ops.to_device(mov)]
mip = ops.zmip_normalize_u8(pages)
# paths the small run above does not reach: uint8 images (scalar-store warp path on the odd width), an odd-width flow,
# untiled Farneback, full iteration windows, the affine resampler and the opt-in flow composition
import torch  # noqa: E402
r8, m8 = synth_pair(301, 333, 2, np.uint8)
d8r, d8m = ops.to_device(r8), ops.to_device(m8)
f8 = ops.farneback_tiles(d8m, d8r, 0, 0, 31, 2)
f8b = ops.farneback_tiles(d8m, d8r, 100, 12, 11, 2, full_windows=True)
w8 = ops.warp_tiles(d8m, f8, 100, 12)
mg = ops.merge_flows_tiles(f8, f8b, 100, 12)
cf = ops.compose_flows_rows(f8, f8b, (0, 301), torch.empty_like(f8))
af = ops.warp_affine(d8m, np.linalg.pinv(np.array([[0.99, 0.02, 1.5], [-0.02, 0.99, -2.0], [0, 0, 1]])), (320, 350), 9, 8)
torch.cuda.synchronize()
print("sanitize run ok", float(np.abs(flow).max()), int(out.max()), int(mip.max()))
