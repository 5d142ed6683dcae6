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
print("sanitize run ok", float(np.abs(flow).max()), int(out.max()), int(mip.max()))
