"""Small driver for ncu: one tiled Farneback call (36 tiles of 1200^2, 3 iterations). Usage under gpurun:
ncu --set full --clock-control none --import-source on -k regex:fb_blur -c 4 -o gpurun_out/prof_blur python scripts/prof_farneback.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from benchdata import synth_pair_large  # noqa: E402
from microaligner_b200 import ops  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 6000
ref, mov = synth_pair_large(N, N, 0)
r, m = torch.from_numpy(ref).cuda(), torch.from_numpy(mov).cuda()
f = ops.farneback_tiles(m, r, 1000, 100, 99, 3)
torch.cuda.synchronize()
print("flow abs max", float(f.abs().max()))
