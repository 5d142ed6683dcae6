"""Small driver for ncu captures of the non-blur stages: DoG, NMI, warp, merge, pyramid on a 6000^2 pair."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from benchdata import synth_pair_large  # noqa: E402
from microaligner_b200 import ops  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 6000
ref, mov = synth_pair_large(N, N, 0)
r, m = torch.from_numpy(ref).cuda(), torch.from_numpy(mov).cuda()
flow = (torch.randn((N, N, 2), device="cuda") * 3).contiguous()
for _ in range(2):
    d1, d2 = ops.dog_u8(r), ops.dog_u8(m)
    s = ops.nmi_chunks(d1, d2, 1000 * 1000)
    w = ops.warp_tiles(m, flow, 1000, 100)
    mg = ops.merge_flows_tiles(flow, flow, 1000, 100)
    p = ops.pyr_down(r)
    u = ops.pyr_up_flow(flow[: N // 2, : N // 2].contiguous(), (N, N), 2.0)
torch.cuda.synchronize()
print("ok", float(s.mean()))
