"""Small driver for ncu captures of the non-blur stages on a 6000^2 pair: DoG, NMI, warp (smooth 3-px flow, as in a
registration), merge, pyramid, z max-projection + normalise, affine resampling, opt-in flow composition."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from benchdata import synth_pair_large  # noqa: E402
from microaligner_b200 import ops  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 6000
ref, mov = synth_pair_large(N, N, 0)
r, m = torch.from_numpy(ref).cuda(), torch.from_numpy(mov).cuda()
import numpy as np  # noqa: E402
yy, xx = torch.meshgrid(torch.arange(N, device="cuda", dtype=torch.float32), torch.arange(N, device="cuda", dtype=torch.float32), indexing="ij")
flow = torch.stack([3 * torch.sin(2 * np.pi * yy / 512), 2 * torch.cos(2 * np.pi * xx / 512)], dim=-1).contiguous()
del yy, xx
tmat = np.array([[0.999, 0.02, 3.5], [-0.02, 0.999, -2.25]])
inv = np.linalg.pinv(np.append(tmat, [[0, 0, 1]], axis=0))
for _ in range(2):
    d1, d2 = ops.dog_u8(r), ops.dog_u8(m)
    s = ops.nmi_chunks(d1, d2, 1000 * 1000)
    w = ops.warp_tiles(m, flow, 1000, 100)
    mg = ops.merge_flows_tiles(flow, flow, 1000, 100)
    p = ops.pyr_down(r)
    u = ops.pyr_up_flow(flow[: N // 2, : N // 2].contiguous(), (N, N), 2.0)
    z = ops.zmip_normalize_u8([r, m, r])
    af = ops.warp_affine(m, inv, (N, N))
    cf = ops.compose_flows_rows(flow, flow, (0, N), torch.empty_like(flow))
torch.cuda.synchronize()
print("ok", float(s.mean()))
