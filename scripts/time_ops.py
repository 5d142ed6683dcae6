"""Per-operator device timings (CUDA events) on representative sizes. Usage: python scripts/time_ops.py [N]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from microaligner_b200 import ops  # noqa: E402


def timeit(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 6000
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(0)
    img = torch.randint(0, 60000, (N, N), device=dev, generator=g, dtype=torch.int32).to(torch.uint16)
    img2 = torch.randint(0, 60000, (N, N), device=dev, generator=g, dtype=torch.int32).to(torch.uint16)
    flow = (torch.randn((N, N, 2), device=dev, generator=g) * 3).contiguous()
    flow2 = (torch.randn((N, N, 2), device=dev, generator=g) * 3).contiguous()
    px = N * N
    T, ov = 1000, 100
    nt = ops.n_tiles(N, N, T)
    tpx = nt * 1200 * 1200
    rows = []
    t = timeit(lambda: ops.warp_tiles(img, flow, T, ov)); rows.append(("warp_tiles u16", t, px * 12 / t / 1e6, px / t / 1e3))
    t = timeit(lambda: ops.pyr_down(img)); rows.append(("pyr_down u16", t, px * 2.5 / t / 1e6, px / t / 1e3))
    half = flow[: N // 2, : N // 2].contiguous()
    t = timeit(lambda: ops.pyr_up_flow(half, (N, N), 2.0)); rows.append(("pyr_up_flow", t, px * 10 / t / 1e6, px / t / 1e3))
    t = timeit(lambda: ops.merge_flows_tiles(flow, flow2, T, ov)); rows.append(("merge_flows", t, px * 40 / t / 1e6, px / t / 1e3))
    t = timeit(lambda: ops.dog_u8(img)); rows.append(("dog_u8", t, px * 13 / t / 1e6, px / t / 1e3))
    d1, d2 = ops.dog_u8(img), ops.dog_u8(img2)
    t = timeit(lambda: ops.nmi_chunks(d1, d2, T * T)); rows.append(("nmi_chunks (noise)", t, px * 2 / t / 1e6, px / t / 1e3))
    sm = torch.from_numpy(np.tile(np.arange(N, dtype=np.uint8) // 8, (N, 1))).to(dev)
    t = timeit(lambda: ops.nmi_chunks(sm, sm, T * T)); rows.append(("nmi_chunks (smooth)", t, px * 2 / t / 1e6, px / t / 1e3))
    for iters in (1, 3):
        t = timeit(lambda: ops.farneback_tiles(img, img2, T, ov, 99, iters), n=3, warm=1)
        b = 44 + 60 + 28 * iters + 68 * (iters - 1)
        rows.append((f"farneback tiled it={iters} ({nt} tiles)", t, tpx * b / t / 1e6, tpx / t / 1e3))
    small = img[:1250, :1250].contiguous(); small2 = img2[:1250, :1250].contiguous()
    t = timeit(lambda: ops.farneback_tiles(small, small2, 0, 0, 99, 3), n=3, warm=1)
    rows.append(("farneback untiled 1250^2 it=3", t, 1250 * 1250 * 324 / t / 1e6, 1250 * 1250 / t / 1e3))
    print(f"N={N}  ({px/1e6:.1f} Mpx, {nt} tiles)")
    print(f"{'op':42s} {'ms':>10s} {'alg GB/s':>10s} {'Mpx/s':>12s}")
    for name, ms, gbs, mpx in rows:
        print(f"{name:42s} {ms:10.3f} {gbs:10.1f} {mpx:12.1f}")


if __name__ == "__main__":
    main()
