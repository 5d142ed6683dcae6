"""Timing of the NMI gate kernel (ma_nmi_chunks) on DoG-like (smooth) and noise-like u8 images.
Usage: python scripts/time_nmi.py [--size 12000]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from microaligner_b200 import _lib, ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=12000)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    n = args.size
    g = torch.Generator(device="cuda").manual_seed(0)
    base = torch.rand((n // 16, n // 16), device="cuda", generator=g)
    smooth = torch.nn.functional.interpolate(base[None, None], size=(n, n), mode="bicubic")[0, 0].clamp(0, 1)
    images = {
        "smooth": ((smooth * 255).to(torch.uint8), (torch.roll(smooth, (1, 2), (0, 1)) * 255).to(torch.uint8)),
        "noise": (torch.randint(0, 256, (n, n), device="cuda", dtype=torch.uint8, generator=g),
                  torch.randint(0, 256, (n, n), device="cuda", dtype=torch.uint8, generator=g)),
        "odd chunk": ((smooth[:3001, :2999] * 255).to(torch.uint8).contiguous(),
                      (smooth[1:3002, :2999] * 255).to(torch.uint8).contiguous()),
    }
    res = {}
    for name, (a, b) in images.items():
        chunk = 1000 * 1000 if name != "odd chunk" else 150 * 150 + 7
        ref = None
        for variant in (0,):
            scores = ops.nmi_chunks(a, b, chunk)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.reps):
                ops.nmi_chunks(a, b, chunk)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.reps
            ref = scores if ref is None else ref
            res[f"{name} v{variant}"] = {"ms": round(ms, 3), "gpx_s": round(a.numel() / ms / 1e6, 1),
                                         "identical_scores": bool(torch.equal(ref, scores))}
            print(name, res[f"{name} v{variant}"], flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/time_nmi.json", "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
