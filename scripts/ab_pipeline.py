"""A/B of the Farneback kernel variants (include/microaligner_b200.h, MA_FB_VARIANT_SHIFT_V / _H / _P): bit-identity of
the stitched flow against the default kernels over ragged / small-window / untiled / uint8 cases, then per-kernel
timings on a batch of 1200^2 tile windows.  Writes gpurun_out/ab_pipeline.json.
Usage: python scripts/ab_pipeline.py [--size 6000] [--variants 0,0,0 0,0,1 2,4,1 ...]  (v,h,p) [--quick]"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from benchdata import synth_pair  # noqa: E402
from microaligner_b200 import _lib, ops  # noqa: E402

CASES = {
    "u16 2300x2500 T1000 ov100 it3": (2300, 2500, np.uint16, 1000, 100, 3, False),
    "u16 625x777 untiled it3": (625, 777, np.uint16, 0, 0, 3, False),
    "u8 700x900 T300 ov40 it2": (700, 900, np.uint8, 300, 40, 2, False),
    "u16 500x460 T200 ov8 it2": (500, 460, np.uint16, 200, 8, 2, False),
    "u16 1100x1300 T500 ov194 it1": (1100, 1300, np.uint16, 500, 194, 1, False),
    "u16 1500x1100 T1000 ov100 it2 contract": (1500, 1100, np.uint16, 1000, 100, 2, True),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=6000)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--variants", nargs="*", default=["0,0,0", "0,0,1", "2,0,0", "3,0,0", "0,3,0", "0,4,0", "2,4,1", "3,4,1", "1,1,0", "0,2,0"])
    ap.add_argument("--quick", action="store_true", help="parity on the first two cases only")
    args = ap.parse_args()
    variants = [tuple(int(x) for x in v.split(",")) for v in args.variants]
    res = {"parity": {}, "timing": {}}
    t_start = time.time()
    good = set(variants)
    for ci, (name, (h, w, dt, T, ov, it, contract)) in enumerate(CASES.items()):
        if args.quick and ci >= 2:
            break
        ref, mov = synth_pair(h, w, 7, dt)
        dref, dmov = ops.to_device(ref), ops.to_device(mov)
        win = ov - (1 - ov % 2) if T > 0 else 99
        base = ops.farneback_tiles(dmov, dref, T, ov, win, it, contract_fma=contract, pipelined=False, variant=(0, 0, 0))
        for v in variants:
            if not any(v):
                continue
            key = f"{','.join(map(str, v))} | {name}"
            try:
                out = ops.farneback_tiles(dmov, dref, T, ov, win, it, contract_fma=contract, pipelined=False, variant=v)
                torch.cuda.synchronize()
                same = bool(torch.equal(base, out))
                res["parity"][key] = {"identical": same, "max_abs_diff": float((base - out).abs().max())}
            except Exception as e:  # noqa: BLE001
                same = False
                res["parity"][key] = {"error": repr(e)}
            if not same:
                good.discard(v)
            print(key, res["parity"][key], flush=True)
            if "error" in res["parity"][key]:   # a trapped kernel poisons the context: stop here
                res["aborted"] = key
                break
        if "aborted" in res:
            break
    res["identical_variants"] = sorted(",".join(map(str, v)) for v in good)
    res["parity_seconds"] = time.time() - t_start

    if "aborted" not in res:
        n = args.size
        g = torch.Generator(device="cuda").manual_seed(0)
        base = torch.rand((n // 8, n // 8), device="cuda", generator=g)
        img = torch.nn.functional.interpolate(base[None, None], size=(n, n), mode="bilinear")[0, 0]
        ref32 = (img * 60000).to(torch.int32)
        ref = ref32.to(torch.uint16)
        mov = torch.roll(ref32, (2, 3), (0, 1)).to(torch.uint16).contiguous()
        for v in variants:
            for contract in (False, True):
                ops.farneback_tiles(mov, ref, 1000, 100, 99, 3, contract_fma=contract, pipelined=False, variant=v)   # warm-up
                torch.cuda.synchronize()
                _lib.lib.ma_profile_reset()
                _lib.lib.ma_profile_enable(1)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(args.reps):
                    ops.farneback_tiles(mov, ref, 1000, 100, 99, 3, contract_fma=contract, pipelined=False, variant=v)
                e1.record()
                torch.cuda.synchronize()
                _lib.lib.ma_profile_enable(0)
                prof = {k: round(v_[0] / args.reps, 3) for k, v_ in _lib.profile_summary().items() if k.startswith("fb_")}
                key = f"{','.join(map(str, v))}{' contract_fma' if contract else ''}"
                res["timing"][key] = {"ms_per_call": round(e0.elapsed_time(e1) / args.reps, 3), **prof}
                print(key, res["timing"][key], flush=True)
    res["seconds"] = time.time() - t_start
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/ab_pipeline.json", "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
