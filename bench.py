"""bench.py -- Mpixel/s registered (Farneback flow + warp), BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--size S]

A step = OptFlowRegistrator.register() + Warper.warp() on one synthetic uint16 pair (BASELINE.json
configs[1]: 20000 x 20000, 4 pyramid levels + full resolution, 3 iterations, 1000-px tiles, 100-px
overlap, no DoG prefilter).  One JSON line is printed by rank 0:

  value      whole-job Mpx/s with ref/mov already resident in HBM (CUDA events, max over ranks)
  e2e        the same step through the drop-in numpy API: page-locked host arrays in, host arrays out
             (H2D of ref+mov, D2H of the flow, H2D of the image for Warper -- the read-only flow array returned by
             register() is device-mirrored, so handing it to Warper costs no upload -- D2H of the warped image)
  roofline   dominant kernel, timed live with CUDA events inside the library during the timed steps
  cpu_baseline  the oracle port of the reference (same cv2 / sklearn calls, all host threads) on a
             bounded crop of the same pair (rank 0, N=1 only)

--impl reference times that CPU path alone (rank 0), same metric/unit/config, K bounded-sample steps.
"""
import argparse
import contextlib
import io
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PARAMS = dict(num_pyr_lvl=4, num_iterations=3, tile_size=1000, overlap=100, use_full_res_img=True, use_dog=False)

# algorithmic HBM bytes per unit (unit = tile-pixel for fb_*, pixel otherwise; u16 input)
ALG_BYTES = {
    "fb_polyexp": 64.0,   # 2 x (2 in + 20 out) + fused first UpdateMatrices (M 20 out; its R0/R1 inputs stay on chip)
    "fb_blur_v": 20.0,    # M in (the transposed V it writes is an intermediate of the blur+solve stage)
    "fb_blur_h": 8.0,     # flow out
    "fb_update": 68.0,    # flow 8 + R0 20 + R1 20 + M 20
    "warp_tiles": 12.0, "tile_max": 16.0, "merge_tiles": 24.0, "pyrdown": 2.5, "pyrup_flow": 10.0,
    "minmax": 2.0, "dog_row": 2.0, "dog_col": 4.0, "dog_quant": 5.0, "nmi_hist": 2.0,
}
# FP32 lane-instructions per unit of the two FP32-bound kernels: 5 planes x (1 + 3 m), m = 49
FP32_INSTR = {"fb_blur_v": 5 * (1 + 3 * 49), "fb_blur_h": 5 * (1 + 3 * 49)}
# dram__bytes_read.sum + dram__bytes_write.sum per unit from the committed ncu --set full capture (profiles/)
NCU_TRAFFIC_PER_UNIT = {  # profiles/r01_ncu_farneback_full.txt: (dram_rd + dram_wr) / 51.84 M tile-px per launch
    "fb_blur_h": 27.85, "fb_blur_v": 39.43, "fb_polyexp": 62.58, "fb_update": 68.08,
}


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    FIELDS = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        busy = [v for v in sm if v > 0]
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def quiet():
    return contextlib.redirect_stdout(io.StringIO())


# ----------------------------------------------------------------------------------- CPU reference arm
def cpu_reference_step(ref, mov, workers):
    from oracle import reference_flow as rf  # the one place bench.py executes oracle/: the CPU baseline
    be = rf.CvBackend(workers=workers)
    t = time.perf_counter()
    flow = rf.register(ref, mov, be=be, **PARAMS)
    rf.warp(mov, flow, PARAMS["tile_size"], PARAMS["overlap"], be)
    return time.perf_counter() - t


def cpu_sample(size, sample):
    from benchdata import synth_pair_large
    s = min(size, sample)
    return synth_pair_large(s, s, seed=0)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import cv2
    workers = os.cpu_count() or 1
    cv2.setNumThreads(1)
    ref, mov = cpu_sample(args.size, args.cpu_sample)
    px = ref.size
    for _ in range(args.warmup):
        cpu_reference_step(ref[:1200, :1200].copy(), mov[:1200, :1200].copy(), workers)
    ts = [cpu_reference_step(ref, mov, workers) for _ in range(args.steps)]
    t = float(np.mean(ts))
    v = px / t / 1e6
    sample = f"{ref.shape[0]}x{ref.shape[1]} crop of the {args.size}x{args.size} pair, same parameters"
    print(json.dumps({
        "impl": "reference", "metric": "Mpixel/s registered (Farneback flow + warp)", "value": v, "unit": "Mpx/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, sample),
        "cpu_baseline": {"value": v, "unit": "Mpx/s", "cores": workers, "kind": "port", "sample": sample,
                         "cv2": cv2.__version__},
        "e2e": {"value": v, "unit": "Mpx/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_config(args, sample=None):
    c = {"workload": f"single {args.size}x{args.size} uint16 synthetic pair, tiled optical flow + warp",
         "params": PARAMS, "l2": "inputs (2 x %.1f GB) larger than the 126 MB L2" % (args.size * args.size * 2 / 1e9)}
    if sample:
        c["sample"] = sample
    return c


# ----------------------------------------------------------------------------------- B200 arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    from benchdata import synth_pair_large
    from microaligner_b200 import OptFlowRegistrator, Warper, _lib, ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the B200 arm has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        from microaligner_b200 import parallel
        parallel.init(dist.group.WORLD)

    S = args.size
    comm = parallel.get() if world > 1 else None
    if rank == 0:
        ref_h = torch.empty((S, S), dtype=torch.uint16, pin_memory=True).numpy()
        mov_h = torch.empty((S, S), dtype=torch.uint16, pin_memory=True).numpy()
        synth_pair_large(S, S, seed=0, out=(ref_h, mov_h))
        ref_d, mov_d = torch.from_numpy(ref_h).to(dev), torch.from_numpy(mov_h).to(dev)
    else:
        ref_h = mov_h = None
        ref_d = torch.empty((S, S), dtype=torch.uint16, device=dev)
        mov_d = torch.empty((S, S), dtype=torch.uint16, device=dev)
    if world > 1:  # inputs are replicated on every GPU; only computed data crosses NVLink afterwards
        comm.broadcast(ref_d, 0)
        comm.broadcast(mov_d, 0)

    ref_sh = mov_sh = None
    if world > 1 and args.sharded_io:
        # the pair lives once in shared host memory; every rank page-locks just the rows it will upload
        from microaligner_b200.engine import Engine
        port = os.environ.get("MASTER_PORT", "0")
        import shutil
        import tempfile
        shm = ["/dev/shm" if os.path.isdir("/dev/shm") and shutil.disk_usage("/dev/shm").free > 5 * S * S else tempfile.gettempdir()]
        dist.broadcast_object_list(shm, src=0)      # rank 0 decides where the pair lives
        paths = [os.path.join(shm[0], f"ma_bench_{port}_{n}.u16") for n in ("ref", "mov")]
        if rank == 0:
            for pth, arr in zip(paths, (ref_h, mov_h)):
                mm = np.memmap(pth, dtype=np.uint16, mode="w+", shape=(S, S))
                mm[...] = arr
                mm.flush()
        dist.barrier()
        ref_sh, mov_sh = (np.memmap(pth, dtype=np.uint16, mode="r+", shape=(S, S)) for pth in paths)
        probe = Engine(PARAMS["tile_size"], PARAMS["overlap"], PARAMS["num_pyr_lvl"], PARAMS["num_iterations"],
                       PARAMS["use_full_res_img"], PARAMS["use_dog"], comm=comm)
        r0, r1 = probe.full_input_rows((S, S))
        wb = probe.warp_band((S, S))
        up_rows = 2 * (r1 - r0) + (min(wb[1] + PARAMS["overlap"], S) - max(wb[0] - PARAMS["overlap"], 0) if wb[1] > wb[0] else 0)
        t_up = torch.tensor([float(up_rows) * S * 2], device=dev, dtype=torch.float64)
        dist.all_reduce(t_up)
        sharded_h2d = int(t_up.item())
        for arr in (ref_sh, mov_sh):
            rows = torch.from_numpy(arr)[r0:r1]
            err = torch.cuda.cudart().cudaHostRegister(rows.data_ptr(), rows.numel() * rows.element_size(), 0)
            if int(err) != 0:
                sys.stderr.write(f"[rank {rank}] cudaHostRegister failed ({err}); uploads will be staged\n")
        dist.barrier()
        if rank == 0:
            for pth in paths:
                os.unlink(pth)       # the mappings keep the memory alive

    reg, wrp = OptFlowRegistrator(), Warper()
    for k, v in PARAMS.items():
        setattr(reg, k, v)
    wrp.tile_size, wrp.overlap = PARAMS["tile_size"], PARAMS["overlap"]
    # N > 1: the flow stays sharded on the GPUs that computed it (each band is what that rank's warp reads);
    # the final NVLink gather is that of the warped image
    reg.gather_flow = False

    def step_device():
        reg.ref_img, reg.mov_img = ref_d, mov_d
        flow = reg.register()
        wrp.image, wrp.flow = mov_d, flow
        return wrp.warp()

    def step_host_sharded():
        # opt-in (--sharded-io): every rank moves only its own share over its own PCIe link -- uploads the rows of
        # ref / mov it reads, downloads its band of the flow and of the warped image
        reg.ref_img, reg.mov_img = ref_sh, mov_sh
        rows, flow = reg.register_sharded()
        wrp.image, wrp.flow = mov_sh, reg.device_flow
        out = wrp.warp_sharded()
        reg.device_flow = None
        return flow, out

    def step_host():
        if world > 1 and args.sharded_io:
            return step_host_sharded()
        if world == 1:  # the drop-in numpy API
            reg.ref_img, reg.mov_img = ref_h, mov_h
            flow = reg.register()
            wrp.image, wrp.flow = mov_h, flow
            return wrp.warp()
        # N ranks: rank 0 owns the host arrays; upload + NVLink broadcast, sharded step, rank 0 downloads
        r = ops.to_device(ref_h, dev) if rank == 0 else torch.empty_like(ref_d)
        m = ops.to_device(mov_h, dev) if rank == 0 else torch.empty_like(mov_d)
        comm.broadcast(r, 0)
        comm.broadcast(m, 0)
        reg.ref_img, reg.mov_img = r, m
        reg.gather_flow = True          # the host wants the whole flow on rank 0
        flow = reg.register()
        reg.gather_flow = False
        wrp.image, wrp.flow = m, flow
        out = wrp.warp()
        if rank == 0:
            return ops.to_host(flow), ops.to_host(out)
        return None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            out = fn()
            del out      # results are not retained across steps (keeps the pinned host buffers in steady state)
        b.record()
        barrier()
        ms = a.elapsed_time(b)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps

    with quiet():
        for _ in range(args.warmup):
            step_device()
        launches0 = _lib.lib.ma_launch_count()
        _lib.lib.ma_profile_reset()
        _lib.lib.ma_profile_enable(1)
        with ClockSampler(local) as clk:
            ms_dev = timed(step_device, args.steps)
        _lib.lib.ma_profile_enable(0)
        launches = _lib.lib.ma_launch_count() - launches0
        prof = _lib.profile_summary()
        # opt-in FMA-contracted window blur (not bit-identical; within the 0.01 / 0.1 px contract per Farneback call)
        reg.exact_arithmetic = False
        step_device()
        ms_fast = timed(step_device, args.steps)
        reg.exact_arithmetic = True
        if args.trace:
            from microaligner_b200.engine import Engine
            Engine.trace = True
            Engine.times.clear()
            step_device()
            Engine.trace = False
            sys.stderr.write(f"[trace rank {rank}] " + json.dumps({k: round(v * 1e3, 2) for k, v in Engine.times.items()}) + "\n")
        # end-to-end through the numpy API (page-locked host arrays)
        for _ in range(max(3, args.warmup)):
            step_host()
        ms_e2e = timed(step_host, args.steps)
    px = S * S
    value = px / (ms_dev * 1e-3) / 1e6
    e2e_val = px / (ms_e2e * 1e-3) / 1e6
    img_b, flow_b = px * 2, px * 8
    if world == 1:
        h2d = 2 * img_b + img_b            # register(ref, mov) + Warper(image); the flow array is device-mirrored
    elif args.sharded_io:
        h2d = sharded_h2d                  # every rank: its rows of ref and mov (+ pyramid / overlap halos) and of the image to warp
    else:
        h2d = 2 * img_b                    # rank 0 uploads ref, mov; the flow stays on the devices
    d2h = flow_b + img_b                   # flow returned by register(), warped image

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    clocks = clk.summary()
    peak, peak_src = measured_hbm_peak()
    N = PARAMS["num_iterations"]
    kernels = {}
    for name, (ms, n, units) in prof.items():
        kernels[name] = {"ms_per_step": ms / args.steps, "launches_per_step": n / args.steps, "units_per_step": units / args.steps}
    dom = max(kernels, key=lambda k: kernels[k]["ms_per_step"]) if kernels else None
    roofline = None
    if dom:
        k = kernels[dom]
        bpu = ALG_BYTES.get(dom)
        avg_ms = k["ms_per_step"] / k["launches_per_step"]
        units_per_launch = k["units_per_step"] / k["launches_per_step"]
        achieved = (bpu or 0) * units_per_launch / (avg_ms * 1e-3) / 1e9
        roofline = {"kernel": dom, "bound": "hbm", "actual_bound": "fp32 pipe" if dom in FP32_INSTR else "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "peak_source": peak_src, "alg_bytes_per_unit": bpu, "units_per_launch": units_per_launch,
                    "avg_launch_ms": avg_ms, "share_of_step": k["ms_per_step"] / ms_dev,
                    "traffic": (NCU_TRAFFIC_PER_UNIT[dom] * units_per_launch) if dom in NCU_TRAFFIC_PER_UNIT else None}
        if dom in FP32_INSTR and clocks.get("sm_mhz"):
            fp_peak = 148 * 128 * clocks["sm_mhz"] * 1e6
            fp_ach = FP32_INSTR[dom] * units_per_launch / (avg_ms * 1e-3)
            roofline["fp32_issue"] = {"note": "kernel is FP32-pipe bound, not HBM bound (DESIGN.md section 4): separately rounded FP32 lane-ops vs 148 SM x 128 lanes x clock", "achieved_ginstr_s": fp_ach / 1e9,
                                      "peak_ginstr_s": fp_peak / 1e9, "frac": fp_ach / fp_peak, "at_sm_mhz": clocks["sm_mhz"]}
    line = {
        "metric": "Mpixel/s registered (Farneback flow + warp)", "value": value, "unit": "Mpx/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev, "higher_is_better": True, "scaling": args.scaling,
        "parallelism": f"tile-row bands over {world} GPU(s), P2P halo exchange + scalar all-reduces (NCCL)",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args),
        "e2e": {"value": e2e_val, "unit": "Mpx/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "api": "OptFlowRegistrator.register() + Warper.warp() on page-locked numpy arrays" if world == 1 else
                       ("every rank: register_sharded() + warp_sharded() -- its own rows of the page-locked host pair up, its band of "
                        "flow and image down, over its own PCIe link" if args.sharded_io else
                        "rank 0: page-locked numpy in -> H2D -> NVLink broadcast -> sharded register()+warp() -> D2H of flow and image")},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
        "contract_fma_mode": {"value": px / (ms_fast * 1e-3) / 1e6, "unit": "Mpx/s", "ms_per_step": ms_fast,
                              "note": "opt-in OptFlowRegistrator.exact_arithmetic=False; NOT the headline: flow no longer bit-identical"},
        "kernels": {k: {kk: round(vv, 4) for kk, vv in v.items()} for k, v in sorted(kernels.items(), key=lambda kv: -kv[1]["ms_per_step"])},
    }
    if world == 1 and not args.no_cpu_baseline:
        import cv2
        cv2.setNumThreads(1)
        workers = os.cpu_count() or 1
        s = min(S, args.cpu_sample)
        cref, cmov = np.ascontiguousarray(ref_h[:s, :s]), np.ascontiguousarray(mov_h[:s, :s])
        with quiet():
            t = cpu_reference_step(cref, cmov, workers)
        line["cpu_baseline"] = {"value": s * s / t / 1e6, "unit": "Mpx/s", "cores": workers, "kind": "port", "seconds": t,
                                "sample": f"{s}x{s} crop of the same pair, same parameters, one pass", "cv2": cv2.__version__}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    # stdout carries exactly ONE JSON line: everything else any library prints at fd level (NCCL's version
    # banner, torch warnings) is routed to stderr for the duration of the run
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w")
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", type=int, default=None,
                    help="image side; default: 20000 at 1 GPU (BASELINE configs[1]) growing with the GPU count up to the "
                         "50000 whole-slide pair of configs[4] at 8 GPUs (weak scaling, image area ~ N)")
    ap.add_argument("--strong", action="store_true", help="keep the 20000^2 pair at every GPU count (strong scaling)")
    ap.add_argument("--cpu-sample", type=int, default=4000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--use-dog", action="store_true", help="BASELINE configs[2]: same pair with the DoG prefilter enabled")
    ap.add_argument("--sharded-io", action="store_true",
                    help="N > 1, e2e leg only: every rank uploads its own rows and downloads its own band (register_sharded / "
                         "warp_sharded) instead of rank 0 moving everything")
    ap.add_argument("--trace", action="store_true", help="extra untimed step with per-phase synchronised wall times (stderr)")
    args = ap.parse_args()
    if args.use_dog:
        PARAMS["use_dog"] = True
    args.scaling = "weak"
    if args.size is None:
        n = max(1, int(os.environ.get("WORLD_SIZE", args.gpus)))
        args.size = 20000 if (args.strong or n == 1) else min(50000, int(round(20000 * n ** 0.5 / 1000.0)) * 1000)
        if args.strong:
            args.scaling = "strong"
    elif int(os.environ.get("WORLD_SIZE", args.gpus)) > 1:
        args.scaling = "strong"  # explicit size: the same image at every GPU count
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)
    sys.stdout.flush()


if __name__ == "__main__":
    main()
