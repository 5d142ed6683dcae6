"""bench.py -- Mpixel/s registered (Farneback flow + warp), BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--size S]

A step = OptFlowRegistrator.register() + Warper.warp() on one synthetic uint16 pair: 4 pyramid levels + full
resolution, 3 iterations, 1000-px tiles, 100-px overlap, no DoG prefilter.  The default workload is the SAME
50000 x 50000 whole-slide pair at every GPU count (BASELINE.json configs[4], strong scaling: the curve over
--gpus 1/2/4/8 is the 1 -> 8 GPU speed-up); on one GPU the 20000 x 20000 pair of configs[1] and its DoG variant
(configs[2]) are measured as well and reported under `configs1` / `configs2`.  --size S selects another pair.

One JSON line is printed by rank 0:

  value      whole-job Mpx/s with ref / mov already resident in HBM (CUDA events, max over ranks)
  e2e        the same step through the drop-in numpy API: host arrays in, host arrays out.  Every rank uploads the
             rows of ref / mov / image it reads and downloads the rows of the flow / warped image it computed, over
             its own PCIe link, inside the timed region; the copies overlap the kernels (copy stream, speculative
             download behind the last level).  The flow handed from register() to Warper is device-mirrored.
  roofline   dominant kernel, timed live with CUDA events inside the library during the timed steps
  parity     correctness bits of THIS run: sharded == single GPU (N > 1), numpy-API result == device-path result,
             GPU == CPU port of the reference on a crop (N = 1)
  phases_ms  where a step goes (device-synchronised wall time per engine phase, one extra untimed step)
  cpu_baseline  the oracle port of the reference (same cv2 / sklearn calls, all host threads) on a bounded crop of
             the same pair (rank 0, N = 1 only)

--impl reference times that CPU path alone (rank 0), same metric / unit / config, K bounded-sample steps.
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
    "minmax": 2.0, "dog_row": 2.0, "dog_col": 4.0, "dog_quant": 5.0, "nmi_hist": 2.0, "nmi_chunk": 2.0,
}
# FP32 lane-instructions per unit of the two FP32-bound kernels: 5 planes x (1 + 3 m), m = 49
FP32_INSTR = {"fb_blur_v": 5 * (1 + 3 * 49), "fb_blur_h": 5 * (1 + 3 * 49)}
# dram__bytes_read.sum + dram__bytes_write.sum per unit from the committed ncu --set full capture
# (profiles/r02_ncu_farneback_variants_full.txt: (dram_rd + dram_wr) per tile-px of the launch window)
NCU_TRAFFIC_PER_UNIT = {"fb_blur_h": 27.85, "fb_blur_v": 39.43, "fb_polyexp": 62.58, "fb_update": 68.08}


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
def cpu_reference_step(ref, mov, workers, params=None):
    from oracle import reference_flow as rf  # the one place bench.py executes oracle/: the CPU baseline
    params = params or PARAMS
    be = rf.CvBackend(workers=workers)
    t = time.perf_counter()
    flow = rf.register(ref, mov, be=be, **params)
    img = rf.warp(mov, flow, params["tile_size"], params["overlap"], be)
    return time.perf_counter() - t, flow, img


def default_cpu_sample():
    # SURVEY 8d asks for a crop large enough to occupy the host: 64 full-resolution tiles when there are >= 32 cores
    return 8000 if (os.cpu_count() or 1) >= 32 else 4000


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import cv2
    from benchdata import synth_pair_large
    workers = os.cpu_count() or 1
    cv2.setNumThreads(1)
    s = min(args.size, args.cpu_sample)
    ref, mov = synth_pair_large(s, s, seed=0)
    for _ in range(args.warmup):
        cpu_reference_step(ref[:1200, :1200].copy(), mov[:1200, :1200].copy(), workers)
    ts = [cpu_reference_step(ref, mov, workers)[0] for _ in range(args.steps)]
    t = float(np.mean(ts))
    v = ref.size / t / 1e6
    sample = f"{s}x{s} crop of the {args.size}x{args.size} pair, same parameters"
    print(json.dumps({
        "impl": "reference", "metric": "Mpixel/s registered (Farneback flow + warp)", "value": v, "unit": "Mpx/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.size, sample),
        "cpu_baseline": {"value": v, "unit": "Mpx/s", "cores": workers, "kind": "port", "sample": sample,
                         "cv2": cv2.__version__},
        "e2e": {"value": v, "unit": "Mpx/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_config(size, sample=None, params=None):
    which = {50000: "whole-slide 50000x50000 uint16 synthetic pair (BASELINE configs[4])",
             20000: "single 20000x20000 uint16 synthetic pair (BASELINE configs[1])"}.get(size, f"{size}x{size} uint16 synthetic pair")
    c = {"workload": which + ", tiled optical flow + warp", "params": params or PARAMS,
         "l2": "inputs (2 x %.1f GB) larger than the 126 MB L2" % (size * size * 2 / 1e9)}
    if sample:
        c["sample"] = sample
    return c


# ----------------------------------------------------------------------------------- B200 arm
class Bench:
    def __init__(self, args):
        import torch
        import torch.distributed as dist
        from microaligner_b200 import parallel
        self.torch, self.dist, self.args = torch, dist, args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device (the B200 arm has no CPU fallback)")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
            parallel.init(dist.group.WORLD)
        self.comm = parallel.get()

    # -- data: one copy of the pair in host memory that every rank can read (node-shared on several ranks) ----------
    def make_pair(self, S, params):
        from benchdata import synth_pair_large
        from microaligner_b200 import ops
        from microaligner_b200.engine import Engine
        torch = self.torch
        if self.world == 1:
            ref = torch.empty((S, S), dtype=torch.uint16, pin_memory=True).numpy()
            mov = torch.empty((S, S), dtype=torch.uint16, pin_memory=True).numpy()
            synth_pair_large(S, S, seed=0, out=(ref, mov))
            return ref, mov
        # page-locked host inputs, as on one GPU: the pair lives once in node-shared memory and every rank locks the rows
        # it is going to upload -- before they are filled, which is when these hosts let shared pages be locked
        ref = self.comm.shared_host_empty((S, S), np.uint16)
        mov = self.comm.shared_host_empty((S, S), np.uint16)
        eng = Engine(params["tile_size"], params["overlap"], params["num_pyr_lvl"], params["num_iterations"],
                     params["use_full_res_img"], params["use_dog"], comm=self.comm)
        wb = eng.warp_band((S, S))
        rows = eng.full_input_rows((S, S))
        rows = (min(rows[0], max(wb[0] - eng.ov, 0)), max(rows[1], min(wb[1] + eng.ov, S))) if wb[1] > wb[0] else rows
        for a in (ref, mov):
            ops.pin_rows(a, rows)
        self.comm.barrier()
        if self.rank == 0:
            synth_pair_large(S, S, seed=0, out=(ref, mov))
        self.comm.barrier()
        return ref, mov

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def timed(self, fn, steps):
        torch = self.torch
        self.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            out = fn()
            del out      # results are not retained across steps (keeps the host buffers in steady state)
        b.record()
        self.barrier()
        ms = a.elapsed_time(b)
        if self.world > 1:
            t = torch.tensor([ms], device=self.dev)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps

    def registrator(self, params):
        from microaligner_b200 import OptFlowRegistrator, Warper
        reg, wrp = OptFlowRegistrator(), Warper()
        for k, v in params.items():
            setattr(reg, k, v)
        wrp.tile_size, wrp.overlap = params["tile_size"], params["overlap"]
        # N > 1, device tensors: flow and registered image stay sharded in HBM -- every rank ends the step holding its band
        # of both, the layout the host download (e2e) and the pipeline's page writers consume.  The replicated result
        # (every rank also receives the other ranks' bands of the image over NVLink) is timed separately: `replicated_result`
        reg.gather_flow = False
        wrp.gather_image = False
        return reg, wrp

    def measure(self, S, params, steps, warmup, profile=False, extras=False):
        """Device-resident and end-to-end timings of one workload; returns a dict."""
        torch = self.torch
        from microaligner_b200 import _lib
        ref_h, mov_h = self.make_pair(S, params)
        # replicated device inputs for the device-resident leg (uploaded from a private copy: on several ranks only part of
        # the node-shared pair is page-locked by this rank, and CUDA rejects copies that straddle the edge of such a region)
        ref_d = torch.from_numpy(ref_h if self.world == 1 else np.array(ref_h)).to(self.dev)
        mov_d = torch.from_numpy(mov_h if self.world == 1 else np.array(mov_h)).to(self.dev)
        reg, wrp = self.registrator(params)

        def step_device():
            reg.ref_img, reg.mov_img = ref_d, mov_d
            flow = reg.register()
            wrp.image, wrp.flow = mov_d, flow
            return wrp.warp()

        def step_host():      # the drop-in numpy API, on every rank
            reg.ref_img, reg.mov_img = ref_h, mov_h
            flow = reg.register()
            wrp.image, wrp.flow = mov_h, flow
            return flow, wrp.warp()

        res = {}
        with quiet():
            for _ in range(warmup):
                step_device()
            if profile:
                launches0 = _lib.lib.ma_launch_count()
                _lib.lib.ma_profile_reset()
                _lib.lib.ma_profile_enable(1)
            with ClockSampler(self.local) as clk:
                res["ms_dev"] = self.timed(step_device, steps)
            if profile:
                _lib.lib.ma_profile_enable(0)
                res["launches"] = _lib.lib.ma_launch_count() - launches0
                res["prof"] = _lib.profile_summary()
            res["clocks"] = clk.summary()
            if extras:
                # opt-in FMA-contracted window blur (not bit-identical; within the 0.01 / 0.1 px contract per Farneback call)
                reg.exact_arithmetic = False
                step_device()
                res["ms_fast"] = self.timed(step_device, steps)
                reg.exact_arithmetic = True
                res["phases"] = self.phases(step_device)
                if self.world > 1:
                    wrp.gather_image = True
                    step_device()
                    res["ms_replicated"] = self.timed(step_device, steps)
                    wrp.gather_image = False
            # end to end through the numpy API
            for _ in range(max(3, warmup)):
                step_host()
            res["ms_e2e"] = self.timed(step_host, steps)
            if extras:
                res["parity"] = self.parity(S, params, ref_h, mov_h, ref_d, mov_d, step_host)
        px = S * S
        res["value"] = px / (res["ms_dev"] * 1e-3) / 1e6
        res["e2e_value"] = px / (res["ms_e2e"] * 1e-3) / 1e6
        res["io"] = self.io_bytes(S, params)
        res["host_pair"] = (ref_h, mov_h)
        del ref_d, mov_d
        torch.cuda.empty_cache()
        return res

    def phases(self, step_device):
        """One extra, untimed step with device-synchronised wall time per engine phase: rank 0's and the max over ranks."""
        from microaligner_b200.engine import Engine
        torch = self.torch
        Engine.trace = True
        Engine.times.clear()
        step_device()
        Engine.trace = False
        names = ["pyramid", "exchange", "warp", "dog", "farneback", "farneback(untiled level)", "nmi gate", "merge+pyrup",
                 "gather flow", "gather image"]
        mine = torch.tensor([Engine.times.get(n, 0.0) * 1e3 for n in names], device=self.dev, dtype=torch.float64)
        mx = mine.clone()
        if self.world > 1:
            self.dist.all_reduce(mx, op=self.dist.ReduceOp.MAX)
        return {"rank0": {n: round(v, 2) for n, v in zip(names, mine.tolist()) if v > 0},
                "max_over_ranks": {n: round(v, 2) for n, v in zip(names, mx.tolist()) if v > 0},
                "note": "every phase is bracketed by a device synchronisation, so the sum exceeds a timed step"}

    def io_bytes(self, S, params):
        """Host <-> device bytes of one end-to-end step, summed over the ranks, from the row ranges the engine uses."""
        from microaligner_b200.engine import Engine, LevelLayout
        torch = self.torch
        eng = Engine(params["tile_size"], params["overlap"], params["num_pyr_lvl"], params["num_iterations"],
                     params["use_full_res_img"], params["use_dog"], comm=self.comm)
        r0, r1 = eng.full_input_rows((S, S))
        wb = eng.warp_band((S, S))
        up_img = (min(wb[1] + params["overlap"], S) - max(wb[0] - params["overlap"], 0)) if (self.world > 1 and wb[1] > wb[0]) else \
            (S if self.world == 1 else 0)
        fr = eng.result_rows(LevelLayout(S, S, eng.T, eng.ov, self.comm), S)
        down_img = (wb[1] - wb[0]) if self.world > 1 else S
        t = torch.tensor([float(2 * (r1 - r0) + up_img) * S * 2, float(fr[1] - fr[0]) * S * 8 + float(down_img) * S * 2],
                         device=self.dev, dtype=torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t)
        return int(t[0].item()), int(t[1].item())

    def parity(self, S, params, ref_h, mov_h, ref_d, mov_d, step_host):
        """Correctness bits of this very run (none of it is timed)."""
        torch = self.torch
        from microaligner_b200 import OptFlowRegistrator, Warper, parallel
        out = {}
        # the numpy-API result (streamed / sharded host I/O) against the device path
        reg, wrp = self.registrator(params)
        reg.gather_flow = wrp.gather_image = True
        reg.ref_img, reg.mov_img = ref_d, mov_d
        flow_d = reg.register()
        wrp.image, wrp.flow = mov_d, flow_d
        img_d = wrp.warp()
        flow_h, img_h = step_host()
        if self.rank == 0:
            ok_f = ok_i = True
            for y0 in range(0, S, 2000):       # compare on the device, in bands
                y1 = min(y0 + 2000, S)
                # np.array: a private copy -- slices of the node-shared results may straddle the rows this rank page-locked
                ok_f &= bool(torch.equal(torch.from_numpy(np.array(flow_h[y0:y1])).to(self.dev), flow_d[y0:y1]))
                ok_i &= bool(torch.equal(torch.from_numpy(np.array(img_h[y0:y1])).to(self.dev), img_d[y0:y1]))
            out["numpy_api_equals_device_path"] = {"flow": ok_f, "image": ok_i}
        del flow_h, img_h
        if self.world > 1:
            # the sharded result against the same engine on ONE GPU (rank 0 re-runs it unsharded on its device tensors)
            torch.cuda.empty_cache()
            if self.rank == 0:
                saved = parallel.get()
                parallel.init(None)
                try:
                    r1, w1 = OptFlowRegistrator(), Warper()
                    for k, v in params.items():
                        setattr(r1, k, v)
                    w1.tile_size, w1.overlap = params["tile_size"], params["overlap"]
                    r1.ref_img, r1.mov_img = ref_d, mov_d
                    f1 = r1.register()
                    ok_f = bool(torch.equal(f1, flow_d))
                    w1.image, w1.flow = mov_d, f1
                    ok_i = bool(torch.equal(w1.warp(), img_d))
                    del f1
                    out["sharded_equals_single"] = {"flow": ok_f, "image": ok_i, "decisions_equal": r1.decisions == reg.decisions}
                except torch.OutOfMemoryError as e:      # the single-GPU run needs ~120 GB at 50000^2 next to the sharded buffers
                    out["sharded_equals_single"] = {"skipped": "out of memory for the single-GPU re-run: " + str(e)[:80]}
                finally:
                    parallel._COMM = saved
            self.barrier()
        del flow_d, img_d
        return out


def kernel_table(prof, steps):
    return {name: {"ms_per_step": ms / steps, "launches_per_step": n / steps, "units_per_step": units / steps}
            for name, (ms, n, units) in prof.items()}


def roofline_of(kernels, ms_dev, clocks):
    peak, peak_src = measured_hbm_peak()
    dom = max(kernels, key=lambda k: kernels[k]["ms_per_step"]) if kernels else None
    if not dom:
        return None
    k = kernels[dom]
    bpu = ALG_BYTES.get(dom)
    avg_ms = k["ms_per_step"] / k["launches_per_step"]
    units_per_launch = k["units_per_step"] / k["launches_per_step"]
    achieved = (bpu or 0) * units_per_launch / (avg_ms * 1e-3) / 1e9
    roofline = {"kernel": dom, "bound": "hbm", "actual_bound": "fp32 pipe" if dom in FP32_INSTR else "hbm", "achieved": achieved,
                "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src, "alg_bytes_per_unit": bpu,
                "units_per_launch": units_per_launch, "avg_launch_ms": avg_ms, "share_of_step": k["ms_per_step"] / ms_dev,
                "traffic": (NCU_TRAFFIC_PER_UNIT[dom] * units_per_launch) if dom in NCU_TRAFFIC_PER_UNIT else None}
    if dom in FP32_INSTR and clocks.get("sm_mhz"):
        fp_peak = 148 * 128 * clocks["sm_mhz"] * 1e6
        fp_ach = FP32_INSTR[dom] * units_per_launch / (avg_ms * 1e-3)
        roofline["fp32_issue"] = {"note": "kernel is FP32-pipe bound, not HBM bound (DESIGN.md section 4): separately rounded FP32 "
                                          "lane-ops vs 148 SM x 128 lanes x clock", "achieved_ginstr_s": fp_ach / 1e9,
                                  "peak_ginstr_s": fp_peak / 1e9, "frac": fp_ach / fp_peak, "at_sm_mhz": clocks["sm_mhz"]}
    return roofline


def run_b200(args):
    b = Bench(args)
    S, world = args.size, b.world
    m = b.measure(S, PARAMS, args.steps, args.warmup, profile=True, extras=True)
    side = {}
    if world == 1 and not args.no_side_configs and S != 20000:
        # BASELINE configs[1] (the 20000^2 pair) and configs[2] (same pair, DoG prefilter) on one GPU
        m1 = b.measure(20000, PARAMS, args.steps, args.warmup)
        side["configs1"] = {"config": workload_config(20000), "value": m1["value"], "unit": "Mpx/s", "ms_per_step": m1["ms_dev"],
                            "e2e": {"value": m1["e2e_value"], "unit": "Mpx/s", "ms_per_step": m1["ms_e2e"],
                                    "h2d_bytes_per_step": m1["io"][0], "d2h_bytes_per_step": m1["io"][1]}}
        p2 = dict(PARAMS, use_dog=True)
        m2 = b.measure(20000, p2, args.steps, args.warmup)
        side["configs2"] = {"config": workload_config(20000, params=p2), "value": m2["value"], "unit": "Mpx/s",
                            "ms_per_step": m2["ms_dev"],
                            "e2e": {"value": m2["e2e_value"], "unit": "Mpx/s", "ms_per_step": m2["ms_e2e"]}}
    if b.rank != 0:
        if world > 1:
            b.dist.destroy_process_group()
        return
    px = S * S
    kernels = kernel_table(m["prof"], args.steps)
    line = {
        "metric": "Mpixel/s registered (Farneback flow + warp)", "value": m["value"], "unit": "Mpx/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": m["ms_dev"], "higher_is_better": True, "scaling": args.scaling,
        "parallelism": f"tile-row bands over {world} GPU(s), P2P halo exchange + scalar all-reduces (NCCL); every rank ends the "
                       "step with its band of the flow and of the registered image in HBM",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(S),
        "e2e": {"value": m["e2e_value"], "unit": "Mpx/s", "ms_per_step": m["ms_e2e"], "h2d_bytes_per_step": m["io"][0],
                "d2h_bytes_per_step": m["io"][1],
                "api": "OptFlowRegistrator.register() + Warper.warp() on host numpy arrays, on every rank: each rank uploads the "
                       "rows it reads and downloads the rows it computed over its own PCIe link; results are full arrays "
                       "(page-locked on one GPU, node-shared memory on several)"},
        "gpu_launches": int(m["launches"]), "clocks": m["clocks"], "roofline": roofline_of(kernels, m["ms_dev"], m["clocks"]),
        "parity": m["parity"], "phases_ms": m["phases"],
        "contract_fma_mode": {"value": px / (m["ms_fast"] * 1e-3) / 1e6, "unit": "Mpx/s", "ms_per_step": m["ms_fast"],
                              "note": "opt-in OptFlowRegistrator.exact_arithmetic=False; NOT the headline: flow no longer bit-identical"},
        "kernels": {k: {kk: round(vv, 4) for kk, vv in v.items()} for k, v in sorted(kernels.items(), key=lambda kv: -kv[1]["ms_per_step"])},
    }
    if "ms_replicated" in m:
        line["replicated_result"] = {"value": px / (m["ms_replicated"] * 1e-3) / 1e6, "unit": "Mpx/s", "ms_per_step": m["ms_replicated"],
                                     "note": "same step with Warper.gather_image=True (the API default): every rank also receives "
                                             "the other ranks' bands of the registered image over NVLink"}
    line.update(side)
    if world == 1 and not args.no_cpu_baseline:
        import cv2
        import torch
        from microaligner_b200 import OptFlowRegistrator, Warper
        cv2.setNumThreads(1)
        workers = os.cpu_count() or 1
        s = min(S, args.cpu_sample)
        cref, cmov = (np.ascontiguousarray(a[:s, :s]) for a in m["host_pair"])
        with quiet():
            t, cflow, cimg = cpu_reference_step(cref, cmov, workers)
            reg, wrp = OptFlowRegistrator(), Warper()
            for k, v in PARAMS.items():
                setattr(reg, k, v)
            reg.ref_img, reg.mov_img = cref, cmov
            gflow = reg.register()
            wrp.image, wrp.flow = cmov, gflow
            gimg = wrp.warp()
        epe = np.sqrt(((np.asarray(gflow) - cflow) ** 2).sum(-1))
        line["cpu_baseline"] = {"value": s * s / t / 1e6, "unit": "Mpx/s", "cores": workers, "kind": "port", "seconds": t,
                                "sample": f"{s}x{s} crop of the same pair, same parameters, one pass", "cv2": cv2.__version__}
        line["parity"]["gpu_equals_cpu_crop"] = {"flow_identical": bool(np.array_equal(gflow, cflow)),
                                                 "image_identical": bool(np.array_equal(gimg, cimg)),
                                                 "flow_mean_epe_px": float(epe.mean()), "flow_max_epe_px": float(epe.max()),
                                                 "crop": f"{s}x{s}"}
    print(json.dumps(line))
    if world > 1:
        b.dist.destroy_process_group()


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
    ap.add_argument("--size", type=int, default=50000,
                    help="image side; default: the 50000^2 whole-slide pair of BASELINE configs[4] at every GPU count "
                         "(strong scaling); 20000 = configs[1]")
    ap.add_argument("--cpu-sample", type=int, default=default_cpu_sample())
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-side-configs", action="store_true", help="one GPU: skip the 20000^2 measurements (configs[1], configs[2])")
    ap.add_argument("--use-dog", action="store_true", help="BASELINE configs[2]: the DoG prefilter enabled for the main workload")
    args = ap.parse_args()
    if args.use_dog:
        PARAMS["use_dog"] = True
    args.scaling = "strong"      # the same pair at every GPU count
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)
    sys.stdout.flush()


if __name__ == "__main__":
    main()
