"""BASELINE configs[3]: a synthetic multiplexed stack -- C cycles x 2 channels x Z z-planes of S x S uint16 -- through the YAML
pipeline entry (python -m microaligner_b200 config.yaml: z-MIP, per-cycle register, every page warped and written to one
BigTIFF stack), on the GPUs of this torchrun launch.

    torchrun --nproc-per-node 8 scripts/bench_pipeline.py [--size 20000] [--cycles 8] [--zplanes 3] [--dir /dev/shm/ma_c4]

The dataset is written as one OME-TIFF per cycle (rank 0), then the pipeline runs on all ranks; rank 0 prints one JSON
line: wall seconds, registered Mpx/s (cycles - 1 registrations), pages/s, per-stage seconds of rank 0."""
import argparse
import contextlib
import io
import json
import os
import shutil
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from benchdata import synth_pair_large  # noqa: E402
from microaligner_b200 import parallel, tiffio  # noqa: E402
from microaligner_b200 import __main__ as cli  # noqa: E402
from tests.pipeline_data import ome_xml  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=20000)
    ap.add_argument("--cycles", type=int, default=8)
    ap.add_argument("--zplanes", type=int, default=3)
    ap.add_argument("--dir", default="/dev/shm/ma_c4")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        parallel.init(dist.group.WORLD)
    S, C, Z = args.size, args.cycles, args.zplanes
    names = ("DAPI", "CD3")
    cfg_path = os.path.join(args.dir, "config.yaml")
    t_gen = time.perf_counter()
    if rank == 0:
        shutil.rmtree(args.dir, ignore_errors=True)
        os.makedirs(args.dir)
        ref, mov = synth_pair_large(S, S, seed=0)
        paths = {}
        for cyc in range(1, C + 1):
            src = ref if cyc == 1 else (mov if cyc % 2 == 0 else np.roll(mov, 1, axis=1))
            p = os.path.join(args.dir, f"cycle{cyc}.ome.tif")
            mm = tiffio.memmap(p, (len(names), Z, S, S), np.uint16, description=ome_xml(names, Z, h=S, w=S))
            for ci in range(len(names)):
                for z in range(Z):
                    mm[ci, z] = (src >> z) if ci == 0 else ((65535 - src) >> (z + 1))
            mm.flush()
            del mm
            paths[f"Cycle {cyc}"] = p
        import yaml
        cfg = {"Input": {"InputImagePaths": paths, "ReferenceCycle": 1, "ReferenceChannel": "DAPI"},
               "Output": {"OutputDir": os.path.join(args.dir, "out"), "OutputPrefix": "c4_", "SaveOutputToCycleStack": True},
               "RegistrationParameters": {"OptFlowReg": dict(NumberPyramidLevels=4, NumberIterationsPerLevel=3, TileSize=1000,
                                                             Overlap=100, NumberOfWorkers=0, UseFullResImage=True, UseDOG=False)}}
        with open(cfg_path, "w") as f:
            yaml.safe_dump(cfg, f, sort_keys=False)
    parallel.get().barrier()
    t_gen = time.perf_counter() - t_gen
    os.environ["MA_PIPELINE_TIMING"] = "1"
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        decisions = cli.main([cfg_path])
    parallel.get().barrier()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    if rank == 0:
        with tiffio.TiffFile(os.path.join(args.dir, "out", "c4_optflow_reg_result_stack.tif")) as tf:
            n_pages = len(tf.pages)
            first_ok = bool(np.array_equal(tf.pages[0].asarray()[:512], tiffio.TiffFile(os.path.join(args.dir, "cycle1.ome.tif")).pages[0].asarray()[:512]))
        better = [[d["better"] for d in decisions[c]] for c in sorted(decisions)]
        print(json.dumps({
            "workload": f"{C} cycles x {len(names)} channels x {Z} z-planes of {S}x{S} uint16 through python -m microaligner_b200 config.yaml "
                        "(BASELINE configs[3])", "n_gpus": world, "wall_s": wall, "dataset_generation_s": t_gen,
            "registered_mpx_per_s": (C - 1) * S * S / wall / 1e6, "pages_per_s": n_pages / wall, "pages": n_pages,
            "gb_read_and_written": 2 * n_pages * S * S * 2 / 1e9, "stage_seconds_rank0": {k: round(v, 3) for k, v in cli.TIMES.items()},
            "first_cycle_passes_through": first_ok, "all_levels_better": all(all(b) for b in better), "stdout_lines": len(buf.getvalue().splitlines())}))
        shutil.rmtree(args.dir, ignore_errors=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
