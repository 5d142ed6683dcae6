"""Where does the device wait?  One register()+warp() step with Engine.timeline on: for every engine phase the host time at
which it was enqueued and the device time at which it started / ended (CUDA events, no synchronisation inside the step).
`starved` = the host entered the phase after the device had finished everything before it (the queue ran dry).
  python scripts/timeline.py [size]            or under torchrun for several ranks (prints rank 0 and the last rank)"""
import contextlib
import io
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from benchdata import synth_pair_large  # noqa: E402
from microaligner_b200 import OptFlowRegistrator, Warper, engine, parallel  # noqa: E402


def step(ref, mov):
    reg = OptFlowRegistrator()
    reg.num_pyr_lvl, reg.num_iterations, reg.tile_size, reg.overlap, reg.use_full_res_img, reg.use_dog = 4, 3, 1000, 100, True, False
    reg.ref_img, reg.mov_img = ref, mov
    reg.gather_flow = False          # as bench.py: the flow stays sharded between register() and warp()
    with contextlib.redirect_stdout(io.StringIO()):
        flow = reg.register()
    w = Warper()
    w.tile_size, w.overlap, w.image, w.flow = 1000, 100, mov, flow
    return w.warp()


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        parallel.init(dist.group.WORLD)
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    ref_h, mov_h = synth_pair_large(S, S, seed=1)
    ref, mov = torch.from_numpy(ref_h).to(dev), torch.from_numpy(mov_h).to(dev)
    for _ in range(2):
        step(ref, mov)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    base_ev = torch.cuda.Event(enable_timing=True)
    base_ev.record()
    torch.cuda.synchronize()
    base_t = time.perf_counter()
    engine.Engine.timeline = []
    step(ref, mov)
    torch.cuda.synchronize()
    total = (time.perf_counter() - base_t) * 1e3
    tl, engine.Engine.timeline = engine.Engine.timeline, None
    rows, prev_end, starved, busy = [], 0.0, 0.0, {}
    for name, t0, e0, t1, e1 in tl:
        g0, g1 = base_ev.elapsed_time(e0), base_ev.elapsed_time(e1)
        c0, c1 = (t0 - base_t) * 1e3, (t1 - base_t) * 1e3
        dry = max(0.0, c0 - max(prev_end, 0.0)) if c0 > prev_end else 0.0
        starved += dry
        busy[name] = busy.get(name, 0.0) + (g1 - g0)
        rows.append((name, c0, c1, g0, g1, dry))
        prev_end = max(prev_end, g1)
    rank = dist.get_rank() if world > 1 else 0
    if rank in (0, world // 2, world - 1):
        out = [f"rank {rank}/{world} size {S} defer_gate={os.environ.get('MA_DEFER_GATE', '1')}: step {total:.2f} ms, "
               f"device queue dry for {starved:.2f} ms when a phase was enqueued"]
        out.append(f"{'phase':26s} {'host in':>9s} {'host out':>9s} {'dev start':>9s} {'dev end':>9s} {'dry':>6s}")
        for name, c0, c1, g0, g1, dry in rows:
            out.append(f"{name:26s} {c0:9.2f} {c1:9.2f} {g0:9.2f} {g1:9.2f} {dry:6.2f}")
        out.append("device time between a phase's events, summed: " + ", ".join(f"{k} {v:.1f}" for k, v in busy.items()))
        print("\n".join(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
