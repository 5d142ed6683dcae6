"""CPU: the row-sharded control flow of microaligner_b200.engine.Engine on 2-3 gloo ranks, with tests/mock_ops.py (cv2 on
CPU tensors, same row-range semantics as the CUDA wrappers) patched in for the device operators and every fresh buffer
poisoned.  Whatever a rank reads without having computed or fetched it shows up as a difference from the single-rank
run -- this is how the partition / halo / exchange logic (and the opt-in band-local pyramid) is checked without a GPU."""
import contextlib
import io
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.util import synth_pair

CASES = {
    # every pyramid level tiled: all stages sharded
    "tiled levels, dog": ((520, 610), np.uint16, dict(tile_size=100, overlap=16, num_pyr_lvl=1, num_iterations=1,
                                                       use_full_res_img=True, use_dog=True)),
    # coarse level untiled (replicated), full-res tiled, uint8, no DoG, flow up-scaled at the end
    "mixed levels": ((430, 380), np.uint8, dict(tile_size=120, overlap=12, num_pyr_lvl=2, num_iterations=1,
                                                use_full_res_img=False, use_dog=False)),
}


class _PoisonTorch:
    """torch, except that empty() / empty_like() return poisoned memory (NaN / random integers)."""

    def __init__(self):
        self._rng = np.random.default_rng(1234)

    def __getattr__(self, name):
        return getattr(torch, name)

    def _poison(self, t):
        if t.dtype.is_floating_point:
            t.fill_(float("nan"))
        else:
            a = t.numpy()
            a[...] = self._rng.integers(0, np.iinfo(a.dtype).max, a.shape, dtype=a.dtype, endpoint=True)
        return t

    def empty(self, *a, **k):
        return self._poison(torch.empty(*a, **k))

    def empty_like(self, *a, **k):
        return self._poison(torch.empty_like(*a, **k))


def _run(ref, mov, kw, local_pyramid=False, corrected=False, forced=None):
    from microaligner_b200 import engine, parallel
    from tests import mock_ops
    saved = engine.ops, engine.torch
    engine.ops, engine.torch = mock_ops, _PoisonTorch()
    try:
        eng = engine.Engine(kw["tile_size"], kw["overlap"], kw["num_pyr_lvl"], kw["num_iterations"], kw["use_full_res_img"],
                            kw["use_dog"], comm=parallel.get(), log=lambda *a: None, corrected=corrected)
        eng.local_pyramid = local_pyramid
        eng.force_decisions = forced
        flow = eng.register(torch.from_numpy(ref), torch.from_numpy(mov))
        img = eng.warp(torch.from_numpy(mov), flow)
        return flow.numpy().copy(), img.numpy().copy(), [d["better"] for d in eng.decisions]
    finally:
        engine.ops, engine.torch = saved      # the single-rank run happens inside the pytest process


def _worker(rank, world, port, case, tmp, local_pyramid, corrected=False, gather_below=None):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    if gather_below is not None:
        os.environ["MA_PYRAMID_GATHER_BELOW"] = str(gather_below)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from microaligner_b200 import parallel
    parallel.init(dist.group.WORLD)
    try:
        shape, dtype, kw = CASES[case]
        ref, mov = synth_pair(shape[0], shape[1], 3, dtype, amp=2.0, period=160.0)
        flow, img, dec = _run(ref, mov, kw, local_pyramid, corrected)
        np.savez(os.path.join(tmp, f"r{rank}.npz"), flow=flow, img=img, dec=np.array(dec))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("local_pyramid", [False, True])
@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("case", list(CASES))
def test_sharded_engine_equals_single_rank(tmp_path, case, world, local_pyramid):
    shape, dtype, kw = CASES[case]
    ref, mov = synth_pair(shape[0], shape[1], 3, dtype, amp=2.0, period=160.0)
    with contextlib.redirect_stdout(io.StringIO()):
        want_flow, want_img, want_dec = _run(ref, mov, kw)
    assert np.isfinite(want_flow).all()
    mp.spawn(_worker, args=(world, _free_port(), case, str(tmp_path), local_pyramid), nprocs=world, join=True)
    for r in range(world):
        got = np.load(tmp_path / f"r{r}.npz")
        assert list(got["dec"]) == want_dec
        assert np.array_equal(got["flow"], want_flow), f"rank {r}: flow differs on {np.count_nonzero(got['flow'] != want_flow)} values"
        assert np.array_equal(got["img"], want_img), f"rank {r}: warped image differs"


@pytest.mark.parametrize("world,case,corrected,gather_below", [(5, "mixed levels", False, None), (2, "tiled levels, dog", True, None),
                                                               (3, "mixed levels", False, 150), (2, "tiled levels, dog", False, 100)])
def test_sharded_engine_corner_cases(tmp_path, world, case, corrected, gather_below):
    """More ranks than tile rows (some ranks own no band at some levels); the opt-in corrected flow composition, which
    gathers the accumulated flow before composing; and the pyramid plans the small test images do not reach by themselves:
    a band-local level above the gathered one (threshold 150: levels of 215 and 108 rows), no gathered level at all."""
    shape, dtype, kw = CASES[case]
    ref, mov = synth_pair(shape[0], shape[1], 3, dtype, amp=2.0, period=160.0)
    want_flow, want_img, want_dec = _run(ref, mov, kw, corrected=corrected)
    mp.spawn(_worker, args=(world, _free_port(), case, str(tmp_path), gather_below is not None, corrected, gather_below),
             nprocs=world, join=True)
    for r in range(world):
        got = np.load(tmp_path / f"r{r}.npz")
        assert list(got["dec"]) == want_dec
        assert np.array_equal(got["flow"], want_flow) and np.array_equal(got["img"], want_img), f"rank {r} differs"


def _worker_host(rank, world, port, case, tmp, forced=None):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from microaligner_b200 import engine, parallel
    from tests import mock_ops
    parallel.init(dist.group.WORLD)
    try:
        engine.ops, engine.torch = mock_ops, _PoisonTorch()
        shape, dtype, kw = CASES[case]
        ref, mov = synth_pair(shape[0], shape[1], 3, dtype, amp=2.0, period=160.0)
        eng = engine.Engine(kw["tile_size"], kw["overlap"], kw["num_pyr_lvl"], kw["num_iterations"], kw["use_full_res_img"],
                            kw["use_dog"], comm=parallel.get(), log=lambda *a: None)
        eng.group_tiles = 1             # one tile row per group: the speculative streaming of the last level is exercised
        eng.force_decisions = forced
        sinks, real = [], mock_ops.HostSink
        mock_ops.HostSink = lambda host: sinks.append(real(host)) or sinks[-1]
        for rep in range(2):            # the second round reuses the node-shared result blocks of the first
            del sinks[:]
            flow, flow_dev = eng.register_host(ref, mov)
            flow_pushes = list(sinks[0].log)
            img = eng.warp_host(mov, flow_dev)
            np.savez(os.path.join(tmp, f"r{rank}_{rep}.npz"), flow=flow, img=img, need=np.array(eng.full_input_rows(ref.shape)),
                     pushes=np.array(flow_pushes))
            del flow, img
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("case", list(CASES))
def test_host_io_on_several_ranks(tmp_path, case, world):
    """register_host / warp_host: every rank uploads only the rows of ref / mov / image it reads (the rest of its device
    buffers is garbage) and delivers only its own rows, yet every rank ends up with the complete single-rank result
    (node-shared result arrays)."""
    shape, dtype, kw = CASES[case]
    ref, mov = synth_pair(shape[0], shape[1], 3, dtype, amp=2.0, period=160.0)
    want_flow, want_img, _ = _run(ref, mov, kw)
    mp.spawn(_worker_host, args=(world, _free_port(), case, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        for rep in range(2):
            got = np.load(tmp_path / f"r{r}_{rep}.npz")
            assert np.array_equal(got["flow"], want_flow), f"rank {r} round {rep}: flow differs"
            assert np.array_equal(got["img"], want_img), f"rank {r} round {rep}: warped image differs"
        if case == "tiled levels, dog" and world == 3:       # every level sharded: a real share, not the whole image
            assert got["need"][1] - got["need"][0] < shape[0]
        if case == "tiled levels, dog" and world == 2:       # 3 whole tile rows per rank: some are merged and sent early
            pushes = [tuple(p) for p in got["pushes"]]
            assert len(pushes) >= 2 and len(set(pushes)) == len(pushes), pushes
            rows = sorted(pushes)
            assert all(a[1] == b[0] for a, b in zip(rows, rows[1:])), pushes       # this rank's band exactly once


@pytest.mark.parametrize("forced", [(True, False), (False, True)])
def test_host_io_on_two_ranks_with_rejected_levels(tmp_path, forced):
    """The same with a rejected level: (True, False) -- the rows streamed out speculatively during the last level are all
    delivered again from the flow that is returned instead; (False, True) -- the first level is rejected after the head
    of the second was enqueued on the assumption that it would be accepted (deferred gate) and is redone."""
    case = "tiled levels, dog"
    shape, dtype, kw = CASES[case]
    ref, mov = synth_pair(shape[0], shape[1], 3, dtype, amp=2.0, period=160.0)
    want_flow, want_img, dec = _run(ref, mov, kw, forced=forced)
    assert dec == list(forced)
    mp.spawn(_worker_host, args=(2, _free_port(), case, str(tmp_path), forced), nprocs=2, join=True)
    for r in range(2):
        got = np.load(tmp_path / f"r{r}_1.npz")
        assert np.array_equal(got["flow"], want_flow), f"rank {r}: flow differs"
        assert np.array_equal(got["img"], want_img), f"rank {r}: warped image differs"


STREAM_CASE = ((640, 530), np.uint16, dict(tile_size=100, overlap=16, num_pyr_lvl=1, num_iterations=1, use_full_res_img=True,
                                           use_dog=False))


@pytest.mark.parametrize("forced", [None, (True, True), (True, False), (False, True)])
def test_speculative_streaming_on_one_rank(forced):
    """One GPU, host result: the last level's Farneback runs in groups of tile rows, each complete tile row is merged and
    delivered right behind it (sink log), and the delivered array equals the plain device-path result -- also when the
    gate rejects the last level and everything is delivered again from the flow that is returned instead."""
    from microaligner_b200 import engine, parallel
    from tests import mock_ops
    shape, dtype, kw = STREAM_CASE
    ref, mov = synth_pair(shape[0], shape[1], 5, dtype, amp=2.0, period=160.0)
    saved = engine.ops, engine.torch
    engine.ops, engine.torch = mock_ops, _PoisonTorch()
    try:
        def make():
            e = engine.Engine(kw["tile_size"], kw["overlap"], kw["num_pyr_lvl"], kw["num_iterations"], kw["use_full_res_img"],
                              kw["use_dog"], comm=parallel.Comm(None), log=lambda *a: None)
            e.force_decisions = forced
            e.group_tiles = 6            # one tile row per group
            return e
        want = make().register(torch.from_numpy(ref), torch.from_numpy(mov)).numpy().copy()
        sinks = []
        real = mock_ops.HostSink
        mock_ops.HostSink = lambda host: sinks.append(real(host)) or sinks[-1]
        try:
            eng = make()
            flow, dev = eng.register_host(ref, mov)
        finally:
            mock_ops.HostSink = real
        assert np.array_equal(flow, want) and np.array_equal(dev.numpy(), want)
        log = sinks[0].log
        ny = -(-shape[0] // kw["tile_size"])
        speculated = [r for r in log if r != (0, shape[0])]
        assert len(speculated) >= 2 and speculated[0][0] == 0 and speculated[-1][1] == shape[0], log
        assert all(a[1] == b[0] for a, b in zip(speculated, speculated[1:])), log       # contiguous, in order
        assert ny >= 6
        if forced is not None and not forced[-1]:
            assert log[-1] == (0, shape[0])          # rejected level: everything delivered again
        else:
            assert log == speculated
    finally:
        engine.ops, engine.torch = saved


DEFER_CASE = ((520, 610), np.uint16, dict(tile_size=100, overlap=16, num_pyr_lvl=2, num_iterations=1, use_full_res_img=True,
                                          use_dog=True))


def _run_deferred(ref, mov, kw, forced, defer):
    from microaligner_b200 import engine, parallel
    from tests import mock_ops
    saved = engine.ops, engine.torch
    engine.ops, engine.torch = mock_ops, _PoisonTorch()
    try:
        lines = []
        eng = engine.Engine(kw["tile_size"], kw["overlap"], kw["num_pyr_lvl"], kw["num_iterations"], kw["use_full_res_img"],
                            kw["use_dog"], comm=parallel.get(), log=lambda *a: lines.append(" ".join(str(x) for x in a)))
        eng.force_decisions, eng.defer_gate = forced, defer
        flow = eng.register(torch.from_numpy(ref), torch.from_numpy(mov))
        return flow.numpy().copy(), [d["better"] for d in eng.decisions], lines
    finally:
        engine.ops, engine.torch = saved


def _worker_deferred(rank, world, port, tmp, forced):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from microaligner_b200 import parallel
    parallel.init(dist.group.WORLD)
    try:
        shape, dtype, kw = DEFER_CASE
        ref, mov = synth_pair(shape[0], shape[1], 3, dtype, amp=2.0, period=160.0)
        flow, dec, lines = _run_deferred(ref, mov, kw, forced, True)
        np.savez(os.path.join(tmp, f"r{rank}.npz"), flow=flow, dec=np.array(dec), lines=np.array(lines))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("forced", [None, (False, True, True), (True, False, True), (False, False, False)])
def test_deferred_gate_equals_immediate_gate(tmp_path, forced):
    """The accept / reject decision of a level is read back behind the next level's pre-warp and DoG images, which are
    enqueued assuming "accepted".  A rejected level (forced here) redoes them from the right flow: flow, decisions and
    the order of the log lines equal those of the run that reads every gate at once -- on one rank and on two."""
    shape, dtype, kw = DEFER_CASE
    ref, mov = synth_pair(shape[0], shape[1], 3, dtype, amp=2.0, period=160.0)
    want = _run_deferred(ref, mov, kw, forced, False)
    got = _run_deferred(ref, mov, kw, forced, True)
    assert len(want[1]) == 3 and got[1] == want[1] and got[2] == want[2]
    order = [ln.strip().split()[0] for ln in got[2]]
    assert order == ["Pyramid", "MI", "Better" if got[1][0] else "Worse", "Pyramid", "MI", "Better" if got[1][1] else "Worse",
                     "Pyramid", "MI", "Better" if got[1][2] else "Worse"]
    assert np.array_equal(got[0], want[0])
    if forced is not None:
        mp.spawn(_worker_deferred, args=(2, _free_port(), str(tmp_path), forced), nprocs=2, join=True)
        for r in range(2):
            res = np.load(tmp_path / f"r{r}.npz")
            assert list(res["dec"]) == want[1] and list(res["lines"]) == (want[2] if r == 0 else [])     # rank 0 reports
            assert np.array_equal(res["flow"], want[0]), f"rank {r}: flow differs"
