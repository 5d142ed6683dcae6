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


def _run(ref, mov, kw, local_pyramid=False, corrected=False):
    from microaligner_b200 import engine, parallel
    from tests import mock_ops
    saved = engine.ops, engine.torch
    engine.ops, engine.torch = mock_ops, _PoisonTorch()
    try:
        eng = engine.Engine(kw["tile_size"], kw["overlap"], kw["num_pyr_lvl"], kw["num_iterations"], kw["use_full_res_img"],
                            kw["use_dog"], comm=parallel.get(), log=lambda *a: None, corrected=corrected)
        eng.local_pyramid = local_pyramid
        flow = eng.register(torch.from_numpy(ref), torch.from_numpy(mov))
        img = eng.warp(torch.from_numpy(mov), flow)
        return flow.numpy().copy(), img.numpy().copy(), [d["better"] for d in eng.decisions]
    finally:
        engine.ops, engine.torch = saved      # the single-rank run happens inside the pytest process


def _worker(rank, world, port, case, tmp, local_pyramid, corrected=False):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
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


@pytest.mark.parametrize("world,case,corrected", [(5, "mixed levels", False), (2, "tiled levels, dog", True)])
def test_sharded_engine_corner_cases(tmp_path, world, case, corrected):
    """More ranks than tile rows (some ranks own no band at some levels); the opt-in corrected flow composition, which
    gathers the accumulated flow before composing."""
    shape, dtype, kw = CASES[case]
    ref, mov = synth_pair(shape[0], shape[1], 3, dtype, amp=2.0, period=160.0)
    want_flow, want_img, want_dec = _run(ref, mov, kw, corrected=corrected)
    mp.spawn(_worker, args=(world, _free_port(), case, str(tmp_path), False, corrected), nprocs=world, join=True)
    for r in range(world):
        got = np.load(tmp_path / f"r{r}.npz")
        assert list(got["dec"]) == want_dec
        assert np.array_equal(got["flow"], want_flow) and np.array_equal(got["img"], want_img), f"rank {r} differs"


def _worker_host_sharded(rank, world, port, case, tmp):
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
        rows, flow_rows, flow_dev = eng.register_host_sharded(ref, mov)
        wrows, img_rows = eng.warp_host_sharded(mov, flow_dev)
        np.savez(os.path.join(tmp, f"r{rank}.npz"), rows=np.array(rows), flow=flow_rows, wrows=np.array(wrows), img=img_rows,
                 need=np.array(eng.full_input_rows(ref.shape)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("case", list(CASES))
def test_sharded_host_io(tmp_path, case, world):
    """register_host_sharded / warp_host_sharded: every rank uploads only the rows of ref / mov / image it reads (the rest
    of its device buffers is garbage) and downloads only its band; the bands of all ranks tile the single-rank result."""
    shape, dtype, kw = CASES[case]
    ref, mov = synth_pair(shape[0], shape[1], 3, dtype, amp=2.0, period=160.0)
    want_flow, want_img, _ = _run(ref, mov, kw)
    mp.spawn(_worker_host_sharded, args=(world, _free_port(), case, str(tmp_path)), nprocs=world, join=True)
    flow_cover, img_cover = np.zeros(shape[0], int), np.zeros(shape[0], int)
    for r in range(world):
        got = np.load(tmp_path / f"r{r}.npz")
        (a, b), (c, d) = got["rows"], got["wrows"]
        assert np.array_equal(got["flow"], want_flow[a:b]), f"rank {r}: flow rows {a}:{b} differ"
        assert np.array_equal(got["img"], want_img[c:d]), f"rank {r}: warped rows {c}:{d} differ"
        flow_cover[a:b] += 1
        img_cover[c:d] += 1
        if case == "tiled levels, dog" and world == 3:       # every level sharded: a real share, not the whole image
            assert got["need"][1] - got["need"][0] < shape[0]
    assert (flow_cover >= 1).all() and (img_cover >= 1).all()
