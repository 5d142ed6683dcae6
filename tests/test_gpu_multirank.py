"""-m gpu: the row-sharded engine with 2 and 3 ranks (gloo, all ranks on cuda:0, halos staged through the
host) must reproduce the single-rank result bit for bit -- partition math, halo exchange, DoG min/max and
NMI score reductions -- without needing a multi-GPU box."""
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

pytestmark = pytest.mark.gpu

CASES = [
    ((700, 820), np.uint16, dict(num_pyr_lvl=2, tile_size=150, overlap=20, use_full_res_img=True, use_dog=True, num_iterations=2)),
    ((640, 530), np.uint8, dict(num_pyr_lvl=2, tile_size=120, overlap=16, use_full_res_img=False, num_iterations=1)),
]
# every pyramid level tiled (>= 2 tiles per side), so that a band-local pyramid really leaves rows uncomputed
LOCAL_CASE = ((1300, 1500), np.uint16, dict(num_pyr_lvl=2, tile_size=150, overlap=20, use_full_res_img=True, use_dog=True,
                                            num_iterations=1))


def _run(ref, mov, kw):
    from microaligner_b200 import OptFlowRegistrator, Warper
    reg = OptFlowRegistrator()
    for k, v in kw.items():
        setattr(reg, k, v)
    reg.ref_img, reg.mov_img = torch.from_numpy(ref).cuda(), torch.from_numpy(mov).cuda()
    with contextlib.redirect_stdout(io.StringIO()):
        flow = reg.register()
    w = Warper()
    w.tile_size, w.overlap = kw["tile_size"], kw["overlap"]
    w.image, w.flow = torch.from_numpy(mov).cuda(), flow
    return flow.cpu().numpy(), w.warp().cpu().numpy(), [d["better"] for d in reg.decisions]


def _worker(rank, world, port, case_id, tmp, pyramid_mode=None):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    if pyramid_mode == "gathered":
        os.environ["MA_LOCAL_PYRAMID"] = "0"    # read by Engine.__init__: pyramid levels computed in slices and gathered
    elif pyramid_mode is not None:
        os.environ["MA_PYRAMID_GATHER_BELOW"] = str(pyramid_mode)   # Engine.pyramid_plan: which level is gathered
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from microaligner_b200 import parallel
    parallel.init(dist.group.WORLD)
    try:
        shape, dtype, kw = LOCAL_CASE if case_id < 0 else CASES[case_id]
        ref, mov = synth_pair(shape[0], shape[1], 7, dtype)
        flow, img, dec = _run(ref, mov, kw)
        np.savez(os.path.join(tmp, f"r{rank}.npz"), flow=flow, img=img, dec=np.array(dec))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("case_id", range(len(CASES)))
@pytest.mark.parametrize("world", [2, 3, 5])
def test_sharded_equals_single(cuda, tmp_path, case_id, world):
    shape, dtype, kw = CASES[case_id]
    ref, mov = synth_pair(shape[0], shape[1], 7, dtype)
    want_flow, want_img, want_dec = _run(ref, mov, kw)
    mp.spawn(_worker, args=(world, _free_port(), case_id, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        got = np.load(tmp_path / f"r{r}.npz")
        assert list(got["dec"]) == want_dec
        assert np.array_equal(got["flow"], want_flow), f"rank {r}: flow differs"
        assert np.array_equal(got["img"], want_img), f"rank {r}: warped image differs"


@pytest.mark.timeout(600)
@pytest.mark.parametrize("world", [2, 3])
def test_sharded_pyramid_plans_equal_single(cuda, tmp_path, world):
    """Every way of building the pyramids on several ranks (Engine.pyramid_plan; LOCAL_CASE has every level tiled, so
    band-local levels really leave rows uncomputed): the default plan (first level gathered, here), a band-local level
    above the gathered one (threshold 400: levels of 650 and 325 rows), all levels band-local (threshold 1), and
    MA_LOCAL_PYRAMID=0, which computes every level in slices and gathers it."""
    shape, dtype, kw = LOCAL_CASE
    ref, mov = synth_pair(shape[0], shape[1], 7, dtype)
    want_flow, want_img, want_dec = _run(ref, mov, kw)
    for mode in (None, 400, 1, "gathered"):
        mp.spawn(_worker, args=(world, _free_port(), -1, str(tmp_path), mode), nprocs=world, join=True)
        for r in range(world):
            got = np.load(tmp_path / f"r{r}.npz")
            assert list(got["dec"]) == want_dec
            assert np.array_equal(got["flow"], want_flow), f"rank {r}: flow differs (pyramid mode {mode})"
            assert np.array_equal(got["img"], want_img), f"rank {r}: warped image differs (pyramid mode {mode})"


def _worker_host(rank, world, port, tmp):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from microaligner_b200 import OptFlowRegistrator, Warper, parallel
    parallel.init(dist.group.WORLD)
    try:
        shape, dtype, kw = LOCAL_CASE
        ref, mov = synth_pair(shape[0], shape[1], 7, dtype)
        for rep in range(2):       # the second round reuses the node-shared result blocks and the page-locked inputs
            reg = OptFlowRegistrator()
            for k, v in kw.items():
                setattr(reg, k, v)
            reg.ref_img, reg.mov_img = ref, mov          # numpy in ...
            with contextlib.redirect_stdout(io.StringIO()):
                flow = reg.register()                    # ... full numpy flow out, on every rank
            w = Warper()
            w.tile_size, w.overlap = kw["tile_size"], kw["overlap"]
            w.image, w.flow = mov, flow
            img = w.warp()
            # a plain (un-mirrored) copy of the flow is uploaded band-wise instead
            w.image, w.flow = mov, flow.copy()
            img2 = w.warp()
            np.savez(os.path.join(tmp, f"r{rank}_{rep}.npz"), flow=flow, img=img, img2=img2)
            del flow, img, img2
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("world", [2, 3])
def test_numpy_api_on_several_ranks(cuda, tmp_path, world):
    """The drop-in numpy API under a process group: every rank uploads only its rows and downloads only its rows, and
    every rank gets the complete flow / warped image (node-shared result arrays)."""
    shape, dtype, kw = LOCAL_CASE
    ref, mov = synth_pair(shape[0], shape[1], 7, dtype)
    want_flow, want_img, _ = _run(ref, mov, kw)
    mp.spawn(_worker_host, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        for rep in range(2):
            got = np.load(tmp_path / f"r{r}_{rep}.npz")
            assert got["flow"].shape == want_flow.shape and np.array_equal(got["flow"], want_flow), f"rank {r}/{rep}: flow differs"
            assert np.array_equal(got["img"], want_img), f"rank {r}/{rep}: warped image differs"
            assert np.array_equal(got["img2"], want_img), f"rank {r}/{rep}: warped image (uploaded flow) differs"
