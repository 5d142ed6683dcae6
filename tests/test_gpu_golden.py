"""-m gpu: the CUDA path against the golden vectors produced by the UNMODIFIED reference in the build container
(scripts/make_golden.py -> tests/golden/*.npz), through the reference-shaped classes where they exist."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def gold(name):
    return np.load(os.path.join(GOLD, name), allow_pickle=False)


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_golden_farneback_tile(cuda):
    from microaligner_b200 import ops
    g = gold("farneback_tile.npz")
    got = ops.farneback_tiles(dev(g["mov"]), dev(g["ref"]), 0, 0, int(g["win"]), int(g["iters"])).cpu().numpy()
    assert np.array_equal(got, g["flow"])
    got8 = ops.farneback_tiles(dev(g["mov8"]), dev(g["ref8"]), 0, 0, int(g["win8"]), int(g["iters8"])).cpu().numpy()
    assert np.array_equal(got8, g["flow8"])


def test_golden_tile_flow_calc_class(cuda):
    from microaligner_b200.optflow_reg.flow_calc import TileFlowCalc
    g = gold("tileflow.npz")
    t = TileFlowCalc()
    t.ref_img, t.mov_img = g["ref"], g["mov"]
    t.tile_size, t.overlap, t.num_iter, t.win_size = int(g["T"]), int(g["ov"]), int(g["iters"]), int(g["win"])
    flow = t.calc_flow()
    assert isinstance(flow, np.ndarray) and np.array_equal(flow, g["flow"])
    assert len(t.ref_img) == 0 and len(t.mov_img) == 0          # inputs are blanked like the reference (flow_calc.py:69,73)


def test_golden_warper_class(cuda):
    from microaligner_b200 import Warper
    g = gold("warper.npz")
    for img, want in ((g["img16"], g["out16"]), (g["img8"], g["out8"])):
        w = Warper()
        w.tile_size, w.overlap = int(g["T"]), int(g["ov"])
        w.image, w.flow = img, g["flow"]
        out = w.warp()
        assert out.dtype == img.dtype and np.array_equal(out, want)


def test_golden_dog_and_mi(cuda):
    from microaligner_b200 import OptFlowRegistrator
    from microaligner_b200.shared_modules.similarity_scoring import check_if_higher_similarity, mi_tiled
    g = gold("dog_nmi.npz")
    reg = OptFlowRegistrator()
    d_ref, d_mov = reg.dog(g["ref"], True), reg.dog(g["mov"], True)
    assert np.array_equal(d_ref, g["d_ref"]) and np.array_equal(d_mov, g["d_mov"])
    assert np.array_equal(reg.dog(g["blobs"], True), g["d_blobs"])
    ref = g["ref"]
    assert reg.dog(ref, False) is ref
    a, b = dev(g["d_ref"]), dev(g["d_mov"])
    assert mi_tiled(a, b, 1000) == pytest.approx(float(g["mi_whole"]), rel=1e-12)
    assert mi_tiled(a, b, int(g["chunk_T"])) == pytest.approx(float(g["mi_chunks"]), rel=1e-12)
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        assert check_if_higher_similarity(a, a, b, 1000) == [True]


def test_golden_merge(cuda):
    from microaligner_b200 import ops
    g = gold("merge.npz")
    got = ops.merge_flows_tiles(dev(g["f1"]), dev(g["f2"]), int(g["T"]), int(g["ov"])).cpu().numpy()
    assert np.array_equal(got, g["merged"])
