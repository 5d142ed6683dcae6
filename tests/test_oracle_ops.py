"""CPU: pin the oracle's numpy restatements against the live cv2 / sklearn of this image (bit-exact)
and against the golden vectors produced by the unmodified reference (scripts/make_golden.py)."""
import os

import numpy as np
import pytest

from oracle import cv_ops, farneback_np
from oracle import reference_flow as rf
from tests.util import blobs, random_flow, synth_pair

cv2 = pytest.importorskip("cv2")
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def gold(name):
    return np.load(os.path.join(GOLD, name), allow_pickle=False)


# ---------------------------------------------------------------- live cv2
@pytest.mark.parametrize("case", [((300, 260), np.uint16, 99, 3), ((257, 301), np.uint8, 99, 1), ((200, 200), np.uint16, 31, 2)])
def test_farneback_vs_cv2(case):
    shape, dtype, win, iters = case
    ref, mov = synth_pair(shape[0], shape[1], 0, dtype)
    want = cv2.calcOpticalFlowFarneback(mov, ref, None, 0.5, 0, win, iters, 1, 1.7, cv2.OPTFLOW_FARNEBACK_GAUSSIAN)
    got = farneback_np.farneback(mov, ref, win, iters)
    assert np.array_equal(got, want)


def test_farneback_constants():
    g, xg, xxg, ig11, ig03, ig33, ig55 = farneback_np.poly_gaussian()
    assert (round(ig11, 8), round(ig03, 8), round(ig33, 8), round(ig55, 8)) == (1.59443929, -2.68225768, 4.27669697, 2.54223662)
    assert g[0] == g[2] and xg[1] == 0
    assert abs(float(farneback_np.blur_taps(99)[:1].sum() + 2 * farneback_np.blur_taps(99)[1:].sum()) - 1) < 1e-6


@pytest.mark.parametrize("dtype", [np.uint8, np.uint16, np.float32])
def test_remap_vs_cv2(dtype):
    rng = np.random.default_rng(1)
    h, w = 211, 177
    src = (rng.standard_normal((h, w, 2)) * 5).astype(np.float32) if dtype == np.float32 else \
        rng.integers(0, np.iinfo(dtype).max + 1, (h, w)).astype(dtype)
    y, x = np.mgrid[0:h, 0:w].astype(np.float32)
    m = np.stack([x + rng.standard_normal((h, w)).astype(np.float32) * 30, y + rng.standard_normal((h, w)).astype(np.float32) * 30], -1)
    m[0, 0] = [-1e6, 5]; m[0, 1] = [1e7, 1e7]; m[1, 1] = [-0.5, -0.5]; m[2, 2] = [w - 1, h - 1]; m[4, 4] = [0.015625, 0.046875]
    assert np.array_equal(cv2.remap(src, m, None, cv2.INTER_LINEAR), cv_ops.remap_linear(src, m))


@pytest.mark.parametrize("dtype", [np.uint8, np.uint16])
@pytest.mark.parametrize("shape", [(200, 300), (201, 303), (257, 128)])
def test_pyrdown_vs_cv2(dtype, shape):
    rng = np.random.default_rng(2)
    src = rng.integers(0, np.iinfo(dtype).max + 1, shape).astype(dtype)
    assert np.array_equal(cv2.pyrDown(src), cv_ops.pyr_down(src))


@pytest.mark.parametrize("case", [(100, 130, 200, 260), (100, 130, 199, 259), (101, 77, 201, 154), (64, 64, 128, 127), (2, 2, 3, 3)])
def test_pyrup_vs_cv2(case):
    h, w, dh, dw = case
    f = random_flow(h, w, 3)
    for sc in (1.0, 2.0, 4.0):
        assert np.array_equal(cv2.pyrUp(f * np.float32(sc), dstsize=(dw, dh)), cv_ops.pyr_up_f32c2(f, (dh, dw), sc))


@pytest.mark.parametrize("case", [((300, 260), np.uint16), ((217, 333), np.uint8), ((128, 256), np.uint16)])
def test_dog_vs_cv2(case):
    shape, dtype = case
    be = rf.CvBackend()
    ref, mov = synth_pair(shape[0], shape[1], 4, dtype)
    for img in (ref, mov, blobs(shape[0], shape[1], 2, dtype)):
        assert np.array_equal(be.dog(img), cv_ops.dog(img))
        f = cv2.normalize(img, None, 0, 1, cv2.NORM_MINMAX, cv2.CV_32F)
        for s in (5, 9):
            assert np.array_equal(cv2.GaussianBlur(f, (41, 41), sigmaX=s, sigmaY=s), cv_ops.sep_blur_41(f, cv_ops.gaussian_kernel_41(s)))
    assert np.array_equal(cv2.getGaussianKernel(41, 5, cv2.CV_32F).ravel(), cv_ops.gaussian_kernel_41(5))


def test_nmi_vs_sklearn():
    sk = pytest.importorskip("sklearn.metrics")
    ref, mov = synth_pair(300, 400, 5, np.uint8)
    a, b = cv_ops.dog(ref), cv_ops.dog(mov)
    assert cv_ops.nmi(a, b) == pytest.approx(sk.normalized_mutual_info_score(a.ravel(), b.ravel()), rel=1e-13)
    z = np.zeros_like(a)
    assert cv_ops.nmi(z, z) == sk.normalized_mutual_info_score(z.ravel(), z.ravel()) == 1.0
    assert cv_ops.nmi(z, b) == sk.normalized_mutual_info_score(z.ravel(), b.ravel()) == 0.0


# ---------------------------------------------------------------- golden vectors from the reference
def test_golden_farneback():
    g = gold("farneback_tile.npz")
    assert np.array_equal(farneback_np.farneback(g["mov"], g["ref"], int(g["win"]), int(g["iters"])), g["flow"])
    assert np.array_equal(farneback_np.farneback(g["mov8"], g["ref8"], int(g["win8"]), int(g["iters8"])), g["flow8"])


def test_golden_tileflow():
    g = gold("tileflow.npz")
    for be in (rf.NpBackend(), rf.CvBackend()):
        got = rf.calc_flow(g["ref"], g["mov"], int(g["T"]), int(g["ov"]), int(g["win"]), int(g["iters"]), be)
        assert np.array_equal(got, g["flow"])


def test_golden_warper():
    g = gold("warper.npz")
    for be in (rf.NpBackend(), rf.CvBackend()):
        assert np.array_equal(rf.warp(g["img16"], g["flow"], int(g["T"]), int(g["ov"]), be), g["out16"])
        assert np.array_equal(rf.warp(g["img8"], g["flow"], int(g["T"]), int(g["ov"]), be), g["out8"])


def test_golden_dog_nmi():
    g = gold("dog_nmi.npz")
    assert np.array_equal(cv_ops.dog(g["ref"]), g["d_ref"])
    assert np.array_equal(cv_ops.dog(g["mov"]), g["d_mov"])
    assert np.array_equal(cv_ops.dog(g["blobs"]), g["d_blobs"])
    be = rf.NpBackend()
    assert rf.mi_tiled(g["d_ref"], g["d_mov"], 1000, be) == pytest.approx(float(g["mi_whole"]), rel=1e-13)
    assert rf.mi_tiled(g["d_ref"], g["d_mov"], int(g["chunk_T"]), be) == pytest.approx(float(g["mi_chunks"]), rel=1e-13)


def test_golden_merge():
    g = gold("merge.npz")
    for be in (rf.NpBackend(), rf.CvBackend()):
        assert np.array_equal(rf.merge_flows_tiled(g["f1"], g["f2"], int(g["T"]), int(g["ov"]), be), g["merged"])


def test_golden_e2e():
    g = gold("e2e_small.npz")
    kw = dict(num_pyr_lvl=int(g["num_pyr_lvl"]), num_iterations=int(g["num_iterations"]), tile_size=int(g["tile_size"]),
              overlap=int(g["overlap"]), use_full_res_img=bool(g["use_full_res_img"]), use_dog=bool(g["use_dog"]))
    log = []
    flow = rf.register(g["ref"], g["mov"], be=rf.CvBackend(), log=log, **kw)
    assert np.array_equal(flow, g["flow"])
    assert np.array_equal(rf.warp(g["mov"], flow, kw["tile_size"], kw["overlap"], rf.CvBackend()), g["warped"])
    lines = str(g["stdout"]).splitlines()
    better = [l.strip().startswith("Better") for l in lines if "alignment than before" in l]
    assert better == [l["better"] for l in log]


# ---------------------------------------------------------------- tile geometry
@pytest.mark.parametrize("shape", [(2048, 2048), (2500, 3100), (999, 2001), (150, 130)])
def test_split_stitch_roundtrip(shape):
    rng = np.random.default_rng(6)
    a = rng.integers(0, 65535, shape).astype(np.uint16)
    tiles = rf.split(a, 1000, 100)
    assert all(t.shape == (1200, 1200) for t in tiles)
    assert np.array_equal(rf.stitch(tiles, shape[0], shape[1], 1000, 100), a)
    f = rng.standard_normal(shape + (2,)).astype(np.float32)
    assert np.array_equal(rf.stitch(rf.split(f, 1000, 100), shape[0], shape[1], 1000, 100), f)


FORCED = [("tft_full", (True, False, True), True), ("ft_nofull", (False, True), False), ("tf_nofull", (True, False), False)]


@pytest.mark.parametrize("name,dec,full_res", FORCED)
def test_golden_forced_decisions(name, dec, full_res):
    g = gold("forced_decisions.npz")
    kw = dict(num_pyr_lvl=int(g["num_pyr_lvl"]), num_iterations=int(g["num_iterations"]), tile_size=int(g["tile_size"]),
              overlap=int(g["overlap"]), use_full_res_img=full_res)
    got = rf.register(g["ref"], g["mov"], be=rf.CvBackend(), force_decisions=list(dec), **kw)
    assert np.array_equal(got, g[name])
