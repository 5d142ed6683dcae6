"""-m gpu: every CUDA operator, called through the C ABI, against the CPU oracle. Bit-exact unless stated."""
import numpy as np
import pytest
import torch

from oracle import cv_ops, farneback_np
from oracle import reference_flow as rf
from tests.util import blobs, random_flow, synth_pair

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops(cuda):
    from microaligner_b200 import ops as _ops
    return _ops


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("dtype", [np.uint8, np.uint16])
@pytest.mark.parametrize("shape", [(200, 300), (201, 303), (257, 129), (3, 5)])
def test_pyrdown(ops, dtype, shape):
    rng = np.random.default_rng(0)
    img = rng.integers(0, np.iinfo(dtype).max + 1, shape).astype(dtype)
    got = ops.pyr_down(dev(img)).cpu().numpy()
    assert np.array_equal(got, cv_ops.pyr_down(img))


@pytest.mark.parametrize("case", [(100, 130, 200, 260), (100, 130, 199, 259), (101, 77, 201, 154), (64, 64, 128, 127), (2, 2, 3, 3)])
@pytest.mark.parametrize("scale", [1.0, 2.0, 4.0])
def test_pyrup_flow(ops, case, scale):
    h, w, dh, dw = case
    f = random_flow(h, w, 1)
    got = ops.pyr_up_flow(dev(f), (dh, dw), scale).cpu().numpy()
    assert np.array_equal(got, cv_ops.pyr_up_f32c2(f, (dh, dw), scale))


@pytest.mark.parametrize("dtype", [np.uint8, np.uint16])
@pytest.mark.parametrize("geom", [((311, 277), 100, 20), ((257, 400), 64, 10), ((90, 130), 1000, 100), ((1234, 777), 500, 33)])
def test_warp_tiles(ops, dtype, geom):
    (h, w), T, ov = geom
    rng = np.random.default_rng(2)
    img = rng.integers(0, np.iinfo(dtype).max + 1, (h, w)).astype(dtype)
    flow = random_flow(h, w, 3, mag=6.0)
    flow[0, 0] = [1e6, -1e6]
    flow[1, 1] = [np.float32(1e10), 3]
    flow[2, 2] = [0.5, 0.5]
    flow[3, 3] = [-ov - 0.5, ov + 2.25]
    got = ops.warp_tiles(dev(img), dev(flow), T, ov).cpu().numpy()
    want = rf.warp(img, flow, T, ov, rf.NpBackend())
    assert np.array_equal(got, want)


@pytest.mark.parametrize("geom", [((311, 277), 100, 20), ((250, 420), 64, 10), ((90, 130), 1000, 100)])
def test_merge_flows(ops, geom):
    (h, w), T, ov = geom
    f1 = random_flow(h, w, 4, mag=3.0)
    f2 = random_flow(h, w, 5, mag=3.0)
    f1[:T, :T] = 0            # a tile whose first flow is all zero -> takes flow2
    f2[:T, T:2 * T] = -1.0    # a tile whose second flow has max < 0 -> general branch
    got = ops.merge_flows_tiles(dev(f1), dev(f2), T, ov).cpu().numpy()
    want = rf.merge_flows_tiled(f1, f2, T, ov, rf.NpBackend())
    assert np.array_equal(got, want)
    z = np.zeros_like(f1)
    assert np.array_equal(ops.merge_flows_tiles(dev(f1), dev(z), T, ov).cpu().numpy(), f1)


@pytest.mark.parametrize("case", [((300, 260), np.uint16, 99, 3), ((257, 301), np.uint8, 99, 1), ((400, 400), np.uint16, 31, 2),
                                  ((128, 131), np.uint8, 9, 2), ((333, 1030), np.uint16, 99, 2)])
def test_farneback_untiled(ops, case):
    shape, dtype, win, iters = case
    ref, mov = synth_pair(shape[0], shape[1], 0, dtype)
    got = ops.farneback_tiles(dev(mov), dev(ref), 0, 0, win, iters).cpu().numpy()
    want = farneback_np.farneback(mov, ref, win, iters)
    assert np.array_equal(got, want), f"max |d| = {np.abs(got - want).max()}"


@pytest.mark.parametrize("case", [((500, 700), np.uint16, 200, 40, 2), ((450, 333), np.uint8, 150, 25, 3)])
def test_farneback_tiled(ops, case):
    shape, dtype, T, ov, iters = case
    ref, mov = synth_pair(shape[0], shape[1], 1, dtype)
    win = ov - (1 - ov % 2)
    got = ops.farneback_tiles(dev(mov), dev(ref), T, ov, win, iters).cpu().numpy()
    want = rf.calc_flow(ref, mov, T, ov, win, iters, rf.NpBackend())
    assert got.shape == want.shape
    assert np.array_equal(got, want), f"max |d| = {np.abs(got - want).max()}"


def test_farneback_zero_and_padded(ops):
    ref, mov = synth_pair(260, 300, 2, np.uint16)
    z = np.zeros_like(ref)
    got = ops.farneback_tiles(dev(z), dev(z), 0, 0, 99, 3).cpu().numpy()
    assert not got.any()
    # small batches (workspace for a single tile) must give the same answer as one big batch
    big = ops.farneback_tiles(dev(mov), dev(ref), 100, 20, 19, 2).cpu().numpy()
    old = ops.FARNEBACK_WORKSPACE_BUDGET
    try:
        ops.FARNEBACK_WORKSPACE_BUDGET = 1
        ops.release_workspaces()
        small = ops.farneback_tiles(dev(mov), dev(ref), 100, 20, 19, 2).cpu().numpy()
    finally:
        ops.FARNEBACK_WORKSPACE_BUDGET = old
        ops.release_workspaces()
    assert np.array_equal(big, small)
    # tile sub-range: only the centres of the requested tiles are written
    part = ops.farneback_tiles(dev(mov), dev(ref), 100, 20, 19, 2, tile_range=(3, 5)).cpu().numpy()
    mask = np.zeros(ref.shape, bool)
    mask[100:200, 0:200] = True   # tiles 3 and 4 of a 3x3 grid: row 1, columns 0-1
    assert np.array_equal(part[mask], big[mask]) and not part[~mask].any()


@pytest.mark.parametrize("case", [((300, 260), np.uint16), ((517, 333), np.uint8), ((128, 1024), np.uint16), ((21, 21), np.uint8)])
def test_dog(ops, case):
    shape, dtype = case
    ref, mov = synth_pair(shape[0], shape[1], 3, dtype)
    for img in (ref, mov, blobs(shape[0], shape[1], 1, dtype)):
        got = ops.dog_u8(dev(img)).cpu().numpy()
        assert np.array_equal(got, cv_ops.dog(img))
    z = np.zeros(shape, dtype)
    assert not ops.dog_u8(dev(z)).cpu().numpy().any()
    c = np.full(shape, 7, dtype)
    assert not ops.dog_u8(dev(c)).cpu().numpy().any()


def test_nmi(ops):
    ref, mov = synth_pair(700, 900, 4, np.uint16)
    a, b = cv_ops.dog(ref), cv_ops.dog(mov)
    whole = ops.nmi_chunks(dev(a), dev(b), a.size).cpu().numpy()
    assert whole.shape == (1,)
    assert whole[0] == pytest.approx(cv_ops.nmi(a, b), rel=1e-12)
    chunk = 300 * 300
    got = ops.nmi_chunks(dev(a), dev(b), chunk).cpu().numpy()
    fa, fb = a.ravel(), b.ravel()
    want = [cv_ops.nmi(fa[s:s + chunk], fb[s:s + chunk]) for s in range(0, fa.size, chunk)]
    assert got.shape == (len(want),)
    np.testing.assert_allclose(got, want, rtol=1e-12)
    # special cases: both constant -> 1, one constant -> 0, identical -> 1
    z = np.zeros_like(a)
    assert ops.nmi_chunks(dev(z), dev(z), z.size).cpu().numpy()[0] == 1.0
    assert ops.nmi_chunks(dev(z), dev(b), z.size).cpu().numpy()[0] == 0.0
    assert ops.nmi_chunks(dev(a), dev(a), a.size).cpu().numpy()[0] == pytest.approx(1.0, rel=1e-12)
    # both comparisons of the gate in one launch == two single launches, bit for bit; unaligned chunk starts; noise
    import torch
    rng = np.random.default_rng(0)
    c = rng.integers(0, 256, a.shape).astype(np.uint8)
    for chunk in (300 * 300, 150 * 150 + 7):
        n = -(-a.size // chunk)
        s0 = torch.zeros(n, dtype=torch.float64, device="cuda")
        s1 = torch.zeros(n, dtype=torch.float64, device="cuda")
        ops.nmi_chunk_range2(dev(a), dev(b), dev(c), chunk, (0, n), s0, s1)
        assert torch.equal(s0, ops.nmi_chunks(dev(a), dev(b), chunk)) and torch.equal(s1, ops.nmi_chunks(dev(a), dev(c), chunk))
        fc = c.ravel()
        np.testing.assert_allclose(s1.cpu().numpy(), [cv_ops.nmi(fa[s:s + chunk], fc[s:s + chunk]) for s in range(0, fa.size, chunk)],
                                   rtol=1e-12)


def test_zmip(ops):
    import cv2
    rng = np.random.default_rng(5)
    for dtype in (np.uint8, np.uint16):
        pages = [rng.integers(3, np.iinfo(dtype).max // 2, (211, 317)).astype(dtype) for _ in range(4)]
        want = cv2.normalize(np.maximum.reduce(pages), None, 0, 255, cv2.NORM_MINMAX, cv2.CV_8U)
        got = ops.zmip_normalize_u8([dev(p) for p in pages]).cpu().numpy()
        assert np.array_equal(got, want)


def test_minmax(ops):
    rng = np.random.default_rng(6)
    img = rng.integers(5, 60000, (300, 401)).astype(np.uint16)
    mm = ops.minmax(dev(img)).cpu().numpy()
    assert mm[0] == img.min() and mm[1] == img.max()


def test_farneback_contract_fma_within_contract(ops):
    """Opt-in FMA-contracted window blur: not bit-identical, but far inside the 0.01 / 0.1 px contract per call."""
    ref, mov = synth_pair(520, 610, 3, np.uint16)
    want = rf.calc_flow(ref, mov, 250, 40, 39, 3, rf.NpBackend())
    exact = ops.farneback_tiles(dev(mov), dev(ref), 250, 40, 39, 3).cpu().numpy()
    fast = ops.farneback_tiles(dev(mov), dev(ref), 250, 40, 39, 3, contract_fma=True).cpu().numpy()
    assert np.array_equal(exact, want)
    epe = np.sqrt(((fast - want) ** 2).sum(-1))
    assert epe.mean() <= 1e-4 and epe.max() <= 1e-3, (epe.mean(), epe.max())
    assert not np.array_equal(fast, want)   # it really is the other arithmetic
