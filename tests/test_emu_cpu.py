"""CPU: lane-level emulations of kernel index logic that cannot be run here (no GPU) against the oracle."""
import numpy as np
import pytest

from oracle import farneback_np as fb
from tests import emu_polyexp_march as emu


@pytest.mark.parametrize("shape", [(40, 70), (24, 28), (25, 29), (49, 57), (8, 8), (97, 31), (30, 113)])
def test_marching_polyexp_index_logic(shape):
    """fb_polyexp_march_kernel (csrc/farneback.cu): virtual rows with REFLECT_101, rolling windows, edge selects,
    shuffle sources and band / strip seams reproduce prefilter3 + polyexp bit for bit."""
    win = np.random.default_rng(shape[0]).integers(0, 65535, shape).astype(np.uint16)
    got = emu.polyexp_march(win)
    assert not np.isnan(got).any()
    assert np.array_equal(got, fb.polyexp(fb.prefilter3(win)))


@pytest.mark.parametrize("K", [4, 8])
def test_sliding_window_rotation(K):
    """conv8x2 (csrc/farneback.cu): with the register windows rotated as the kernels
    do, output j of a thread meets exactly in[j + i] and in[j - i] at tap i, for every half-width m."""
    for m in range(0, 97):
        wp, wm = list(range(K)), list(range(K))          # slot -> input row relative to the centre of output 0
        seen = [[] for _ in range(K)]
        pp, pm, i = K, -1, 1

        def step(i, s):
            wp[s] = pp + s
            wm[(K - 1 - s) & (K - 1)] = pm - s
            for j in range(K):
                assert wp[(j + s + 1) & (K - 1)] == j + i + s and wm[(j + 8 * K - 1 - s) & (K - 1)] == j - i - s
                seen[j].append(i + s)

        while i + K - 1 <= m:
            for s in range(K):
                step(i, s)
            pp, pm, i = pp + K, pm - K, i + K
        for s in range(K - 1):
            if i + s <= m:
                step(i, s)
        assert all(taps == list(range(1, m + 1)) for taps in seen)


@pytest.mark.parametrize("addr_a,addr_b", [(256, 512), (256, 515), (259, 1027)])
def test_nmi_run_length_partition(addr_a, addr_b):
    """nmi_hist_rle_kernel (csrc/nmi.cu): scalar head up to the first 16-byte boundary, 16-pixel vectors with run-length
    merging, scalar tail -- every pixel is counted exactly once, whatever the alignment and chunk size."""
    rng = np.random.default_rng(0)
    n, gsz = 1000, 64
    a, b = rng.integers(0, 4, n), rng.integers(0, 3, n)
    for beg, chunk in [(0, 1000), (0, 337), (337, 337), (674, 337), (5, 7), (990, 337)]:
        H = np.zeros(65536, np.int64)
        end = min(beg + chunk, n)
        ln = end - beg
        pa, pb = addr_a + beg, addr_b + beg
        vec = ((pa ^ pb) & 15) == 0
        head = min(((16 - (pa & 15)) & 15) if vec else ln, ln)
        nvec = (ln - head) // 16
        tail0 = head + nvec * 16
        hits = np.zeros(ln, int)
        for t in range(gsz):
            for i in list(range(t, head, gsz)) + list(range(tail0 + t, ln, gsz)):
                H[(a[beg + i] << 8) | b[beg + i]] += 1
                hits[i] += 1
            for v in range(t, nvec, gsz):
                run, cnt = 0, 0
                for q in range(16):
                    i = head + v * 16 + q
                    key = (a[beg + i] << 8) | b[beg + i]
                    hits[i] += 1
                    if cnt and key == run:
                        cnt += 1
                    else:
                        if cnt:
                            H[run] += cnt
                        run, cnt = key, 1
                H[run] += cnt
        assert (hits == 1).all()
        want = np.zeros(65536, np.int64)
        np.add.at(want, (a[beg:end] << 8) | b[beg:end], 1)
        assert np.array_equal(H, want)


def _pyrdown_march_emu(img, base_misaligned, rows_per_cta=32):
    """pyrdown_march_kernel (csrc/pyramid.cu) thread by thread: interior columns read three aligned PAIRS per source row
    -- (cx-2, cx-1) (cx, cx+1) (cx+2, cx+3), or (cx-3, cx-2) (cx-1, cx) (cx+1, cx+2) when the row starts in the middle of a
    pair -- the others five scalar REFLECT_101 taps; a five-deep rolling window of row sums marches down the column."""
    h, w = img.shape
    bits = 8 * img.dtype.itemsize
    mask = (1 << bits) - 1
    oh, ow = (h + 1) // 2, (w + 1) // 2
    out = np.zeros((oh, ow), img.dtype)
    refl = lambda i, n: -i if i < 0 else (2 * n - 2 - i if i >= n else i)          # noqa: E731
    flat = img.reshape(-1).astype(np.int64)
    # element offset of the first pixel from a pair boundary: the device tensor is dense, so row y starts at y*w (+ base)
    base = 1 if base_misaligned else 0

    def pair(elem):                       # the aligned pair whose LOW half is element `elem` (must be on a pair boundary)
        assert (elem + base) % 2 == 0 and 0 <= elem and elem + 1 < flat.size + 1
        lo = flat[elem]
        hi = flat[elem + 1] if elem + 1 < flat.size else 0
        return int(lo | (hi << bits))

    for ox in range(ow):
        cx = 2 * ox
        interior = cx >= 3 and cx + 3 < w

        def hrow(y):
            yy = refl(y, h)
            if interior:
                mid = (yy * w + base) % 2 != 0
                q = yy * w + cx - (3 if mid else 2)
                a, b, c = pair(q), pair(q + 2), pair(q + 4)
                if mid:
                    return (a >> bits) + 4 * (b & mask) + 6 * (b >> bits) + 4 * (c & mask) + (c >> bits)
                return (a & mask) + 4 * (a >> bits) + 6 * (b & mask) + 4 * (b >> bits) + (c & mask)
            xs = [refl(cx + d - 2, w) for d in range(5)]
            r = img[yy].astype(np.int64)
            return int(r[xs[0]] + 4 * r[xs[1]] + 6 * r[xs[2]] + 4 * r[xs[3]] + r[xs[4]])

        for oy0 in range(0, oh, rows_per_cta):
            r0, r1, r2 = hrow(2 * oy0 - 2), hrow(2 * oy0 - 1), hrow(2 * oy0)
            for oy in range(oy0, min(oy0 + rows_per_cta, oh)):
                r3, r4 = hrow(2 * oy + 1), hrow(2 * oy + 2)
                out[oy, ox] = (r0 + 4 * r1 + 6 * r2 + 4 * r3 + r4 + 128) >> 8
                r0, r1, r2 = r2, r3, r4
    return out


@pytest.mark.parametrize("dtype", [np.uint8, np.uint16])
@pytest.mark.parametrize("shape", [(3, 3), (7, 8), (9, 7), (34, 41), (70, 12), (33, 66)])
@pytest.mark.parametrize("misaligned", [False, True])
def test_marching_pyrdown_pair_logic(shape, dtype, misaligned):
    """Odd widths make the row alignment alternate from row to row; either way the pair unpacking, the interior test and
    the rolling window reproduce cv2.pyrDown bit for bit."""
    import cv2
    rng = np.random.default_rng(shape[0] * 100 + shape[1])
    img = rng.integers(0, np.iinfo(dtype).max, shape, dtype=dtype, endpoint=True)
    assert np.array_equal(_pyrdown_march_emu(img, misaligned, rows_per_cta=5), cv2.pyrDown(img))
