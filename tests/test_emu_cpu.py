"""CPU: lane-level emulations of kernel index logic that cannot be run here (no GPU) against the oracle."""
import numpy as np
import pytest

from oracle import farneback_np as fb
from tests import emu_polyexp_march as emu


@pytest.mark.parametrize("shape", [(40, 70), (24, 28), (25, 29), (49, 57), (8, 8), (97, 31), (30, 113)])
def test_marching_polyexp_index_logic(shape):
    """fb_polyexp_march_kernel (csrc/farneback_variants.cuh): virtual rows with REFLECT_101, rolling windows, edge selects,
    shuffle sources and band / strip seams reproduce prefilter3 + polyexp bit for bit."""
    win = np.random.default_rng(shape[0]).integers(0, 65535, shape).astype(np.uint16)
    got = emu.polyexp_march(win)
    assert not np.isnan(got).any()
    assert np.array_equal(got, fb.polyexp(fb.prefilter3(win)))


@pytest.mark.parametrize("K", [4, 8])
def test_sliding_window_rotation(K):
    """conv8x2 / convKx2 (csrc/farneback.cu, farneback_variants.cuh): with the register windows rotated as the kernels
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
