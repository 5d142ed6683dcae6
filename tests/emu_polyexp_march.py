"""CPU emulation (numpy, lanes as a vector of 32) of the warp-marching polynomial-expansion kernel
(fb_polyexp_march_kernel): validates the index logic -- virtual rows with REFLECT_101, rolling windows, edge selects,
shuffle source lanes -- bit-exactly against oracle.farneback_np.polyexp(prefilter3(window)).
Test infrastructure (imports the oracle); driven by tests/test_emu_cpu.py, or run it directly."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import farneback_np as fb  # noqa: E402

F32, F64 = np.float32, np.float64
OUTW = 28          # output columns per warp (lanes 2..29)
BAND = 24          # output rows per warp task (the kernel uses a larger band; small here to exercise the seams)


def refl(i, n):
    i = abs(i)
    return 2 * (n - 1) - i if i >= n else i


def march_strip(win, xs, ya, yb, out):
    Sh, Sw = win.shape
    g, xg, xxg, ig11, ig03, ig33, ig55 = fb.poly_gaussian(1, 1.7)
    g0, g1, xg1, xxg1 = g[1], g[2], xg[2], xxg[2]
    lanes = np.arange(32)
    xl = xs - 2 + lanes
    valid = (xl >= 0) & (xl < Sw)
    # shuffle sources (lane indices), clamped into the warp; lanes whose source is clamped hold don't-care values
    srcL = np.array([min(max(refl(x - 1, Sw) - (xs - 2), 0), 31) if 0 <= x < Sw else l for l, x in enumerate(xl)])
    srcR = np.array([min(max(refl(x + 1, Sw) - (xs - 2), 0), 31) if 0 <= x < Sw else l for l, x in enumerate(xl)])
    tL = np.array([min(max(min(max(x - 1, 0), Sw - 1) - (xs - 2), 0), 31) for x in xl])
    tR = np.array([min(max(min(max(x + 1, 0), Sw - 1) - (xs - 2), 0), 31) for x in xl])
    th0 = th1 = th2 = np.zeros(32, F32)
    P0 = P1 = P2 = np.zeros(32, F32)
    for v in range(ya - 2, yb + 2):            # virtual th row
        r = refl(v, Sh)
        raw = np.where(valid, win[r, np.clip(xl, 0, Sw - 1)], 0).astype(F32)
        th_new = raw * F32(0.5) + (raw[srcL] + raw[srcR]) * F32(0.25)
        th0, th1, th2 = th1, th2, th_new
        if v < ya:
            continue
        P_new = th1 * F32(0.5) + (th0 + th2) * F32(0.25)     # P at virtual row v - 1
        P0, P1, P2 = P1, P2, P_new
        y = v - 2
        if y < ya or y >= yb or y >= Sh:
            continue
        s0 = P1 if y == 0 else P0
        s1 = P1 if y == Sh - 1 else P2
        sc = P1
        pp = s0 + s1
        t0 = sc * g0 + g1 * pp
        t1 = F32(0) + xg1 * (s1 - s0)
        t2 = F32(0) + xxg1 * pp
        t0l, t0r, t1l, t1r, t2l, t2r = t0[tL], t0[tR], t1[tL], t1[tR], t2[tL], t2[tR]
        b1 = (t0 * g0).astype(F64)
        b3 = (t1 * g0).astype(F64)
        b5 = (t2 * g0).astype(F64)
        tg = (t0r + t0l).astype(F64)
        b1 = b1 + tg * F64(g1)
        b4 = tg * F64(xxg1)
        b2 = ((t0r - t0l) * xg1).astype(F64)
        b3 = b3 + ((t1r + t1l) * g1).astype(F64)
        b6 = ((t1r - t1l) * xg1).astype(F64)
        b5 = b5 + ((t2r + t2l) * g1).astype(F64)
        R = [(b3 * ig11).astype(F32), (b2 * ig11).astype(F32), (b1 * ig03 + b5 * ig33).astype(F32),
             (b1 * ig03 + b4 * ig33).astype(F32), (b6 * ig55).astype(F32)]
        for l in range(2, 30):
            x = xl[l]
            if x < Sw:
                for k in range(5):
                    out[k, y, x] = R[k][l]


def polyexp_march(win):
    Sh, Sw = win.shape
    out = np.full((5, Sh, Sw), np.nan, F32)
    for xs in range(0, Sw, OUTW):
        for ya in range(0, Sh, BAND):
            march_strip(win, xs, ya, min(ya + BAND, Sh), out)
    return out


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for shape in [(40, 70), (24, 28), (25, 29), (49, 57), (8, 8), (97, 31), (30, 113)]:
        win = rng.integers(0, 65535, shape).astype(np.uint16)
        want = fb.polyexp(fb.prefilter3(win))
        got = polyexp_march(win)
        assert not np.isnan(got).any(), shape
        assert np.array_equal(got, want), (shape, np.argwhere(got != want)[:5])
        print(shape, "ok")
