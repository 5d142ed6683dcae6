"""ORACLE (test infrastructure only -- never imported by the product path).

numpy restatement of ``cv2.calcOpticalFlowFarneback`` exactly as the reference calls it
(/root/reference/microaligner/optflow_reg/flow_calc.py:30-47):

    levels=0, pyr_scale=0.5, poly_n=1, poly_sigma=1.7, flags=OPTFLOW_FARNEBACK_GAUSSIAN,
    winsize=win, iterations=N, prev = moving tile, next = reference tile.

The arithmetic lives in OpenCV (``modules/video/src/optflowgf.cpp``; the reference pins
opencv-contrib-python==4.5.5.64, environment.yaml:75; this image ships 4.13.0 whose
baseline build is SSE3, i.e. *no FMA contraction* in this translation unit).  Every
float32 / float64 rounding step below follows the published C++ algorithm so that the
result is bit-identical to cv2 on this image (pinned by tests/test_oracle_ops.py and by
tests/golden/farneback_*.npz).

All arrays are numpy; float32 ops are done on float32 arrays so each op rounds once.
"""
import numpy as np

F32 = np.float32
F64 = np.float64


# ----------------------------------------------------------------------------- constants
def _chol_inv(A: np.ndarray) -> np.ndarray:
    """cv::Mat::inv(DECOMP_CHOLESKY) for a small SPD double matrix: OpenCV's in-place
    Cholesky (L stores reciprocal diagonals) followed by forward/back substitution on I."""
    m = A.shape[0]
    L = A.astype(F64).copy()
    b = np.eye(m, dtype=F64)
    for i in range(m):
        for j in range(i):
            s = L[i, j]
            for k in range(j):
                s -= L[i, k] * L[j, k]
            L[i, j] = s * L[j, j]
        s = L[i, i]
        for k in range(i):
            t = L[i, k]
            s -= t * t
        L[i, i] = 1.0 / np.sqrt(s)
    for i in range(m):
        for j in range(m):
            s = b[i, j]
            for k in range(i):
                s -= L[i, k] * b[k, j]
            b[i, j] = s * L[i, i]
    for i in range(m - 1, -1, -1):
        for j in range(m):
            s = b[i, j]
            for k in range(m - 1, i, -1):
                s -= L[k, i] * b[k, j]
            b[i, j] = s * L[i, i]
    return b


def poly_gaussian(n: int = 1, sigma: float = 1.7):
    """FarnebackPrepareGaussian: taps g, xg, xxg (f32, index n = centre) and the inverse-Gram
    constants ig11, ig03, ig33, ig55 (f64).  The Gram sums use *float* products widened to
    double, as the C++ does (g[y]*g[x]*x*x is a float expression) -- this changes the
    constants at the 1e-8 level, which is visible after rounding R to f32."""
    if sigma < 1.1920929e-07:  # FLT_EPSILON
        sigma = n * 0.3
    g = np.zeros(2 * n + 1, F32)
    s = 0.0
    for x in range(-n, n + 1):
        g[x + n] = F32(np.exp(-x * x / (2 * sigma * sigma)))
        s += float(g[x + n])
    s = 1.0 / s
    xg = np.zeros_like(g)
    xxg = np.zeros_like(g)
    for x in range(-n, n + 1):
        g[x + n] = F32(float(g[x + n]) * s)
        xg[x + n] = F32(F32(x) * g[x + n])
        xxg[x + n] = F32(F32(x * x) * g[x + n])
    G = np.zeros((6, 6), F64)
    for y in range(-n, n + 1):
        for x in range(-n, n + 1):
            gg = F32(g[y + n] * g[x + n])
            fx, fy = F32(x), F32(y)
            G[0, 0] += float(gg)
            G[1, 1] += float(F32(F32(gg * fx) * fx))
            G[3, 3] += float(F32(F32(F32(F32(gg * fx) * fx) * fx) * fx))
            G[5, 5] += float(F32(F32(F32(F32(gg * fx) * fx) * fy) * fy))
    G[2, 2] = G[0, 3] = G[0, 4] = G[3, 0] = G[4, 0] = G[1, 1]
    G[4, 4] = G[3, 3]
    G[3, 4] = G[4, 3] = G[5, 5]
    invG = _chol_inv(G)
    return g, xg, xxg, float(invG[1, 1]), float(invG[0, 3]), float(invG[3, 3]), float(invG[5, 5])


def blur_taps(win: int):
    """FarnebackUpdateFlow_GaussianBlur taps: kernel[0..m], m=win//2, sigma=0.3*m."""
    m = win // 2
    sigma = m * 0.3
    k = np.zeros(m + 1, F32)
    s = 1.0
    k[0] = F32(1.0)
    for i in range(1, m + 1):
        t = F32(np.exp(-i * i / (2 * sigma * sigma)))
        k[i] = t
        s += float(t) * 2
    s = 1.0 / s
    for i in range(m + 1):
        k[i] = F32(float(k[i]) * s)
    return k


BORDER = np.array([0.14, 0.14, 0.4472, 0.4472, 0.4472], F32)


# ----------------------------------------------------------------------------- stages
def _r101(i, n):
    i = np.abs(i)
    return np.where(i >= n, 2 * (n - 1) - i, i)


def prefilter3(img: np.ndarray) -> np.ndarray:
    """convertTo(CV_32F) + GaussianBlur 3x3 (taps 1/4,1/2,1/4; REFLECT_101), rows then cols.

    Products by 0.25/0.5 are exact, so only the additions round; the symmetric
    (a+b)*0.25 + c*0.5 form is value-identical to any other association as long as
    no intermediate overflows -- verified bit-exact vs cv2 in tests."""
    f = img.astype(F32)
    h, w = f.shape
    xs = np.arange(w)
    l, r = _r101(xs - 1, w), _r101(xs + 1, w)
    t = f[:, xs] * F32(0.5) + (f[:, l] + f[:, r]) * F32(0.25)
    ys = np.arange(h)
    u, d = _r101(ys - 1, h), _r101(ys + 1, h)
    return t[ys] * F32(0.5) + (t[u] + t[d]) * F32(0.25)


def polyexp(I: np.ndarray, n: int = 1, sigma: float = 1.7) -> np.ndarray:
    """FarnebackPolyExp for n=1 -> (5,H,W) planar f32 [R0..R4] (OpenCV stores them interleaved)."""
    assert n == 1
    g, xg, xxg, ig11, ig03, ig33, ig55 = poly_gaussian(n, sigma)
    h, w = I.shape
    ys = np.arange(h)
    up = np.maximum(ys - 1, 0)
    dn = np.minimum(ys + 1, h - 1)
    s0, s1 = I[up], I[dn]
    p = s0 + s1
    t0 = I * g[1] + g[2] * p                       # row0 = srow*g0 ; += g[k]*p
    t1 = F32(0) + xg[2] * (s1 - s0)
    t2 = F32(0) + xxg[2] * p
    # horizontal: columns replicate
    xs = np.arange(w)
    lf = np.maximum(xs - 1, 0)
    rt = np.minimum(xs + 1, w - 1)
    g0 = g[1]
    gk, xgk, xxgk = g[2], xg[2], xxg[2]
    b1 = (t0 * g0).astype(F64)
    b3 = (t1 * g0).astype(F64)
    b5 = (t2 * g0).astype(F64)
    tg = (t0[:, rt] + t0[:, lf]).astype(F64)       # float add, then widened
    b1 = b1 + tg * F64(gk)
    b4 = tg * F64(xxgk)
    b2 = ((t0[:, rt] - t0[:, lf]) * xgk).astype(F64)   # float*float, then widened
    b3 = b3 + ((t1[:, rt] + t1[:, lf]) * gk).astype(F64)
    b6 = ((t1[:, rt] - t1[:, lf]) * xgk).astype(F64)
    b5 = b5 + ((t2[:, rt] + t2[:, lf]) * gk).astype(F64)
    R = np.empty((5, h, w), F32)
    R[1] = (b2 * ig11).astype(F32)
    R[0] = (b3 * ig11).astype(F32)
    R[3] = (b1 * ig03 + b4 * ig33).astype(F32)
    R[2] = (b1 * ig03 + b5 * ig33).astype(F32)
    R[4] = (b6 * ig55).astype(F32)
    return R


def update_matrices(R0: np.ndarray, R1: np.ndarray, flow: np.ndarray) -> np.ndarray:
    """FarnebackUpdateMatrices over the whole tile. R0,R1: (5,H,W); flow: (H,W,2) -> M (5,H,W)."""
    _, h, w = R0.shape
    x = np.arange(w, dtype=F32)[None, :]
    y = np.arange(h, dtype=F32)[:, None]
    dx = flow[..., 0]
    dy = flow[..., 1]
    fx = x + dx
    fy = y + dy
    x1 = np.floor(fx).astype(np.int64)
    y1 = np.floor(fy).astype(np.int64)
    # cvFloor works on the float value; int conversion of huge values is UB in C -- clamp
    fx = fx - x1.astype(F32)
    fy = fy - y1.astype(F32)
    inside = (x1 >= 0) & (x1 < w - 1) & (y1 >= 0) & (y1 < h - 1)
    xc = np.clip(x1, 0, max(w - 2, 0))
    yc = np.clip(y1, 0, max(h - 2, 0))
    one = F32(1)
    a00 = (one - fx) * (one - fy)
    a01 = fx * (one - fy)
    a10 = (one - fx) * fy
    a11 = fx * fy
    xc1 = np.minimum(xc + 1, w - 1)
    yc1 = np.minimum(yc + 1, h - 1)

    def samp(c):
        P = R1[c]
        return ((a00 * P[yc, xc] + a01 * P[yc, xc1]) + a10 * P[yc1, xc]) + a11 * P[yc1, xc1]

    r2 = np.where(inside, samp(0), F32(0))
    r3 = np.where(inside, samp(1), F32(0))
    r4 = np.where(inside, (R0[2] + samp(2)) * F32(0.5), R0[2])
    r5 = np.where(inside, (R0[3] + samp(3)) * F32(0.5), R0[3])
    r6 = np.where(inside, (R0[4] + samp(4)) * F32(0.25), R0[4] * F32(0.5))
    r2 = (R0[0] - r2) * F32(0.5)
    r3 = (R0[1] - r3) * F32(0.5)
    r2 = r2 + (r4 * dy + r6 * dx)
    r3 = r3 + (r6 * dy + r5 * dx)
    # border damping: scale = bx_lo * bx_hi * by_lo * by_hi (left-to-right float products)
    sx_lo = np.ones(w, F32)
    sx_hi = np.ones(w, F32)
    sy_lo = np.ones(h, F32)
    sy_hi = np.ones(h, F32)
    for i in range(5):
        if i < w:
            sx_lo[i] = BORDER[i]
        if w - 1 - i >= 0:
            sx_hi[w - 1 - i] = BORDER[i]
        if i < h:
            sy_lo[i] = BORDER[i]
        if h - 1 - i >= 0:
            sy_hi[h - 1 - i] = BORDER[i]
    scale = ((sx_lo * sx_hi)[None, :] * sy_lo[:, None]) * sy_hi[:, None]
    xi = np.arange(w)[None, :]
    yi = np.arange(h)[:, None]
    damp = (xi < 5) | (xi >= w - 5) | (yi < 5) | (yi >= h - 5)
    scale = np.where(damp, scale, F32(1)).astype(F32)
    r2, r3, r4, r5, r6 = (np.where(damp, v * scale, v) for v in (r2, r3, r4, r5, r6))
    M = np.empty((5, h, w), F32)
    M[0] = r4 * r4 + r6 * r6
    M[1] = (r4 + r5) * r6
    M[2] = r5 * r5 + r6 * r6
    M[3] = r4 * r2 + r6 * r3
    M[4] = r6 * r2 + r5 * r3
    return M


def gaussian_blur_sym(M: np.ndarray, k: np.ndarray) -> np.ndarray:
    """Separable window blur of (C,H,W) f32: vertical (rows clamped) then horizontal
    (columns replicated), each `s = c*k0; for i: s += (a_{+i} + a_{-i}) * k_i`, mul and add
    rounded separately (SSE2 baseline, no FMA)."""
    m = len(k) - 1
    C, h, w = M.shape
    ys = np.arange(h)
    V = M * k[0]
    for i in range(1, m + 1):
        V = V + (M[:, np.minimum(ys + i, h - 1)] + M[:, np.maximum(ys - i, 0)]) * k[i]
    xs = np.arange(w)
    Hh = V * k[0]
    for i in range(1, m + 1):
        Hh = Hh + k[i] * (V[:, :, np.maximum(xs - i, 0)] + V[:, :, np.minimum(xs + i, w - 1)])
    return Hh


def solve_flow(B: np.ndarray) -> np.ndarray:
    g11, g12, g22, h1, h2 = (B[i].astype(F64) for i in range(5))
    idet = 1.0 / (g11 * g22 - g12 * g12 + 1e-3)
    flow = np.empty(B.shape[1:] + (2,), F32)
    flow[..., 0] = ((g11 * h2 - g12 * h1) * idet).astype(F32)
    flow[..., 1] = ((g22 * h1 - g12 * h2) * idet).astype(F32)
    return flow


def farneback(prev: np.ndarray, nxt: np.ndarray, win: int, iters: int, return_intermediates=False):
    """Flow (H,W,2) f32 such that prev(p) ~ next(p + flow(p))."""
    R0 = polyexp(prefilter3(prev))
    R1 = polyexp(prefilter3(nxt))
    h, w = prev.shape
    flow = np.zeros((h, w, 2), F32)
    M = update_matrices(R0, R1, flow)
    k = blur_taps(win)
    inter = dict(R0=R0, R1=R1, M0=M.copy())
    for it in range(iters):
        B = gaussian_blur_sym(M, k)
        flow = solve_flow(B)
        if it < iters - 1:
            M = update_matrices(R0, R1, flow)
    if return_intermediates:
        return flow, inter
    return flow
