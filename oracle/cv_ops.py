"""ORACLE (test infrastructure only -- never imported by the product path).

numpy restatements of the OpenCV / scikit-learn calls that sit on the reference's
non-linear registration path, each with the reference call site it stands for:

  remap_linear      cv.remap(.., INTER_LINEAR)      optflow_reg/warper.py:65, optflow_registrator.py:45
  pyr_down          cv.pyrDown                      optflow_reg/optflow_registrator.py:194
  pyr_up_f32c2      cv.pyrUp(flow*k, dstsize)       optflow_reg/optflow_registrator.py:140,150,164,169,212
  dog               OptFlowRegistrator.dog          optflow_reg/optflow_registrator.py:249-274
  nmi               sklearn normalized_mutual_info_score   shared_modules/similarity_scoring.py:36,44

The arithmetic itself is OpenCV's (not vendored by the reference; pinned
opencv-contrib-python==4.5.5.64 in environment.yaml:75, 4.13.0 in this image) and
scikit-learn's (pinned 1.0.2, environment.yaml:87; 1.9.0 here).  These restatements follow the
published algorithms; tests/test_oracle_ops.py pins them against the live cv2 / sklearn of
this image and against tests/golden/*.npz.
"""
import numpy as np

F32 = np.float32
F64 = np.float64


def r101(i, n):
    """BORDER_REFLECT_101 index map (valid for |overshoot| < n)."""
    i = np.abs(np.asarray(i))
    return np.where(i >= n, 2 * (n - 1) - i, i)


# ------------------------------------------------------------------------------- remap
def remap_linear(src: np.ndarray, mapxy: np.ndarray) -> np.ndarray:
    """cv.remap(src, mapxy(H,W,2 f32), None, INTER_LINEAR), BORDER_CONSTANT 0.

    Coordinates are quantised to 1/32 px: s = cvRound(map*32) (round half to even on the f32
    product), integer part s>>5 saturated to int16, fraction s&31.  u8 uses 15-bit integer
    weights, u16/f32 float weights summed left to right without FMA; u16 rounds half-even."""
    h, w = src.shape[:2]
    sx = np.rint(mapxy[..., 0] * F32(32)).astype(np.int64)
    sy = np.rint(mapxy[..., 1] * F32(32)).astype(np.int64)
    sx = np.clip(sx, -(2 ** 31), 2 ** 31 - 1)
    sy = np.clip(sy, -(2 ** 31), 2 ** 31 - 1)
    ix = np.clip(sx >> 5, -32768, 32767)
    iy = np.clip(sy >> 5, -32768, 32767)
    ax = (sx & 31)
    ay = (sy & 31)

    def tap(yy, xx):
        ok = (yy >= 0) & (yy < h) & (xx >= 0) & (xx < w)
        v = src[np.clip(yy, 0, h - 1), np.clip(xx, 0, w - 1)]
        if src.ndim == 3:
            ok = ok[..., None]
        return np.where(ok, v, 0)

    t00, t01, t10, t11 = tap(iy, ix), tap(iy, ix + 1), tap(iy + 1, ix), tap(iy + 1, ix + 1)
    if src.dtype == np.uint8:
        w00 = (32 - ay) * (32 - ax) * 32
        w01 = (32 - ay) * ax * 32
        w10 = ay * (32 - ax) * 32
        w11 = ay * ax * 32
        acc = t00.astype(np.int64) * w00 + t01.astype(np.int64) * w01 \
            + t10.astype(np.int64) * w10 + t11.astype(np.int64) * w11
        return ((acc + 16384) >> 15).astype(np.uint8)
    fx = ax.astype(F32) * F32(1 / 32)
    fy = ay.astype(F32) * F32(1 / 32)
    one = F32(1)
    w00 = (one - fy) * (one - fx)
    w01 = (one - fy) * fx
    w10 = fy * (one - fx)
    w11 = fy * fx
    if src.ndim == 3:
        w00, w01, w10, w11 = (v[..., None] for v in (w00, w01, w10, w11))
    acc = ((t00.astype(F32) * w00 + t01.astype(F32) * w01) + t10.astype(F32) * w10) + t11.astype(F32) * w11
    if src.dtype == np.uint16:
        return np.clip(np.rint(acc), 0, 65535).astype(np.uint16)
    return acc.astype(F32)


# ------------------------------------------------------------------------------- pyramids
def pyr_down(img: np.ndarray) -> np.ndarray:
    """cv.pyrDown for u8/u16: 5x5 binomial, REFLECT_101, (s + 128) >> 8, out ((H+1)//2,(W+1)//2)."""
    h, w = img.shape
    oh, ow = (h + 1) // 2, (w + 1) // 2
    k = (1, 4, 6, 4, 1)
    a = img.astype(np.int64)
    xs = 2 * np.arange(ow)
    row = sum(k[d + 2] * a[:, r101(xs + d, w)] for d in range(-2, 3))
    ys = 2 * np.arange(oh)
    s = sum(k[d + 2] * row[r101(ys + d, h)] for d in range(-2, 3))
    return ((s + 128) >> 8).astype(img.dtype)


def _pyr_up_axis_last(s: np.ndarray, dn: int) -> np.ndarray:
    """Horizontal half of cv.pyrUp on (..., n) f32 -> (..., dn), *unscaled* (sum of weights 8).

    interior even: (s[i-1] + s[i]*6) + s[i+1]     odd: (s[i] + s[i+1])*4
    left edge even: s[0]*6 + s[1]*2               right edge even: s[n-2] + s[n-1]*7, odd: s[n-1]*8
    """
    n = s.shape[-1]
    out = np.empty(s.shape[:-1] + (2 * n,), F32)
    c6, c4, c2, c7, c8 = F32(6), F32(4), F32(2), F32(7), F32(8)
    if n == 1:
        out[..., 0] = s[..., 0] * c8
        out[..., 1] = s[..., 0] * c8
        return out[..., :dn]
    out[..., 2:2 * n - 2:2] = (s[..., 0:n - 2] + s[..., 1:n - 1] * c6) + s[..., 2:n]
    out[..., 1:2 * n - 1:2] = (s[..., 0:n - 1] + s[..., 1:n]) * c4
    out[..., 0] = s[..., 0] * c6 + s[..., 1] * c2
    out[..., 2 * n - 2] = s[..., n - 2] + s[..., n - 1] * c7
    out[..., 2 * n - 1] = s[..., n - 1] * c8
    return out[..., :dn]


def pyr_up_f32c2(flow: np.ndarray, dsize_hw, scale: float = 1.0) -> np.ndarray:
    """cv.pyrUp(flow * scale, dstsize=(W',H')) for (h,w,2) f32, H' in {2h-1,2h}, W' in {2w-1,2w}.

    Rows then columns; the vertical pass uses the generic 3-row form with the row index map
    top: reflect-101, bottom: replicate; result * 1/64 (exact)."""
    h, w, _ = flow.shape
    dh, dw = dsize_hw
    assert abs(dh - 2 * h) == dh % 2 and abs(dw - 2 * w) == dw % 2
    f = flow * F32(scale)
    rows = np.empty((h, dw, 2), F32)
    for c in range(2):
        rows[..., c] = _pyr_up_axis_last(np.ascontiguousarray(f[..., c]), dw)
    ys = np.arange(h)
    r0 = rows[r101(ys - 1, h) if h > 1 else np.zeros(h, int)]
    r1 = rows
    r2 = rows[np.minimum(ys + 1, h - 1)]
    out = np.empty((2 * h, dw, 2), F32)
    out[0::2] = ((r0 + r1 * F32(6)) + r2) * F32(1 / 64)
    out[1::2] = ((r1 + r2) * F32(4)) * F32(1 / 64)
    return np.ascontiguousarray(out[:dh])


# ------------------------------------------------------------------------------- DoG
def gaussian_kernel_41(sigma: float) -> np.ndarray:
    """cv.getGaussianKernel(41, sigma, CV_32F): exp(-x^2/2s^2)/sum in f64, stored as f32."""
    x = np.arange(-20, 21, dtype=F64)
    k = np.exp(-(x * x) / (2.0 * sigma * sigma))
    return (k / k.sum()).astype(F32)


def _fma32(a, b, c):
    """float32 fused multiply-add emulated in float64 (exact product, one rounding...)."""
    # a*b is exact in f64 for f32 inputs (48-bit product); the f64 sum rounds once to 53 bits
    # and then to 24 -- a double rounding that can differ from a true fma only when the f64
    # sum lands within 2^-29 relative of an f32 tie; vanishingly rare and noted in tests.
    return (a.astype(F64) * b.astype(F64) + c.astype(F64)).astype(F32)


def normalize_minmax_f32(img: np.ndarray):
    """cv.normalize(img, None, 0, 1, NORM_MINMAX, CV_32F).  OpenCV computes
    scale = 1/(max-min) in f64, then -- for a CV_32F result -- rounds it to f32 *first* and forms
    shift = (float)0 - (float)(min*scale_f32); convertTo applies them as one f32 FMA per pixel."""
    smin, smax = float(img.min()), float(img.max())
    scale = 1.0 / (smax - smin) if smax - smin > 2.220446049250313e-16 else 0.0
    scale = float(F32(scale))
    shift = 0.0 - float(F32(smin * scale))
    return _fma32(img.astype(F32), np.full(1, scale, F32), np.full(1, shift, F32)), scale, shift


def sep_blur_41(f: np.ndarray, k: np.ndarray) -> np.ndarray:
    """cv.GaussianBlur(f32, (41,41), sigma) = row filter then column filter, REFLECT_101,
    as executed by OpenCV's AVX2-dispatched separable filter (filter.simd.hpp) on an
    AVX2+FMA host -- which is what both this container and the B200 hosts are:

    Row pass: s = 0; for j=0..40: s = fma(x[j-20], k[j], s) in the vector body (columns
    < W & ~3); the scalar tail columns use mul-then-add (no contraction).
    Column pass (symmetric): s = k[20]*c; for j=1..20: s = fma(x[+j] + x[-j], k[20+j], s)
    in the vector body (columns < W & ~7); tail columns mul-then-add.
    Pinned bit-exact vs cv2 4.13.0 for W >= 21 in tests/test_oracle_ops.py."""
    h, w = f.shape
    assert h >= 21 and w >= 21, "single-reflection REFLECT_101 only"
    xs = np.arange(w)
    acc = np.zeros_like(f)
    acc_t = np.zeros_like(f)
    for j in range(41):
        col = f[:, r101(xs + j - 20, w)]
        acc = _fma32(col, k[j:j + 1], acc)
        acc_t = acc_t + col * k[j]
    wv = (w // 4) * 4
    acc[:, wv:] = acc_t[:, wv:]
    ys = np.arange(h)
    out = acc * k[20]
    out_t = acc * k[20]
    for j in range(1, 21):
        pr = acc[r101(ys + j, h)] + acc[r101(ys - j, h)]
        out = _fma32(pr, k[20 + j:21 + j], out)
        out_t = out_t + pr * k[20 + j]
    wv = (w // 8) * 8
    out[:, wv:] = out_t[:, wv:]
    return out


def dog(img: np.ndarray, low_sigma: float = 5, high_sigma: float = 9) -> np.ndarray:
    """OptFlowRegistrator.dog(img, True) (optflow_registrator.py:249-274)."""
    if img.max() == 0:
        return img
    f, _, _ = normalize_minmax_f32(img)
    ls = sep_blur_41(f, gaussian_kernel_41(low_sigma))
    hs = sep_blur_41(f, gaussian_kernel_41(high_sigma))
    d = hs - ls
    dmin, dmax = float(d.min()), float(d.max())
    scale = 255.0 * (1.0 / (dmax - dmin) if dmax - dmin > 2.220446049250313e-16 else 0.0)
    shift = 0.0 - dmin * scale
    q = _fma32(d, np.full(1, scale, F32), np.full(1, shift, F32))
    return np.clip(np.rint(q), 0, 255).astype(np.uint8)


def normalize_minmax_u8(img: np.ndarray) -> np.ndarray:
    """cv.normalize(img, None, 0, 255, NORM_MINMAX, CV_8U) (shared_modules/utils.py:94): scale, shift in f64,
    applied as one f32 FMA, round half to even, saturate."""
    smin, smax = float(img.min()), float(img.max())
    scale = 255.0 * (1.0 / (smax - smin) if smax - smin > 2.220446049250313e-16 else 0.0)
    shift = 0.0 - smin * scale
    q = _fma32(img.astype(F32), np.full(1, scale, F32), np.full(1, shift, F32))
    return np.clip(np.rint(q), 0, 255).astype(np.uint8)


# ------------------------------------------------------------------------------- NMI
def nmi(a: np.ndarray, b: np.ndarray) -> float:
    """sklearn.metrics.normalized_mutual_info_score(a, b) (arithmetic mean normaliser, natural log)
    via the joint histogram of the two label arrays."""
    a = np.asarray(a).ravel()
    b = np.asarray(b).ravel()
    ua, ia = np.unique(a, return_inverse=True)
    ub, ib = np.unique(b, return_inverse=True)
    if len(ua) == len(ub) == 1 or len(ua) == len(ub) == 0:
        return 1.0
    J = np.zeros((len(ua), len(ub)), np.int64)
    np.add.at(J, (ia, ib), 1)
    return nmi_from_hist(J)


def nmi_from_hist(J: np.ndarray) -> float:
    J = J.astype(np.int64)
    J = J[J.sum(1) > 0][:, J.sum(0) > 0]
    if J.shape[0] == 1 and J.shape[1] == 1:
        return 1.0
    n = float(J.sum())
    pi = J.sum(1).astype(F64)
    pj = J.sum(0).astype(F64)
    nzx, nzy = np.nonzero(J)
    nz = J[nzx, nzy].astype(F64)
    log_c = np.log(nz)
    c_sum = nz / n
    outer = pi[nzx].astype(np.int64) * pj[nzy].astype(np.int64)
    log_outer = -np.log(outer.astype(F64)) + np.log(pi.sum()) + np.log(pj.sum())
    mi = c_sum * (log_c - np.log(n)) + c_sum * log_outer
    mi = np.where(np.abs(mi) < np.finfo(F64).eps, 0.0, mi)
    mi = float(np.clip(mi.sum(), 0.0, None))

    def ent(p):
        p = p[p > 0]
        if len(p) == 1:
            return 0.0
        s = p.sum()
        return float(-np.sum((p / s) * (np.log(p) - np.log(s))))

    if mi <= np.finfo(F64).eps:  # sklearn >= 1.x shortcut
        return 0.0
    ha, hb = ent(pi), ent(pj)
    norm = 0.5 * (ha + hb)
    return float(mi / max(norm, np.finfo(F64).eps))
