"""ORACLE (test infrastructure only -- never imported by the product path).

CPU restatement of the *control flow* of the reference's non-linear registration path:

  register()            optflow_reg/optflow_registrator.py:93-173
  _generate_img_pyr     optflow_reg/optflow_registrator.py:175-202
  _upscale_flow_...     optflow_reg/optflow_registrator.py:204-215
  merge_two_flows       optflow_reg/optflow_registrator.py:37-47, 217-240
  dog()                 optflow_reg/optflow_registrator.py:249-274
  TileFlowCalc          optflow_reg/flow_calc.py:59-98
  Warper.warp           optflow_reg/warper.py:37-76
  tiles / stitch        shared_modules/slicer.py:23-118, shared_modules/stitcher.py:25-118
  mi_tiled / gate       shared_modules/similarity_scoring.py:27-68

written independently (tile windows are gathered by index arithmetic, centres scattered
back) with a pluggable arithmetic backend:

  CvBackend  -- the same third-party calls the reference makes (cv2.*, sklearn NMI); this is
                what `bench.py --impl reference` and the `cpu_baseline` leg time, optionally
                with a thread pool over tiles (the reference's dask.delayed fan-out).
  NpBackend  -- the numpy restatements of oracle/cv_ops.py and oracle/farneback_np.py.

Pinned against the unmodified reference (imported through oracle/ref_shim.py in the build
container) by tests/test_oracle_flow.py and tests/golden/e2e_*.npz.
"""
from concurrent.futures import ThreadPoolExecutor
from math import log2

import numpy as np

from . import cv_ops, farneback_np


# ------------------------------------------------------------------------------- backends
class CvBackend:
    name = "cv2"

    def __init__(self, workers: int = 1):
        import cv2
        from sklearn.metrics import normalized_mutual_info_score
        self.cv = cv2
        self._nmi = normalized_mutual_info_score
        self.workers = workers

    def map(self, fn, items):
        if self.workers <= 1 or len(items) <= 1:
            return [fn(*it) for it in items]
        with ThreadPoolExecutor(self.workers) as ex:
            return list(ex.map(lambda it: fn(*it), items))

    def farneback(self, mov, ref, win, iters):
        return self.cv.calcOpticalFlowFarneback(mov, ref, None, pyr_scale=0.5, levels=0, winsize=win,
                                                iterations=iters, poly_n=1, poly_sigma=1.7,
                                                flags=self.cv.OPTFLOW_FARNEBACK_GAUSSIAN)

    def remap(self, src, mapxy):
        return self.cv.remap(src, mapxy, None, self.cv.INTER_LINEAR)

    def pyr_down(self, img):
        return self.cv.pyrDown(img)

    def pyr_up(self, flow, dsize_hw, scale):
        f = flow * np.float32(scale) if scale != 1 else flow
        return self.cv.pyrUp(f, dstsize=(dsize_hw[1], dsize_hw[0]))

    def dog(self, img):
        cv = self.cv
        if img.max() == 0:
            return img
        fimg = cv.normalize(img, None, 0, 1, cv.NORM_MINMAX, cv.CV_32F)
        ls = cv.GaussianBlur(fimg, (41, 41), sigmaX=5, dst=None, sigmaY=5)
        hs = cv.GaussianBlur(fimg, (41, 41), sigmaX=9, dst=None, sigmaY=9)
        d = hs - ls
        return cv.normalize(d, None, 0, 255, cv.NORM_MINMAX, cv.CV_8U)

    def nmi(self, a, b):
        return float(self._nmi(a, b))

    def normalize_u8(self, img):
        return self.cv.normalize(img, None, 0, 255, self.cv.NORM_MINMAX, self.cv.CV_8U)


class NpBackend:
    name = "numpy"
    workers = 1

    def map(self, fn, items):
        return [fn(*it) for it in items]

    def farneback(self, mov, ref, win, iters):
        return farneback_np.farneback(mov, ref, win, iters)

    def remap(self, src, mapxy):
        return cv_ops.remap_linear(src, mapxy)

    def pyr_down(self, img):
        return cv_ops.pyr_down(img)

    def pyr_up(self, flow, dsize_hw, scale):
        return cv_ops.pyr_up_f32c2(flow, dsize_hw, scale)

    def dog(self, img):
        return cv_ops.dog(img)

    def nmi(self, a, b):
        return cv_ops.nmi(a, b)

    def normalize_u8(self, img):
        return cv_ops.normalize_minmax_u8(img)


# ------------------------------------------------------------------------------- tiles
def tile_grid(h, w, T):
    return -(-h // T), -(-w // T)


def gather_tile(arr, i, j, T, ov):
    """Window [iT-ov,(i+1)T+ov) x [jT-ov,(j+1)T+ov), zero outside the image (slicer.py:23-66)."""
    S = T + 2 * ov
    h, w = arr.shape[:2]
    y0, x0 = i * T - ov, j * T - ov
    ya, yb = max(y0, 0), min(y0 + S, h)
    xa, xb = max(x0, 0), min(x0 + S, w)
    out = np.zeros((S, S) + arr.shape[2:], arr.dtype)
    out[ya - y0:yb - y0, xa - x0:xb - x0] = arr[ya:yb, xa:xb]
    return out


def split(arr, T, ov):
    ny, nx = tile_grid(arr.shape[0], arr.shape[1], T)
    return [gather_tile(arr, i, j, T, ov) for i in range(ny) for j in range(nx)]


def stitch(tiles, h, w, T, ov):
    """out[y,x] = tile(y//T, x//T)[ov + y%T, ov + x%T] (stitcher.py:72-118)."""
    ny, nx = tile_grid(h, w, T)
    out = np.zeros((h, w) + tiles[0].shape[2:], tiles[0].dtype)
    n = 0
    for i in range(ny):
        for j in range(nx):
            yb, xb = min((i + 1) * T, h), min((j + 1) * T, w)
            out[i * T:yb, j * T:xb] = tiles[n][ov:ov + yb - i * T, ov:ov + xb - j * T]
            n += 1
    return out


# ------------------------------------------------------------------------------- ops on tiles
def warp(image, flow, T=1000, ov=100, be=None):
    """Warper.warp (warper.py:37-76): per tile remap with map = tile-local grid - flow."""
    be = be or CvBackend()
    h, w = image.shape[:2]
    S = T + 2 * ov
    gx = np.arange(S)
    gy = np.arange(S).reshape(-1, 1)

    def one(img_t, flow_t):
        m = np.negative(flow_t)
        m[:, :, 0] += gx
        m[:, :, 1] += gy
        return be.remap(img_t, m)

    tiles = [one(a, b) for a, b in zip(split(image, T, ov), split(flow, T, ov))]  # sequential in the reference
    return stitch(tiles, h, w, T, ov)


def merge_two_flows(f1, f2, be):
    if f1.max() == 0:
        return f2
    if f2.max() == 0:
        return f1
    return f1 + be.remap(f2, -f1)


def merge_flows_tiled(f1, f2, T, ov, be):
    h, w = f1.shape[:2]
    tiles = be.map(lambda a, b: merge_two_flows(a, b, be), list(zip(split(f1, T, ov), split(f2, T, ov))))
    return stitch(tiles, h, w, T, ov)


def calc_flow(ref, mov, T, ov, win, iters, be):
    """TileFlowCalc.calc_flow (flow_calc.py:59-79): prev = moving, next = reference."""
    if max(ref.shape) / T < 2:
        return be.farneback(mov, ref, win, iters)
    h, w = ref.shape
    tiles = be.map(lambda m, r: be.farneback(m, r, win, iters), list(zip(split(mov, T, ov), split(ref, T, ov))))
    return stitch(tiles, h, w, T, ov)


def mi_tiled(a, b, T, be):
    if max(a.shape) / T < 2:
        return be.nmi(a.flatten(), b.flatten())
    fa, fb = a.flatten(), b.flatten()
    n = T * T
    items = [(fa[s:s + n], fb[s:s + n]) for s in range(0, fa.size, n)]
    return float(np.mean(be.map(be.nmi, items)))


def image_pyramid(arr, num_pyr_lvl, use_full_res_img, be):
    if num_pyr_lvl < 0:
        raise ValueError("Number of pyramid levels cannot be less than 0")
    if num_pyr_lvl == 0 and not use_full_res_img:
        raise ValueError("Number of pyramid levels is 0 and use_full_res_img is False. "
                         "Please change one of the parameters")
    pyr, factors = [], []
    cur = arr
    for lvl in range(num_pyr_lvl):
        factor = 2 ** (lvl + 1)
        if arr.shape[0] / factor < 100 or arr.shape[1] / factor < 100:
            break
        cur = be.pyr_down(cur)
        pyr.append(cur)
        factors.append(factor)
    pyr.reverse()
    factors.reverse()
    if use_full_res_img:
        pyr.append(arr)
        factors.append(1)
    return pyr, factors


def compose_flows(f1, f2, be):
    """Opt-in 'corrected' composition used by microaligner_b200 (NOT the reference): f2 + f1(p - f2) in image coordinates."""
    h, w = f1.shape[:2]
    m = np.negative(f2)
    m[:, :, 0] += np.arange(w)
    m[:, :, 1] += np.arange(h).reshape(-1, 1)
    return f2 + be.remap(f1, m)


def register(ref_img, mov_img, num_pyr_lvl=4, num_iterations=3, tile_size=1000, overlap=100,
             use_full_res_img=False, use_dog=False, be=None, log=None, force_decisions=None, corrected=False):
    """OptFlowRegistrator.register (optflow_registrator.py:93-173).  `log` (a list) receives one
    dict per level: factor, mi_after, mi_before, better.  `force_decisions` overrides the gate
    (test hook used to exercise the 'Worse' branches)."""
    be = be or CvBackend()
    T, ov = tile_size, overlap
    win = ov - (1 - ov % 2)
    full_hw = ref_img.shape

    def dog(img, use_it):
        return be.dog(img) if use_it else img

    def upscale_to_full(flow, factor):
        if abs(flow.shape[0] - full_hw[0]) <= 1:
            return flow
        out = flow
        n = int(log2(factor))
        for i in range(n):
            out = be.pyr_up(flow, full_hw, 2 if corrected else 1) if i == n - 1 else be.pyr_up(flow, (2 * flow.shape[0], 2 * flow.shape[1]), 1)
        return out

    ref_pyr, factors = image_pyramid(ref_img, num_pyr_lvl, use_full_res_img, be)
    mov_pyr, _ = image_pyramid(mov_img, num_pyr_lvl, use_full_res_img, be)
    num_lvl = len(factors)
    m_flow = None
    for lvl, factor in enumerate(factors):
        mov_l = mov_pyr[lvl]
        if lvl > 0:
            mov_l = warp(mov_l, m_flow, T, ov, be)
        this_flow = calc_flow(dog(ref_pyr[lvl], use_dog), dog(mov_l, use_dog), T, ov, win, num_iterations, be)
        mov_l = warp(mov_l, this_flow, T, ov, be)
        rd = dog(ref_pyr[lvl], True)
        after = mi_tiled(rd, dog(mov_l, True), T, be)
        before = mi_tiled(rd, dog(mov_pyr[lvl], True), T, be)
        better = bool(after > before)
        if force_decisions is not None:
            better = bool(force_decisions[lvl])
        if log is not None:
            log.append(dict(factor=factor, mi_after=after, mi_before=before, better=better))
        nxt = mov_pyr[lvl + 1].shape if lvl + 1 < num_lvl else None
        if better:
            if lvl == 0:
                m_flow = be.pyr_up(this_flow, nxt, 2) if num_lvl > 1 else upscale_to_full(this_flow, factor)
            elif lvl == num_lvl - 1:
                m_flow = compose_flows(m_flow, this_flow, be) if corrected else merge_flows_tiled(m_flow, this_flow, T, ov, be)
                if not use_full_res_img:
                    m_flow = upscale_to_full(m_flow, factor)
            else:
                merged = compose_flows(m_flow, this_flow, be) if corrected else merge_flows_tiled(m_flow, this_flow, T, ov, be)
                m_flow = be.pyr_up(merged, nxt, 2)
        else:
            if lvl == 0:
                shape = nxt if num_lvl > 1 else mov_img.shape
                m_flow = np.zeros(tuple(shape) + (2,), np.float32)
            elif lvl == num_lvl - 1:
                if not use_full_res_img:
                    m_flow = be.pyr_up(m_flow, mov_img.shape, 2)
            else:
                m_flow = be.pyr_up(m_flow, nxt, 4)
    if m_flow is None:
        raise UnboundLocalError("cannot access local variable 'm_flow' where it is not associated with a value")
    return m_flow


# ------------------------------------------------------------------------------- pipeline dispatch
def max_project(pages, be):
    """read_and_max_project_pages (shared_modules/utils.py:75-95)."""
    m = pages[0]
    for p in pages[1:]:
        m = np.maximum(m, p)
    return be.normalize_u8(m)


def register_cycles(dataset, ref_channel, be=None, **params):
    """register_and_save_ofreg_imgs (__main__.py:320-437) on an in-memory dataset
    {cycle: {channel: {z: page}}}; returns {(cycle, channel, z): image}."""
    be = be or CvBackend()
    T, ov = params.get("tile_size", 1000), params.get("overlap", 100)
    out, ref_img = {}, None
    for cyc_id, cyc in enumerate(dataset):
        mip = max_project(list(dataset[cyc][ref_channel].values()), be)
        if cyc_id == 0:
            ref_img = mip
            for ch, pages in dataset[cyc].items():
                for z, page in pages.items():
                    out[(cyc, ch, z)] = page
            continue
        flow = register(ref_img, mip, be=be, **params)
        ref_img = warp(mip, flow, T, ov, be)
        for ch, pages in dataset[cyc].items():
            for z, page in pages.items():
                out[(cyc, ch, z)] = warp(page, flow, T, ov, be)
    return out
