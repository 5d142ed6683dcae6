"""TEST INFRASTRUCTURE: a CPU stand-in for microaligner_b200.ops (the wrappers of the CUDA library) built on cv2 /
the oracle, with the same call signatures and the same *row-range* semantics, operating on CPU torch tensors.

It exists so that the sharding logic of microaligner_b200.engine -- which rows every rank computes, exchanges and
reads -- can be exercised without a GPU: tests/test_engine_sim_cpu.py runs Engine.register() on several gloo ranks
with this module patched in for `engine.ops` and requires the result to equal the single-rank one.  Every operator
computes ONLY the rows / tiles it is asked for and reads ONLY the rows its stencil reaches, and fresh buffers are
poisoned (tests patch torch.empty), so a rank that reads rows it neither owns nor fetched produces a wrong result.
The arithmetic is NOT that of the CUDA kernels (cv2 calls, not bit-identical restatements); only locality matters."""
import numpy as np
import torch

from oracle import cv_ops, reference_flow as rf

_BE = rf.CvBackend(1)


def _np(t: torch.Tensor) -> np.ndarray:
    return t.numpy()


def to_device(arr, device=None):
    return arr if isinstance(arr, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(arr))


def to_host(t, mirror=False):
    return t.numpy().copy()


_POISON = np.random.default_rng(99)


def to_device_rows(arr, rows, device=None):
    """Only rows [rows) of the host array arrive; everything else is garbage (as torch.empty would leave it)."""
    a = np.ascontiguousarray(arr)
    out = _POISON.integers(0, np.iinfo(a.dtype).max, a.shape, dtype=a.dtype, endpoint=True)
    out[int(rows[0]):int(rows[1])] = a[int(rows[0]):int(rows[1])]
    return torch.from_numpy(out)


def to_host_rows(t, rows):
    return t.numpy()[int(rows[0]):int(rows[1])].copy()


def upload_rows(arr, rows, device=None):
    return to_device_rows(arr, rows, device), None


def wait_upload(event):
    assert event is None


def pin_rows(arr, rows):
    return True


def host_result(shape, dtype):
    a = np.empty(tuple(shape), dtype)
    a.view(np.uint8)[...] = 0xAB        # poison: rows nobody delivers stay recognisable
    return a


class HostSink:
    """Synchronous stand-in for ops.HostSink; records which rows were delivered, in order."""

    def __init__(self, host):
        self.array, self.log = host, []

    def push(self, dev, rows):
        r0, r1 = int(rows[0]), int(rows[1])
        if r1 > r0:
            self.array[r0:r1] = dev.numpy()[r0:r1]
            self.log.append((r0, r1))

    def wait(self):
        pass


def n_tiles(h, w, T):
    return (-(-h // T)) * (-(-w // T))


# ------------------------------------------------------------------ pyramid
def pyr_down_rows(img, rows, out):
    """5-tap REFLECT_101 pyrDown: output rows [a, b) read input rows 2a-2 .. 2b+1 only."""
    a, b = int(rows[0]), int(rows[1])
    if b <= a:
        return out
    src = _np(img)
    h = src.shape[0]
    lo, hi = max(2 * a - 2, 0), min(2 * b + 2, h)
    ys = cv_ops.r101(np.arange(2 * a - 2, 2 * b + 2), h)           # the rows the stencil reaches, reflected
    assert ys.min() >= lo and ys.max() < hi
    import cv2
    # filter the needed rows only: build a padded strip whose border rows are the reflected ones
    strip = src[ys]
    full = cv2.pyrDown(strip, borderType=cv2.BORDER_REFLECT_101)
    _np(out)[a:b] = full[1:1 + (b - a), :out.shape[1]]
    return out


def pyr_down(img):
    h, w = img.shape
    out = torch.empty(((h + 1) // 2, (w + 1) // 2), dtype=img.dtype)
    return pyr_down_rows(img, (0, out.shape[0]), out)


def pyr_up_flow_rows(flow, dsize_hw, scale, rows, out):
    """cv.pyrUp(flow * scale): destination rows [a, b) read source rows a//2 - 1 .. (b+1)//2 + 1."""
    a, b = int(rows[0]), int(rows[1])
    if b <= a:
        return out
    import cv2
    src = _np(flow)
    h = src.shape[0]
    lo, hi = max(a // 2 - 2, 0), min((b + 1) // 2 + 2, h)
    part = np.ascontiguousarray(src[lo:hi]) * np.float32(scale)
    # rows near the cut of `part` are wrong (border handling) unless the cut is the true image border:
    # keep a 2-row margin, which [lo, hi) provides by construction
    up = cv2.pyrUp(part, dstsize=(int(dsize_hw[1]), 2 * part.shape[0] if hi < h or 2 * h == dsize_hw[0] else 2 * part.shape[0] - 1))
    _np(out)[a:b] = up[a - 2 * lo:b - 2 * lo]
    return out


# ------------------------------------------------------------------ warp / merge
def _tile_rows_of(rows, T, h):
    return range(int(rows[0]) // T, -(-int(rows[1]) // T))


def warp_tiles_rows(img, flow, tile_size, overlap, rows, out):
    a, b = int(rows[0]), int(rows[1])
    if b <= a:
        return out
    T, ov = int(tile_size), int(overlap)
    image, fl, dst = _np(img), _np(flow), _np(out)
    h, w = image.shape
    ny, nx = rf.tile_grid(h, w, T)
    for i in _tile_rows_of((a, b), T, h):
        for j in range(nx):
            it, ft = rf.gather_tile(image, i, j, T, ov), rf.gather_tile(fl, i, j, T, ov)
            S0, S1 = it.shape
            gx, gy = np.meshgrid(np.arange(S1, dtype=np.float32), np.arange(S0, dtype=np.float32))
            wt = _BE.remap(it, np.stack([gx - ft[..., 0], gy - ft[..., 1]], axis=-1))
            y0, y1 = max(i * T, a), min((i + 1) * T, b, h)
            x0, x1 = j * T, min((j + 1) * T, w)
            dst[y0:y1, x0:x1] = wt[ov + y0 - i * T:ov + y1 - i * T, ov:ov + x1 - x0]
    return out


def warp_tiles_host_streamed(image, flow, tile_size, overlap, rows=None, out=None):
    """Only rows [rows[0] - overlap, rows[1] + overlap) of the host image are uploaded; only rows [rows) come back."""
    h = image.shape[0]
    r0, r1 = (0, h) if rows is None else (int(rows[0]), int(rows[1]))
    res = np.empty_like(image) if out is None else out
    if r1 > r0:
        dev = to_device_rows(image, (max(r0 - int(overlap), 0), min(r1 + int(overlap), h)))
        got = torch.from_numpy(_POISON.integers(0, 200, image.shape).astype(image.dtype))
        warp_tiles_rows(dev, flow, tile_size, overlap, (r0, r1), got)
        res[r0:r1] = got.numpy()[r0:r1]
    return res


def warp_tiles(img, flow, tile_size, overlap, out=None):
    out = torch.empty_like(img) if out is None else out
    return warp_tiles_rows(img, flow, tile_size, overlap, (0, img.shape[0]), out)


def merge_flows_tile_rows(f1, f2, tile_size, overlap, tile_rows, out):
    T, ov = int(tile_size), int(overlap)
    a, b, dst = _np(f1), _np(f2), _np(out)
    h, w = a.shape[:2]
    ny, nx = rf.tile_grid(h, w, T)
    for i in range(int(tile_rows[0]), int(tile_rows[1])):
        for j in range(nx):
            m = rf.merge_two_flows(rf.gather_tile(a, i, j, T, ov), rf.gather_tile(b, i, j, T, ov), _BE)
            y1, x1 = min((i + 1) * T, h), min((j + 1) * T, w)
            dst[i * T:y1, j * T:x1] = m[ov:ov + y1 - i * T, ov:ov + x1 - j * T]
    return out


def compose_flows_rows(f1, f2, rows, out):
    a, b = int(rows[0]), int(rows[1])
    if b > a:
        _np(out)[a:b] = rf.compose_flows(_np(f1), _np(f2), _BE)[a:b]
    return out


# ------------------------------------------------------------------ Farneback
def farneback_tiles(mov, ref, tile_size, overlap, win, iters, tile_range=None, out=None, contract_fma=False, full_windows=None):
    m, r = _np(mov), _np(ref)
    h, w = r.shape
    T, ov = int(tile_size), int(overlap)
    if out is None:
        out = torch.zeros((h, w, 2), dtype=torch.float32)
    dst = _np(out)
    if T <= 0:
        dst[...] = _BE.farneback(m, r, win, iters)
        return out
    ny, nx = rf.tile_grid(h, w, T)
    t0, t1 = tile_range if tile_range is not None else (0, ny * nx)
    for t in range(int(t0), int(t1)):
        i, j = divmod(t, nx)
        f = _BE.farneback(rf.gather_tile(m, i, j, T, ov), rf.gather_tile(r, i, j, T, ov), win, iters)
        y1, x1 = min((i + 1) * T, h), min((j + 1) * T, w)
        dst[i * T:y1, j * T:x1] = f[ov:ov + y1 - i * T, ov:ov + x1 - j * T]
    return out


# ------------------------------------------------------------------ DoG / NMI
def minmax_rows(img, rows, out=None):
    if out is None:
        out = torch.empty(2, dtype=torch.float32)
    a, b = int(rows[0]), int(rows[1])
    if b <= a:
        out[0], out[1] = float("inf"), float("-inf")
    else:
        part = _np(img)[a:b]
        out[0], out[1] = float(part.min()), float(part.max())
    return out


def dog_diff_rows(img, src_minmax, rows, dmm=None):
    """Band-sized plane (row 0 = image row rows[0]) of blur9(f) - blur5(f), f = (img - min) / (max - min); reads image
    rows rows[0]-20 .. rows[1]+20 (REFLECT_101 at the image border only)."""
    import cv2
    a, b = int(rows[0]), int(rows[1])
    n = max(b - a, 0)
    src = _np(img)
    h, w = src.shape
    diff = torch.full((max(n, 1), w), float("nan"), dtype=torch.float32)
    if dmm is None:
        dmm = torch.empty(2, dtype=torch.float32)
    if n == 0:
        dmm[0], dmm[1] = float("inf"), float("-inf")
        return diff, dmm
    lo, hi = float(src_minmax[0]), float(src_minmax[1])
    ys = cv_ops.r101(np.arange(a - 20, b + 20), h)
    f = (src[ys].astype(np.float32) - np.float32(lo)) * np.float32(1.0 / (hi - lo) if hi > lo else 0.0)
    d = cv2.GaussianBlur(f, (41, 41), 9, borderType=cv2.BORDER_REFLECT_101) - cv2.GaussianBlur(f, (41, 41), 5, borderType=cv2.BORDER_REFLECT_101)
    d = d[20:20 + n]
    diff.numpy()[:n] = d
    dmm[0], dmm[1] = float(d.min()), float(d.max())
    return diff, dmm


def dog_quantize_rows(diff, h, w, diff_minmax, rows, out):
    a, b = int(rows[0]), int(rows[1])
    if b > a:
        lo, hi = float(diff_minmax[0]), float(diff_minmax[1])
        scale = 255.0 / (hi - lo) if hi > lo else 0.0
        q = np.rint((diff.numpy()[:b - a, :w].astype(np.float64) - lo) * scale)
        _np(out)[a:b] = np.clip(q, 0, 255).astype(np.uint8)
    return out


def nmi_chunk_range2(a, b0, b1, chunk, chunk_range, scores0, scores1):
    nmi_chunk_range(a, b0, chunk, chunk_range, scores0)
    nmi_chunk_range(a, b1, chunk, chunk_range, scores1)


def nmi_chunk_range(a, b, chunk, chunk_range, scores):
    fa, fb_ = _np(a).ravel(), _np(b).ravel()
    for c in range(int(chunk_range[0]), int(chunk_range[1])):
        scores[c] = cv_ops.nmi(fa[c * chunk:(c + 1) * chunk], fb_[c * chunk:(c + 1) * chunk])
    return scores


def nmi_chunks(a, b, chunk):
    n = a.numel()
    chunk = int(min(chunk, n))
    nchunks = -(-n // chunk)
    return nmi_chunk_range(a, b, chunk, (0, nchunks), torch.empty(nchunks, dtype=torch.float64))
