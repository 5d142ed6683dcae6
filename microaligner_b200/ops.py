"""Device operators: thin, allocation-aware wrappers of the C ABI on torch CUDA tensors.

torch is used for device memory, streams and (in parallel.py) the process group only; all
arithmetic happens in libmicroaligner_b200.so.  Images are 2-D uint8/uint16 tensors (uint16 is
carried as torch.uint16), flows are (H, W, 2) float32 tensors, all C-contiguous on one device."""
import ctypes
import os
import sys
import weakref
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import MA_F32, MA_U8, MA_U16, check, lib

_DTYPES = {torch.uint8: MA_U8, torch.uint16: MA_U16, torch.float32: MA_F32}


def _code(t: torch.Tensor) -> int:
    try:
        return _DTYPES[t.dtype]
    except KeyError:
        raise TypeError(f"unsupported image dtype {t.dtype}; expected uint8 or uint16") from None


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _req(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise MicroalignerDeviceError(f"{what} must live on a CUDA device (no CPU fallback)")
    if not t.is_contiguous():
        raise ValueError(f"{what} must be C-contiguous")


class MicroalignerDeviceError(RuntimeError):
    pass


def _bytes(n: int, device) -> torch.Tensor:
    return torch.empty(max(int(n), 16), dtype=torch.uint8, device=device)


# --------------------------------------------------------------------------------- host <-> device
_MIRRORS = {}   # id(host array) -> (weakref to the array, device tensor it mirrors)


def mirrored(arr):
    """The device tensor a read-only result array mirrors (see mirror()), or None."""
    m = _MIRRORS.get(id(arr))
    if m is not None and m[0]() is arr and not arr.flags.writeable:
        return m[1]
    return None


def to_device(arr, device=None) -> torch.Tensor:
    """numpy (uint8/uint16/float32) or torch tensor -> contiguous CUDA tensor.  A read-only array handed out
    by to_host(mirror=True) is recognised by identity and mapped back to its device copy without a transfer."""
    if isinstance(arr, torch.Tensor):
        t = arr if arr.is_cuda else arr.to(device or "cuda")
        return t.contiguous()
    m = mirrored(arr)
    if m is not None:
        return m
    a = np.ascontiguousarray(arr)
    if a.dtype not in (np.uint8, np.uint16, np.float32):
        raise TypeError(f"unsupported dtype {a.dtype}; expected uint8, uint16 or float32")
    return torch.from_numpy(a).to(device or "cuda", non_blocking=False)


def to_host(t: torch.Tensor, mirror: bool = False) -> np.ndarray:
    """Device tensor -> numpy array backed by page-locked memory (torch's caching host allocator recycles
    the pinned block once the array is garbage collected), one DMA transfer.

    mirror=True marks the array READ-ONLY and remembers the device tensor it was copied from for as long as
    the array lives: passing that very array back (e.g. register() -> Warper.flow) then costs no upload.
    Read-only is what makes this safe -- the host copy cannot silently diverge from the device copy; call
    .copy() to get a writeable, un-mirrored array."""
    host = torch.empty(t.shape, dtype=t.dtype, device="cpu", pin_memory=True)
    host.copy_(t, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    arr = host.numpy()
    if mirror:
        arr.flags.writeable = False
        key = id(arr)
        _MIRRORS[key] = (weakref.ref(arr, lambda _r, key=key: _MIRRORS.pop(key, None)), t)
    return arr


def to_device_rows(arr: np.ndarray, rows, device=None) -> torch.Tensor:
    """Full-shape device tensor holding only rows [rows[0], rows[1]) of the host array (the rest stays uninitialised):
    what one rank of a sharded run uploads over its own PCIe link."""
    a = np.ascontiguousarray(arr)
    if a.dtype not in (np.uint8, np.uint16, np.float32):
        raise TypeError(f"unsupported dtype {a.dtype}; expected uint8, uint16 or float32")
    src = torch.from_numpy(a)
    t = torch.empty(src.shape, dtype=src.dtype, device=device or "cuda")
    r0, r1 = int(rows[0]), int(rows[1])
    if r1 > r0:
        t[r0:r1].copy_(src[r0:r1], non_blocking=True)
    return t


def to_host_rows(t: torch.Tensor, rows) -> np.ndarray:
    """Rows [rows[0], rows[1]) of a device tensor as a page-locked numpy array (one DMA transfer)."""
    r0, r1 = int(rows[0]), int(rows[1])
    host = torch.empty((max(r1 - r0, 0),) + tuple(t.shape[1:]), dtype=t.dtype, device="cpu", pin_memory=True)
    if r1 > r0:
        host.copy_(t[r0:r1], non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return host.numpy()


# ---- caller-owned host memory: page-locking on reuse, asynchronous row uploads, streamed downloads -------------------
_SEEN = {}        # (address, nbytes) -> times a pageable host range was uploaded
_REGISTERED = []  # [start, end) byte intervals page-locked with cudaHostRegister (whole pages), kept sorted and disjoint
PIN_MIN_BYTES = 32 << 20
_PAGE = 4096


def _clear_cuda_error():
    """cudaHostRegister failures are not sticky, but they stay behind as the runtime's 'last error' and torch would report
    them at its next launch check; cudaGetLastError of the runtime library torch itself loaded clears them."""
    import ctypes
    try:
        with open("/proc/self/maps") as f:
            paths = sorted({ln.split()[-1] for ln in f if "libcudart" in ln})
    except OSError:
        paths = []
    for path in paths + ["libcudart.so.12"]:
        try:
            ctypes.CDLL(path).cudaGetLastError()
        except OSError:
            continue


def _unregister(intervals):
    for iv in intervals:
        if iv in _REGISTERED:
            _REGISTERED.remove(iv)
            try:
                torch.cuda.cudart().cudaHostUnregister(iv[0])
            except Exception:  # interpreter shutdown
                pass


def _root(arr: np.ndarray) -> np.ndarray:
    while isinstance(getattr(arr, "base", None), np.ndarray):
        arr = arr.base
    return arr


def _register(lo: int, hi: int, owner, touch=None) -> bool:
    """cudaHostRegister of the whole pages under the byte range [lo, hi); pages that are registered already are skipped.
    `owner`: the registration is dropped when this object is garbage collected."""
    lo, hi = lo // _PAGE * _PAGE, -(-hi // _PAGE) * _PAGE
    todo, cur = [], lo
    for a, b in sorted(_REGISTERED):
        if b <= cur or a >= hi:
            continue
        if a > cur:
            todo.append((cur, a))
        cur = max(cur, b)
    if cur < hi:
        todo.append((cur, hi))
    done = []
    for a, b in todo:
        rc = int(torch.cuda.cudart().cudaHostRegister(a, b - a, 0))
        if rc != 0:
            # Not page-lockable: copies from / to this range stay staged (correct, ~5x slower).  Seen on hosts whose memory
            # is hot-added on demand: cudaErrorOperatingSystem for ranges that are already populated, or larger than
            # ~16 GB, while fresh (untouched) ranges of a few GB lock fine -- so lock result blocks when they are created.
            _clear_cuda_error()
            sys.stderr.write(f"microaligner_b200: cudaHostRegister({b - a} bytes) failed with error {rc}; copies stay staged\n")
            _unregister(done)
            return False
        _REGISTERED.append((a, b))
        done.append((a, b))
    if done:
        try:
            weakref.finalize(owner, _unregister, done)
        except TypeError:
            pass
    return True


def pin_on_reuse(arr: np.ndarray) -> bool:
    """Page-lock the caller's WHOLE array with cudaHostRegister the SECOND time it is uploaded from: a one-off upload of
    pageable memory is cheaper staged (registration costs about as much as one staged copy), a repeated one -- the same
    reference image against many moving images, bench loops -- then runs at DMA line rate (55 vs 11 GB/s on the B200
    hosts).  Whole arrays only: CUDA rejects a copy whose host range straddles the edge of a registered region, so a
    partially locked array would break the caller's own copies of other slices of it.  The registration is dropped when
    the array is garbage collected."""
    root = _root(arr)
    if root.nbytes < PIN_MIN_BYTES:
        return False
    if arr.size and torch.from_numpy(arr.reshape(-1)[:1]).is_pinned():
        return True
    key = (root.ctypes.data, root.nbytes)
    n = _SEEN.get(key, 0) + 1
    if n < 0:
        return False
    _SEEN[key] = n
    if n < 2:
        return False
    ok = _register(key[0], key[0] + key[1], root)
    if not ok:
        _SEEN[key] = -(1 << 30)        # do not try again
    return ok


def pin_rows(arr: np.ndarray, rows=None) -> bool:
    """Page-lock rows [rows[0], rows[1]) of a C-contiguous host array now (idempotent; None = all rows): the blocks this
    package owns and recycles from call to call -- each rank's rows of the node-shared results of
    parallel.Comm.shared_host_empty, locked while they are still untouched -- and input arrays a caller prepares for
    repeated use.  Copies issued by this package stay inside the locked rows; a caller who hands slices that straddle
    their edge to CUDA directly must copy them first (np.array)."""
    r0, r1 = (0, arr.shape[0]) if rows is None else (int(rows[0]), int(rows[1]))
    if r1 <= r0 or not arr.flags.c_contiguous:
        return r1 <= r0
    row_bytes = arr.strides[0]
    base = arr.ctypes.data
    return _register(base + r0 * row_bytes, base + r1 * row_bytes, _root(arr))


def upload_rows_async(arr: np.ndarray, rows, device, stream: "torch.cuda.Stream"):
    """Full-shape device tensor whose rows [rows[0], rows[1]) are being copied from the host array on `stream`; returns
    (tensor, event recorded after the copy).  Rows outside the range stay uninitialised."""
    a = np.ascontiguousarray(arr)
    if a.dtype not in (np.uint8, np.uint16, np.float32):
        raise TypeError(f"unsupported dtype {a.dtype}; expected uint8, uint16 or float32")
    src = torch.from_numpy(a)
    t = torch.empty(src.shape, dtype=src.dtype, device=device or "cuda")
    r0, r1 = int(rows[0]), int(rows[1])
    ev = torch.cuda.Event()
    stream.wait_stream(torch.cuda.current_stream())      # the allocation above is ordered on the current stream
    with torch.cuda.stream(stream):
        if r1 > r0:
            pin_on_reuse(a)
            t[r0:r1].copy_(src[r0:r1], non_blocking=True)
        ev.record(stream)
    return t, ev


_COPY_STREAMS = {}


def upload_rows(arr: np.ndarray, rows, device=None):
    """upload_rows_async on this device's upload stream; returns (tensor, event) -- call wait_upload(event) on the
    stream that reads the tensor, as late as possible."""
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    st = _COPY_STREAMS.get(dev)
    if st is None:
        st = _COPY_STREAMS[dev] = torch.cuda.Stream(dev)
    return upload_rows_async(arr, rows, dev, st)


def wait_upload(event):
    if event is not None:
        torch.cuda.current_stream().wait_event(event)


def host_result(shape, dtype) -> np.ndarray:
    """Fresh page-locked host array for a result (torch's caching host allocator recycles the block once the array dies)."""
    tdt = dtype if isinstance(dtype, torch.dtype) else torch.from_numpy(np.empty(0, dtype)).dtype
    return torch.empty(tuple(shape), dtype=tdt, pin_memory=True).numpy()


class HostSink:
    """Streams rows of a device tensor into a host array on a side stream while the kernels that follow keep running
    (PCIe copies overlap compute).  push() orders the copy after everything enqueued so far on the current stream; a
    later push() of the same rows overwrites an earlier one (copies on the side stream run in order)."""

    def __init__(self, host: np.ndarray):
        self.array = host
        self._host = torch.from_numpy(host)
        self._stream = torch.cuda.Stream()
        self._keep = []

    def push(self, dev: torch.Tensor, rows):
        r0, r1 = int(rows[0]), int(rows[1])
        if r1 <= r0:
            return
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        self._stream.wait_event(ev)
        with torch.cuda.stream(self._stream):
            self._host[r0:r1].copy_(dev[r0:r1], non_blocking=True)
        self._keep.append(dev)          # the source must outlive the copy

    def wait(self):
        self._stream.synchronize()
        self._keep.clear()


def mirror(arr: np.ndarray, dev: torch.Tensor) -> np.ndarray:
    """Mark a result array READ-ONLY and remember the device tensor it was copied from for as long as the array lives
    (see to_host(mirror=True))."""
    arr.flags.writeable = False
    key = id(arr)
    _MIRRORS[key] = (weakref.ref(arr, lambda _r, key=key: _MIRRORS.pop(key, None)), dev)
    return arr


def pinned_like(arr: np.ndarray) -> np.ndarray:
    """Copy of a numpy array in page-locked host memory (fast H2D for the drop-in numpy API)."""
    t = torch.empty(arr.shape, dtype=torch.from_numpy(arr[:0]).dtype, pin_memory=True)
    out = t.numpy()
    out[...] = arr
    return out


# --------------------------------------------------------------------------------- pyramid
def pyr_down(img: torch.Tensor) -> torch.Tensor:
    _req(img, "image")
    h, w = img.shape
    out = torch.empty(((h + 1) // 2, (w + 1) // 2), dtype=img.dtype, device=img.device)
    es = img.element_size()
    check(lib.ma_pyrdown(img.data_ptr(), w * es, h, w, _code(img), out.data_ptr(), out.shape[1] * es, _stream()), "ma_pyrdown")
    return out


def pyr_down_rows(img: torch.Tensor, rows, out: torch.Tensor) -> torch.Tensor:
    """Rows [rows[0], rows[1]) of cv.pyrDown(img) written into `out` ((h+1)//2, (w+1)//2)."""
    h, w = img.shape
    es = img.element_size()
    check(lib.ma_pyrdown_rows(img.data_ptr(), w * es, h, w, _code(img), out.data_ptr(), out.shape[1] * es, int(rows[0]),
                              int(rows[1]), _stream()), "ma_pyrdown_rows")
    return out


def pyr_up_flow(flow: torch.Tensor, dsize_hw: Sequence[int], scale: float = 1.0) -> torch.Tensor:
    _req(flow, "flow")
    h, w, _ = flow.shape
    dh, dw = int(dsize_hw[0]), int(dsize_hw[1])
    out = torch.empty((dh, dw, 2), dtype=torch.float32, device=flow.device)
    check(lib.ma_pyrup_flow(flow.data_ptr(), h, w, out.data_ptr(), dh, dw, float(scale), _stream()), "ma_pyrup_flow")
    return out


# --------------------------------------------------------------------------------- warp / merge
def warp_tiles(img: torch.Tensor, flow: torch.Tensor, tile_size: int, overlap: int,
               out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _req(img, "image")
    _req(flow, "flow")
    h, w = img.shape
    if tuple(flow.shape) != (h, w, 2) or flow.dtype != torch.float32:
        raise ValueError(f"flow must be float32 of shape {(h, w, 2)}, got {flow.dtype} {tuple(flow.shape)}")
    if out is None:
        out = torch.empty_like(img)
    es = img.element_size()
    check(lib.ma_warp_tiles(img.data_ptr(), w * es, _code(img), flow.data_ptr(), h, w, int(tile_size), int(overlap),
                            out.data_ptr(), w * es, _stream()), "ma_warp_tiles")
    return out


def warp_tiles_host_streamed(image: np.ndarray, flow: torch.Tensor, tile_size: int, overlap: int, rows=None,
                            out: Optional[np.ndarray] = None) -> np.ndarray:
    """Warper.warp for a HOST image and a device-resident flow: the image goes up, is warped and comes back one
    tile row at a time on three streams, so the H2D and D2H transfers overlap (PCIe is full duplex) instead of
    running back to back.  Same kernel, same rows, same result as warp_tiles.
    rows = (r0, r1), a multiple-of-tile_size aligned band: only these output rows are produced (one rank of several);
    out = host array to fill (page-locked or registered for asynchronous copies), default a fresh page-locked one."""
    a = np.ascontiguousarray(image)
    if a.dtype not in (np.uint8, np.uint16):
        raise TypeError(f"unsupported image dtype {a.dtype}; expected uint8 or uint16")
    h, w = a.shape
    T, ov = int(tile_size), int(overlap)
    r0, r1 = (0, h) if rows is None else (int(rows[0]), int(rows[1]))
    dev = flow.device
    src = torch.from_numpy(a)
    img_d = torch.empty((h, w), dtype=src.dtype, device=dev)
    out_d = torch.empty_like(img_d)
    out_h = torch.empty((h, w), dtype=src.dtype, pin_memory=True) if out is None else torch.from_numpy(out)
    if r1 <= r0:
        return out_h.numpy()
    cur = torch.cuda.current_stream()
    up, comp, down = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
    for s in (up, comp, down):
        s.wait_stream(cur)          # the flow (and the allocations) were produced on the current stream
    uploaded = max(r0 - ov, 0)
    if rows is not None:
        pin_on_reuse(a)
    for y0 in range(r0, r1, T):
        y1 = min(y0 + T, r1)
        need = min(y1 + ov, h)      # a tile row reads image rows [y0 - ov, y1 + ov)
        if need > uploaded:
            with torch.cuda.stream(up):
                img_d[uploaded:need].copy_(src[uploaded:need], non_blocking=True)
            uploaded = need
        comp.wait_stream(up)
        with torch.cuda.stream(comp):
            warp_tiles_rows(img_d, flow, T, ov, (y0, y1), out_d)
        down.wait_stream(comp)
        with torch.cuda.stream(down):
            out_h[y0:y1].copy_(out_d[y0:y1], non_blocking=True)
    down.synchronize()
    cur.wait_stream(comp)
    return out_h.numpy()


def warp_affine(img: torch.Tensor, inv3x3: np.ndarray, out_shape: Sequence[int], pad_top: int = 0, pad_left: int = 0) -> torch.Tensor:
    """The page `img`, placed at (pad_top, pad_left) of a zero frame of out_shape, resampled through the 3x3
    output->input matrix `inv3x3` with skimage's order-1 / constant-0 / float64 arithmetic (ma_warp_affine)."""
    _req(img, "image")
    if img.dim() != 2:
        raise ValueError("image must be 2-D")
    m = np.ascontiguousarray(inv3x3, dtype=np.float64)
    if m.shape != (3, 3):
        raise ValueError("inv3x3 must be a 3x3 matrix")
    oh, ow = int(out_shape[0]), int(out_shape[1])
    out = torch.empty((oh, ow), dtype=img.dtype, device=img.device)
    es = img.element_size()
    check(lib.ma_warp_affine(img.data_ptr(), img.stride(0) * es, _code(img), img.shape[0], img.shape[1], int(pad_top), int(pad_left),
                             m.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), out.data_ptr(), ow * es, oh, ow, _stream()),
          "ma_warp_affine")
    return out


def merge_flows_tiles(f1: torch.Tensor, f2: torch.Tensor, tile_size: int, overlap: int) -> torch.Tensor:
    _req(f1, "flow1")
    _req(f2, "flow2")
    h, w, _ = f1.shape
    out = torch.empty_like(f1)
    ws = _bytes(lib.ma_merge_workspace_bytes(h, w, int(tile_size)), f1.device)
    check(lib.ma_merge_flows_tiles(f1.data_ptr(), f2.data_ptr(), h, w, int(tile_size), int(overlap), out.data_ptr(),
                                   ws.data_ptr(), _stream()), "ma_merge_flows_tiles")
    return out


# --------------------------------------------------------------------------------- Farneback
_FB_WS = {}
FARNEBACK_WORKSPACE_BUDGET = 24 << 30  # bytes of HBM the tile batch may use
# MA_FB_FULL_WINDOWS=1: no dependency-cone trimming of the iterations (A/B measurements; same results)
FB_FULL_WINDOWS = os.environ.get("MA_FB_FULL_WINDOWS", "0") not in ("", "0")

def _fb_workspace(device, nbytes: int) -> torch.Tensor:
    ws = _FB_WS.get(device)
    if ws is None or ws.numel() < nbytes:
        _FB_WS.pop(device, None)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
        _FB_WS[device] = ws
    return ws


def release_workspaces():
    _FB_WS.clear()


def n_tiles(h: int, w: int, tile_size: int) -> int:
    return (-(-h // tile_size)) * (-(-w // tile_size))


def farneback_tiles(mov: torch.Tensor, ref: torch.Tensor, tile_size: int, overlap: int, win: int, iters: int,
                    tile_range=None, out: Optional[torch.Tensor] = None, contract_fma: bool = False,
                    full_windows: Optional[bool] = None) -> torch.Tensor:
    """Stitched flow of the tiled (tile_size > 0) or untiled (tile_size <= 0) Farneback.
    contract_fma=True trades bit parity for speed in the window blur (MA_FB_CONTRACT_FMA).
    full_windows=True disables the dependency-cone trimming of the iterations (MA_FB_FULL_WINDOWS, same results)."""
    _req(mov, "moving image")
    _req(ref, "reference image")
    if mov.shape != ref.shape or mov.dtype != ref.dtype:
        raise ValueError("moving and reference image must have the same shape and dtype")
    h, w = ref.shape
    T = int(tile_size)
    ntot = n_tiles(h, w, T) if T > 0 else 1
    t0, t1 = tile_range if tile_range is not None else (0, ntot)
    if out is None:
        out = torch.zeros((h, w, 2), dtype=torch.float32, device=ref.device) if (t0, t1) != (0, ntot) else \
            torch.empty((h, w, 2), dtype=torch.float32, device=ref.device)
    slot = lib.ma_farneback_workspace_bytes(h, w, T, int(overlap), 1)
    nb = max(1, min(t1 - t0, FARNEBACK_WORKSPACE_BUDGET // slot))
    ws = _fb_workspace(ref.device, nb * slot)
    es = ref.element_size()
    check(lib.ma_farneback_tiles_ex(mov.data_ptr(), ref.data_ptr(), w * es, _code(ref), h, w, T, int(overlap), int(win),
                                    int(iters), int(t0), int(t1), out.data_ptr(), ws.data_ptr(), ws.numel(),
                                    (1 if contract_fma else 0) | (4 if (FB_FULL_WINDOWS if full_windows is None else full_windows) else 0),
                                    _stream()), "ma_farneback_tiles")
    return out


# --------------------------------------------------------------------------------- DoG / NMI
def dog_u8(img: torch.Tensor) -> torch.Tensor:
    _req(img, "image")
    h, w = img.shape
    out = torch.empty((h, w), dtype=torch.uint8, device=img.device)
    ws = _bytes(lib.ma_dog_workspace_bytes(h, w), img.device)
    check(lib.ma_dog_u8(img.data_ptr(), w * img.element_size(), _code(img), h, w, out.data_ptr(), w, ws.data_ptr(),
                        _stream()), "ma_dog_u8")
    return out


def nmi_chunks(a: torch.Tensor, b: torch.Tensor, chunk: int) -> torch.Tensor:
    """Per-chunk NMI (float64 device tensor) of two uint8 label images over raster chunks."""
    _req(a, "labels a")
    _req(b, "labels b")
    if a.dtype != torch.uint8 or b.dtype != torch.uint8 or a.numel() != b.numel():
        raise ValueError("NMI labels must be uint8 tensors of equal size")
    n = a.numel()
    chunk = int(min(chunk, n))
    nchunks = -(-n // chunk)
    scores = torch.empty(nchunks, dtype=torch.float64, device=a.device)
    check(lib.ma_nmi_chunks(a.data_ptr(), b.data_ptr(), n, chunk, scores.data_ptr(), None, _stream()),
          "ma_nmi_chunks")
    return scores


def minmax(img: torch.Tensor) -> torch.Tensor:
    _req(img, "image")
    h, w = img.shape
    out = torch.empty(2, dtype=torch.float32, device=img.device)
    check(lib.ma_minmax(img.data_ptr(), w * img.element_size(), _code(img), h, w, out.data_ptr(), _stream()), "ma_minmax")
    return out


def zmip_normalize_u8(pages: Sequence[torch.Tensor]) -> torch.Tensor:
    """z max-projection of same-shape pages followed by min-max normalisation to uint8."""
    for p in pages:
        _req(p, "page")
    h, w = pages[0].shape
    out = torch.empty((h, w), dtype=torch.uint8, device=pages[0].device)
    ws = _bytes(lib.ma_zmip_workspace_bytes(h, w, _code(pages[0])), pages[0].device)
    arr = (ctypes.c_void_p * len(pages))(*[p.data_ptr() for p in pages])
    check(lib.ma_zmip_normalize_u8(arr, len(pages), w * pages[0].element_size(), _code(pages[0]), h, w, out.data_ptr(), w,
                                   ws.data_ptr(), _stream()), "ma_zmip_normalize_u8")
    return out


# --------------------------------------------------------------------------------- row-range variants (row-sharded engine)
def warp_tiles_rows(img, flow, tile_size, overlap, rows, out):
    """Rows [rows[0], rows[1]) of Warper.warp written into `out` (full-size buffer)."""
    h, w = img.shape
    es = img.element_size()
    check(lib.ma_warp_tiles_rows(img.data_ptr(), w * es, _code(img), flow.data_ptr(), h, w, int(tile_size), int(overlap),
                                 out.data_ptr(), w * es, int(rows[0]), int(rows[1]), _stream()), "ma_warp_tiles_rows")
    return out


def pyr_up_flow_rows(flow, dsize_hw, scale, rows, out):
    h, w, _ = flow.shape
    check(lib.ma_pyrup_flow_rows(flow.data_ptr(), h, w, out.data_ptr(), int(dsize_hw[0]), int(dsize_hw[1]), float(scale),
                                 int(rows[0]), int(rows[1]), _stream()), "ma_pyrup_flow_rows")
    return out


def merge_flows_tile_rows(f1, f2, tile_size, overlap, tile_rows, out):
    h, w, _ = f1.shape
    ws = _bytes(lib.ma_merge_workspace_bytes(h, w, int(tile_size)), f1.device)
    check(lib.ma_merge_flows_tile_rows(f1.data_ptr(), f2.data_ptr(), h, w, int(tile_size), int(overlap), out.data_ptr(),
                                       ws.data_ptr(), int(tile_rows[0]), int(tile_rows[1]), _stream()), "ma_merge_flows_tile_rows")
    return out


def compose_flows_rows(f1, f2, rows, out):
    """Opt-in corrected composition: out(p) = f2(p) + bilinear(f1, p - f2(p)) for image rows [rows)."""
    h, w, _ = f1.shape
    check(lib.ma_compose_flows_rows(f1.data_ptr(), f2.data_ptr(), h, w, out.data_ptr(), int(rows[0]), int(rows[1]), _stream()),
          "ma_compose_flows_rows")
    return out


def minmax_rows(img, rows, out=None):
    """[min, max] (float32 device tensor) of image rows [rows[0], rows[1])."""
    h, w = img.shape
    if out is None:
        out = torch.empty(2, dtype=torch.float32, device=img.device)
    if rows[1] <= rows[0]:
        out[0], out[1] = float("inf"), float("-inf")
        return out
    es = img.element_size()
    check(lib.ma_minmax(img.data_ptr() + int(rows[0]) * w * es, w * es, _code(img), int(rows[1] - rows[0]), w, out.data_ptr(),
                        _stream()), "ma_minmax")
    return out


def dog_diff_rows(img, src_minmax, rows, dmm=None):
    """Rows [rows) of the un-normalised difference of Gaussians (band-sized plane: row 0 = image row rows[0])
    and their [min, max]; see ma_dog_diff_rows."""
    h, w = img.shape
    n = max(int(rows[1] - rows[0]), 0)
    wp = lib.ma_dog_diff_pitch_floats(w)
    diff = torch.empty((max(n, 1), wp), dtype=torch.float32, device=img.device)
    if dmm is None:
        dmm = torch.empty(2, dtype=torch.float32, device=img.device)
    ws = _bytes(lib.ma_dog_band_workspace_bytes(w, n), img.device)
    check(lib.ma_dog_diff_rows(img.data_ptr(), w * img.element_size(), _code(img), h, w, src_minmax.data_ptr(), int(rows[0]),
                               int(rows[0]) + n, diff.data_ptr(), dmm.data_ptr(), ws.data_ptr(), _stream()), "ma_dog_diff_rows")
    return diff, dmm


def dog_quantize_rows(diff, h, w, diff_minmax, rows, out):
    """uint8 rows [rows) from a band-sized diff plane produced by dog_diff_rows for the same rows."""
    if rows[1] > rows[0]:
        check(lib.ma_dog_quantize_rows(diff.data_ptr(), int(rows[0]), int(h), int(w), diff_minmax.data_ptr(), int(rows[0]),
                                       int(rows[1]), out.data_ptr(), int(w), _stream()), "ma_dog_quantize_rows")
    return out


def nmi_chunk_range2(a, b0, b1, chunk, chunk_range, scores0, scores1):
    """Both comparisons of the similarity gate in one launch: scores0 = NMI(a, b0), scores1 = NMI(a, b1) per chunk."""
    check(lib.ma_nmi_chunk_range2(a.data_ptr(), b0.data_ptr(), b1.data_ptr(), a.numel(), int(chunk), int(chunk_range[0]),
                                  int(chunk_range[1]), scores0.data_ptr(), scores1.data_ptr(), _stream()), "ma_nmi_chunk_range2")


def nmi_chunk_range(a, b, chunk, chunk_range, scores):
    n = a.numel()
    check(lib.ma_nmi_chunk_range(a.data_ptr(), b.data_ptr(), n, int(chunk), int(chunk_range[0]), int(chunk_range[1]),
                                 scores.data_ptr(), None, _stream()), "ma_nmi_chunk_range")
    return scores
