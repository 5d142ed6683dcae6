"""Minimal uncompressed (Big)TIFF page I/O for the registration pipeline.

The reference reads pages with ``tifffile.TiffFile(p).series[0].pages[i].asarray()``
(shared_modules/utils.py:69-72) and writes the registered stack into ``tifffile.memmap(path, shape=(1, C, Z, Y, X),
dtype, bigtiff=True, description=ome_xml, contiguous=True)`` (__main__.py:116-132).  tifffile is not part of this
environment, and at GPU speed the pipeline is I/O bound, so this module provides exactly that surface, pure
Python + numpy, for the only pixel layout the pipeline produces and that a B200 can be fed at line rate:

* classic TIFF ("II*\\0" / "MM\\0*") and BigTIFF ("II+\\0"), any byte order;
* single-sample, unsigned 8 / 16 bit (or 32-bit float) pages in strips or tiles; uncompressed, deflate or LZW
  (decoded on the host; horizontal predictor supported) -- the fast path is the uncompressed one;
* uncompressed pages whose strips are contiguous in the file are returned as zero-copy ``np.memmap`` views
  (``TiffPage.asarray()``), or read straight into a caller-supplied, page-locked array (``read_into``) so the next
  H2D copy is a single DMA transfer;
* ``memmap(path, shape, dtype, description)`` creates a contiguous multi-page BigTIFF and returns the pixel block
  as one writeable ``np.memmap`` of the requested shape -- the reference's output idiom.

Anything else (JPEG / other codecs, multiple samples, sub-IFDs) raises ``TiffFormatError`` rather than guessing."""
import mmap
import os
import struct
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

# tag ids
IMAGE_WIDTH, IMAGE_LENGTH, BITS_PER_SAMPLE, COMPRESSION, PHOTOMETRIC = 256, 257, 258, 259, 262
IMAGE_DESCRIPTION, STRIP_OFFSETS, SAMPLES_PER_PIXEL, ROWS_PER_STRIP, STRIP_BYTE_COUNTS = 270, 273, 277, 278, 279
PLANAR_CONFIG, SOFTWARE, SAMPLE_FORMAT, TILE_WIDTH = 284, 305, 339, 322
PREDICTOR, TILE_LENGTH, TILE_OFFSETS, TILE_BYTE_COUNTS = 317, 323, 324, 325

_TYPE_SIZES = {1: 1, 2: 1, 3: 2, 4: 4, 5: 8, 6: 1, 7: 1, 8: 2, 9: 4, 10: 8, 11: 4, 12: 8, 16: 8, 17: 8, 18: 8}
_TYPE_FMT = {1: "B", 2: "c", 3: "H", 4: "I", 6: "b", 7: "B", 8: "h", 9: "i", 11: "f", 12: "d", 16: "Q", 17: "q", 18: "Q"}


class TiffFormatError(ValueError):
    pass


_POOL = None
_FDS: Dict[str, int] = {}


def write_page(mm: np.memmap, index: tuple, data: np.ndarray):
    """mm[index] = data for one page of a memory-mapped stack, through pwrite() on the file behind the mapping: filling
    fresh pages of a tmpfs / page-cache file through the mapping costs a page fault per 4 KB (1.3 GB/s measured), the
    system call fills them in bulk (2 GB/s); the mapping sees the same page cache."""
    view = mm[index]
    if not isinstance(mm, np.memmap) or mm.filename is None or not view.flags.c_contiguous or not data.flags.c_contiguous \
            or view.shape != data.shape or view.dtype != data.dtype:
        view[...] = data
        return
    path = os.fspath(mm.filename)
    fd = _FDS.get(path)
    if fd is None:
        fd = _FDS[path] = os.open(path, os.O_RDWR)
    off = mm.offset + (view.ctypes.data - mm.ctypes.data)
    buf = memoryview(data.reshape(-1).view(np.uint8))
    done = 0
    while done < len(buf):
        done += os.pwrite(fd, buf[done:done + (1 << 30)], off + done)


def close_writers():
    for fd in _FDS.values():
        os.close(fd)
    _FDS.clear()


def copy_rows(dst: np.ndarray, src: np.ndarray, threads: int = 8):
    """dst[...] = src for large 2-D arrays, rows split over a few threads: numpy's copy releases the GIL, and both the
    page-cache reads behind a memory-mapped TIFF page and the first-touch page faults of a fresh output mapping are
    per-thread costs (one thread moves ~1-2 GB/s through them, eight ~10 GB/s)."""
    global _POOL
    n = dst.shape[0]
    if dst.nbytes < (64 << 20) or threads <= 1 or n < threads:
        dst[...] = src
        return
    if _POOL is None:
        from concurrent.futures import ThreadPoolExecutor
        _POOL = ThreadPoolExecutor(max_workers=threads)
    step = -(-n // threads)

    def part(i):
        dst[i:i + step] = src[i:i + step]

    list(_POOL.map(part, range(0, n, step)))


def _lzw_decode(data: bytes, expected: int) -> bytes:
    """TIFF LZW (MSB-first codes, 9..12 bits, ClearCode 256, EOI 257, 'early change'): the native decoder of the
    shared library (ma_tiff_lzw_decode, host code) when it is built, else the same algorithm in Python (~1 MB/s)."""
    try:
        from ._lib import lib
    except ImportError:
        return _lzw_decode_py(data, expected)
    out = np.empty(max(expected, 1), np.uint8)
    n = lib.ma_tiff_lzw_decode(bytes(data), len(data), out.ctypes.data, expected)
    if n < 0:
        raise TiffFormatError("corrupt LZW stream")
    return out[:n].tobytes()


def _lzw_decode_py(data: bytes, expected: int) -> bytes:
    out = bytearray()
    table = [bytes([i]) for i in range(256)] + [b"", b""]
    bits, nbits, pos, n = 0, 0, 0, len(data)
    width, prev = 9, None
    while True:
        while nbits < width and pos < n:
            bits = (bits << 8) | data[pos]
            pos += 1
            nbits += 8
        if nbits < width:
            break
        code = (bits >> (nbits - width)) & ((1 << width) - 1)
        nbits -= width
        if code == 257:
            break
        if code == 256:
            table = table[:258]
            width, prev = 9, None
            continue
        if prev is None:
            entry = table[code]
        else:
            entry = table[code] if code < len(table) else prev + prev[:1]
            table.append(prev + entry[:1])
        out += entry
        prev = entry
        if len(table) >= (1 << width) - 1 and width < 12:
            width += 1
        if len(out) >= expected:
            break
    return bytes(out[:expected])


class TiffPage:
    def __init__(self, tif: "TiffFile", tags: Dict[int, tuple]):
        self._tif = tif
        self.tags = tags
        self.width = int(tags[IMAGE_WIDTH][0])
        self.height = int(tags[IMAGE_LENGTH][0])
        bits = int(tags.get(BITS_PER_SAMPLE, (1,))[0])
        spp = int(tags.get(SAMPLES_PER_PIXEL, (1,))[0])
        fmt = int(tags.get(SAMPLE_FORMAT, (1,))[0])
        self.compression = int(tags.get(COMPRESSION, (1,))[0])
        self.predictor = int(tags.get(PREDICTOR, (1,))[0])
        if self.compression not in (1, 5, 8, 32946):
            raise TiffFormatError(f"compressed TIFF (Compression={self.compression}) is not supported "
                                  "(none, LZW and deflate are)")
        if self.predictor not in (1, 2):
            raise TiffFormatError(f"Predictor={self.predictor} is not supported")
        if spp != 1:
            raise TiffFormatError(f"SamplesPerPixel={spp}: only single-channel pages are supported")
        kinds = {(8, 1): "u1", (16, 1): "u2", (32, 1): "u4", (32, 3): "f4", (16, 2): "i2", (8, 2): "i1"}
        if (bits, fmt) not in kinds:
            raise TiffFormatError(f"unsupported sample type: {bits} bits, SampleFormat {fmt}")
        self.dtype = np.dtype(tif.byteorder + kinds[(bits, fmt)]) if bits > 8 else np.dtype(kinds[(bits, fmt)])
        self.shape = (self.height, self.width)
        self.nbytes = self.height * self.width * self.dtype.itemsize
        # segments: (file offset, byte count, y0, x0, rows, cols) -- strips or tiles
        self.segments = []
        if TILE_WIDTH in tags:
            tw, tl = int(tags[TILE_WIDTH][0]), int(tags[TILE_LENGTH][0])
            offs, cnts = tags[TILE_OFFSETS], tags[TILE_BYTE_COUNTS]
            across = -(-self.width // tw)
            for i, (o, c) in enumerate(zip(offs, cnts)):
                self.segments.append((int(o), int(c), (i // across) * tl, (i % across) * tw, tl, tw))
            self.tiled = True
        else:
            rps = int(tags.get(ROWS_PER_STRIP, (self.height,))[0])
            rps = min(rps, self.height) or self.height
            for i, (o, c) in enumerate(zip(tags[STRIP_OFFSETS], tags[STRIP_BYTE_COUNTS])):
                y0 = i * rps
                self.segments.append((int(o), int(c), y0, 0, min(rps, self.height - y0), self.width))
            self.tiled = False
        self.offsets = [sg[0] for sg in self.segments]
        self.counts = [sg[1] for sg in self.segments]
        plain = self.compression == 1 and not self.tiled
        if plain and sum(self.counts) != self.nbytes:
            raise TiffFormatError("strip byte counts do not add up to an uncompressed page")
        self.is_contiguous = plain and all(self.offsets[i] + self.counts[i] == self.offsets[i + 1]
                                           for i in range(len(self.offsets) - 1))

    @property
    def description(self) -> Optional[str]:
        v = self.tags.get(IMAGE_DESCRIPTION)
        return v[0] if v else None

    def asarray(self) -> np.ndarray:
        """The page as a (height, width) array in native byte order; a zero-copy memmap view for uncompressed pages
        whose strips are contiguous in the file."""
        if self.is_contiguous:
            # np.memmap owns its own mapping of the file, so the array stays valid after TiffFile.close() -- the
            # reference returns the page after its with-block (shared_modules/utils.py:69-72)
            a = np.memmap(self._tif.path, dtype=self.dtype, mode="r", offset=self.offsets[0], shape=self.shape)
            return a if a.dtype.isnative else a.astype(a.dtype.newbyteorder("="))
        return self.read_into(np.empty(self.shape, self.dtype.newbyteorder("=")))

    def read_into(self, out: np.ndarray) -> np.ndarray:
        """Fill a caller-owned (e.g. page-locked) array of the page's shape and dtype."""
        if out.shape != self.shape or out.dtype.itemsize != self.dtype.itemsize or not out.flags.c_contiguous:
            raise ValueError(f"read_into needs a C-contiguous array of shape {self.shape} and item size {self.dtype.itemsize}")
        if self.compression == 1 and not self.tiled:
            if self.is_contiguous:      # one run of bytes in the file: a zero-copy view of the mapping, copied once
                src = np.frombuffer(self._tif._map, dtype=np.uint8, count=self.nbytes, offset=self.offsets[0])
                copy_rows(out.view(np.uint8).reshape(self.height, -1), src.reshape(self.height, -1))
            else:
                dst = memoryview(out.reshape(-1).view(np.uint8))
                src = memoryview(self._tif._map)
                pos = 0
                for off, n in zip(self.offsets, self.counts):
                    dst[pos:pos + n] = src[off:off + n]
                    pos += n
                del src
            if not self.dtype.isnative:
                out.byteswap(inplace=True)
            return out
        for off, n, y0, x0, rows, cols in self.segments:
            raw = self._tif._map[off:off + n]
            want = rows * cols * self.dtype.itemsize
            if self.compression in (8, 32946):
                import zlib
                raw = zlib.decompress(raw)
            elif self.compression == 5:
                raw = _lzw_decode(bytes(raw), want)
            if len(raw) < want:
                raise TiffFormatError("truncated TIFF segment")
            seg = np.frombuffer(raw, self.dtype, count=rows * cols).reshape(rows, cols)
            if not seg.dtype.isnative:
                seg = seg.astype(seg.dtype.newbyteorder("="))
            if self.predictor == 2:          # horizontal differencing, wraps modulo 2^bits
                seg = np.cumsum(seg, axis=1, dtype=seg.dtype)
            y1, x1 = min(y0 + rows, self.height), min(x0 + cols, self.width)
            out[y0:y1, x0:x1] = seg[:y1 - y0, :x1 - x0].view(out.dtype)
        return out


class TiffSeries:
    def __init__(self, pages: List[TiffPage], shape: Tuple[int, ...], axes: str):
        self.pages, self.shape, self.axes = pages, shape, axes
        self.dtype = pages[0].dtype.newbyteorder("=") if pages else None


class TiffFile:
    """``with TiffFile(path) as tif: tif.series[0].pages[i].asarray()`` -- the reference's read idiom."""

    def __init__(self, path):
        self.path = os.fspath(path)
        self._fh = open(self.path, "rb")
        self._map = mmap.mmap(self._fh.fileno(), 0, access=mmap.ACCESS_READ)
        head = self._map[:16]
        if head[:2] == b"II":
            self.byteorder = "<"
        elif head[:2] == b"MM":
            self.byteorder = ">"
        else:
            raise TiffFormatError("not a TIFF file")
        magic = struct.unpack(self.byteorder + "H", head[2:4])[0]
        if magic == 42:
            self.bigtiff = False
            first = struct.unpack(self.byteorder + "I", head[4:8])[0]
        elif magic == 43:
            self.bigtiff = True
            if struct.unpack(self.byteorder + "HH", head[4:8]) != (8, 0):
                raise TiffFormatError("malformed BigTIFF header")
            first = struct.unpack(self.byteorder + "Q", head[8:16])[0]
        else:
            raise TiffFormatError(f"unknown TIFF magic {magic}")
        self.pages = self._read_ifds(first)
        self.series = [self._make_series()]

    # -- parsing --------------------------------------------------------------------------------
    def _read_ifds(self, offset: int) -> List[TiffPage]:
        bo, big = self.byteorder, self.bigtiff
        pages, seen = [], set()
        while offset:
            if offset in seen:
                raise TiffFormatError("IFD loop")
            seen.add(offset)
            if big:
                n = struct.unpack(bo + "Q", self._map[offset:offset + 8])[0]
                pos, esz = offset + 8, 20
            else:
                n = struct.unpack(bo + "H", self._map[offset:offset + 2])[0]
                pos, esz = offset + 2, 12
            tags = {}
            for i in range(n):
                e = self._map[pos + i * esz:pos + (i + 1) * esz]
                tag, typ = struct.unpack(bo + "HH", e[:4])
                cnt = struct.unpack(bo + ("Q" if big else "I"), e[4:12] if big else e[4:8])[0]
                val = e[12:20] if big else e[8:12]
                size = _TYPE_SIZES.get(typ)
                if size is None:
                    continue
                nbytes = size * cnt
                if nbytes > len(val):
                    ptr = struct.unpack(bo + ("Q" if big else "I"), val)[0]
                    raw = self._map[ptr:ptr + nbytes]
                else:
                    raw = val[:nbytes]
                if typ == 2:
                    tags[tag] = (raw.rstrip(b"\x00").decode("utf-8", "replace"),)
                elif typ in (5, 10):
                    f = "I" if typ == 5 else "i"
                    v = struct.unpack(bo + f * (2 * cnt), raw)
                    tags[tag] = tuple(v[2 * k] / v[2 * k + 1] if v[2 * k + 1] else 0.0 for k in range(cnt))
                else:
                    tags[tag] = struct.unpack(bo + _TYPE_FMT[typ] * cnt, raw)
            nxt = pos + n * esz
            offset = struct.unpack(bo + ("Q" if big else "I"), self._map[nxt:nxt + (8 if big else 4)])[0]
            pages.append(TiffPage(self, tags))
        return pages

    def _make_series(self) -> TiffSeries:
        p0 = self.pages[0]
        same = all(p.shape == p0.shape and p.dtype == p0.dtype for p in self.pages)
        shape = ((len(self.pages),) + p0.shape) if (same and len(self.pages) > 1) else p0.shape
        desc = self.ome_metadata
        axes = "YX" if len(shape) == 2 else "IYX"
        if desc and same:
            dims = _ome_dims(desc)
            if dims and dims[0] * dims[1] * dims[2] == len(self.pages):
                t, c, z = dims
                shape, axes = tuple(v for v in (t, c, z) if True) + p0.shape, "TCZYX"
                # the reference requires 4-D CZYX inputs (img_checks.py:50-65): drop a singleton T
                if t == 1:
                    shape, axes = shape[1:], "CZYX"
        return TiffSeries(self.pages, shape, axes)

    @property
    def ome_metadata(self) -> Optional[str]:
        d = self.pages[0].description if self.pages else None
        return d if d and "OME" in d[:400] else None

    def asarray(self) -> np.ndarray:
        s = self.series[0]
        return np.stack([p.asarray() for p in s.pages]).reshape(s.shape) if len(s.pages) > 1 else s.pages[0].asarray()

    def close(self):
        try:
            self._map.close()
        finally:
            self._fh.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def _ome_dims(xml: str) -> Optional[Tuple[int, int, int]]:
    """(SizeT, SizeC, SizeZ) of the first Pixels element of an OME-XML string, if present."""
    import re
    m = re.search(r"<Pixels\b[^>]*>", xml)
    if not m:
        return None
    out = []
    for k in ("SizeT", "SizeC", "SizeZ"):
        mm = re.search(k + r'="(\d+)"', m.group(0))
        if not mm:
            return None
        out.append(int(mm.group(1)))
    return tuple(out)


# ------------------------------------------------------------------------------------- writing
def _ifd_entry(tag, typ, count, value) -> bytes:
    """BigTIFF IFD entry with an inline (<= 8 byte) value or an offset."""
    return struct.pack("<HHQ", tag, typ, count) + value.ljust(8, b"\x00")


def memmap(path, shape: Sequence[int], dtype, description: Optional[str] = None, software: str = "microaligner_b200"):
    """Create a little-endian BigTIFF holding prod(shape[:-2]) contiguous uncompressed pages of shape[-2:] and return
    the pixel block as a writeable np.memmap of `shape` (tifffile.memmap(..., bigtiff=True, contiguous=True))."""
    dtype = np.dtype(dtype).newbyteorder("<")
    if dtype.kind not in "uf" or dtype.itemsize not in (1, 2, 4):
        raise TiffFormatError(f"unsupported dtype {dtype}")
    shape = tuple(int(s) for s in shape)
    if len(shape) < 2:
        raise ValueError("shape needs at least (Y, X)")
    h, w = shape[-2:]
    npages = int(np.prod(shape[:-2])) if len(shape) > 2 else 1
    page_bytes = h * w * dtype.itemsize
    desc = (description or "").encode("utf-8") + b"\x00"
    soft = software.encode() + b"\x00"
    # layout: header (16) | description | software | IFDs | pad to 4096 | pixel block
    pos = 16
    desc_off, pos = pos, pos + len(desc)
    soft_off, pos = pos, pos + len(soft)
    pos = (pos + 7) // 8 * 8
    n_tags = 12
    ifd_size = 8 + n_tags * 20 + 8
    ifd0 = pos
    data_off = (ifd0 + npages * ifd_size + 4095) // 4096 * 4096
    with open(os.fspath(path), "wb") as f:
        f.write(struct.pack("<2sHHHQ", b"II", 43, 8, 0, ifd0))
        f.write(desc)
        f.write(soft)
        f.write(b"\x00" * (ifd0 - f.tell()))
        sample_format = 3 if dtype.kind == "f" else 1
        for i in range(npages):
            entries = [
                _ifd_entry(IMAGE_WIDTH, 4, 1, struct.pack("<I", w)),
                _ifd_entry(IMAGE_LENGTH, 4, 1, struct.pack("<I", h)),
                _ifd_entry(BITS_PER_SAMPLE, 3, 1, struct.pack("<H", dtype.itemsize * 8)),
                _ifd_entry(COMPRESSION, 3, 1, struct.pack("<H", 1)),
                _ifd_entry(PHOTOMETRIC, 3, 1, struct.pack("<H", 1)),
                _ifd_entry(IMAGE_DESCRIPTION, 2, len(desc), struct.pack("<Q", desc_off) if len(desc) > 8 else desc),
                _ifd_entry(STRIP_OFFSETS, 16, 1, struct.pack("<Q", data_off + i * page_bytes)),
                _ifd_entry(SAMPLES_PER_PIXEL, 3, 1, struct.pack("<H", 1)),
                _ifd_entry(ROWS_PER_STRIP, 4, 1, struct.pack("<I", h)),
                _ifd_entry(STRIP_BYTE_COUNTS, 16, 1, struct.pack("<Q", page_bytes)),
                _ifd_entry(SOFTWARE, 2, len(soft), struct.pack("<Q", soft_off) if len(soft) > 8 else soft),
                _ifd_entry(SAMPLE_FORMAT, 3, 1, struct.pack("<H", sample_format)),
            ]
            assert len(entries) == n_tags
            nxt = ifd0 + (i + 1) * ifd_size if i + 1 < npages else 0
            f.write(struct.pack("<Q", n_tags) + b"".join(entries) + struct.pack("<Q", nxt))
        f.write(b"\x00" * (data_off - f.tell()))
        f.truncate(data_off + npages * page_bytes)
    return np.memmap(os.fspath(path), dtype=dtype, mode="r+", offset=data_off, shape=shape)


def memmap_existing(path, shape: Sequence[int], dtype) -> np.memmap:
    """Map the pixel block of a file written by memmap() (contiguous, uncompressed pages) for reading and writing --
    how the other ranks of a multi-GPU run open the output file rank 0 created."""
    shape = tuple(int(s) for s in shape)
    with TiffFile(path) as tf:
        pages = tf.pages
        n = int(np.prod(shape[:-2])) if len(shape) > 2 else 1
        if len(pages) != n or pages[0].shape != shape[-2:] or not all(p.is_contiguous for p in pages):
            raise TiffFormatError(f"{path} does not hold {n} contiguous pages of shape {shape[-2:]}")
        first, nbytes = pages[0].offsets[0], pages[0].nbytes
        if any(p.offsets[0] != first + i * nbytes for i, p in enumerate(pages)):
            raise TiffFormatError(f"{path}: pages are not back to back")
        if pages[0].dtype.newbyteorder("=") != np.dtype(dtype).newbyteorder("="):
            raise TiffFormatError(f"{path}: dtype {pages[0].dtype} does not match {np.dtype(dtype)}")
        dt = pages[0].dtype
    return np.memmap(os.fspath(path), dtype=dt, mode="r+", offset=first, shape=shape)


def imwrite(path, data: np.ndarray, description: Optional[str] = None):
    mm = memmap(path, data.shape, data.dtype, description)
    mm[...] = data
    mm.flush()
    del mm


# ------------------------------------------------------------------------------------- pipeline glue
class TiffPageProvider:
    """cycle -> channel -> z -> callable returning the page in a page-locked staging buffer, for
    pipeline.register_and_save_ofreg_imgs.  `layout[cycle][channel][z] = (path, page index)`."""

    def __init__(self, layout: Dict[int, Dict[str, Dict[int, Tuple[str, int]]]], pinned: bool = True, n_buffers: int = 3):
        self.layout = layout
        self._files: Dict[str, TiffFile] = {}
        self._pinned = pinned
        self._bufs: List[np.ndarray] = []
        self._next = 0
        self._n = n_buffers

    def _file(self, path) -> TiffFile:
        if path not in self._files:
            self._files[path] = TiffFile(path)
        return self._files[path]

    def _staging(self, shape, dtype) -> np.ndarray:
        if not self._pinned:
            return np.empty(shape, dtype)
        if len(self._bufs) < self._n or self._bufs[0].shape != shape or self._bufs[0].dtype != dtype:
            import torch
            self._bufs = [torch.empty(shape, dtype=torch.from_numpy(np.empty(0, dtype)).dtype, pin_memory=True).numpy()
                          for _ in range(self._n)]
        b = self._bufs[self._next % self._n]
        self._next += 1
        return b

    def page(self, path, index) -> np.ndarray:
        p = self._file(path).pages[index]
        return p.read_into(self._staging(p.shape, p.dtype.newbyteorder("=")))

    def dataset(self):
        return {cyc: {ch: {z: (lambda pi=pi: self.page(*pi)) for z, pi in zs.items()} for ch, zs in chs.items()}
                for cyc, chs in self.layout.items()}

    def close(self):
        for f in self._files.values():
            f.close()
        self._files.clear()


class TiffStackSink:
    """sink(cycle, channel, z, image) writing into one contiguous BigTIFF stack (1, C_total, Z, Y, X) --
    the reference's SaveOutputToCycleStack layout (__main__.py:372-398)."""

    def __init__(self, path, channel_index: Dict[Tuple[int, str], int], n_z: int, yx: Tuple[int, int], dtype,
                 description: Optional[str] = None):
        self.index = channel_index
        self.mm = memmap(path, (1, len(channel_index), n_z) + tuple(yx), dtype, description)

    def __call__(self, cyc, ch, z, image: np.ndarray):
        self.mm[0, self.index[(cyc, ch)], z] = image

    def close(self):
        self.mm.flush()
        del self.mm
