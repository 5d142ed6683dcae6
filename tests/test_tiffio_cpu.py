"""CPU: the minimal (Big)TIFF reader / writer against Pillow (libtiff) and cv2.imreadmulti."""
import struct

import numpy as np
import pytest

from microaligner_b200 import tiffio

PIL = pytest.importorskip("PIL.Image")
cv2 = pytest.importorskip("cv2")

OME = ('<?xml version="1.0"?><OME xmlns="http://www.openmicroscopy.org/Schemas/OME/2016-06"><Image ID="Image:0">'
       '<Pixels ID="Pixels:0" DimensionOrder="XYZCT" Type="uint16" SizeX="70" SizeY="50" SizeZ="3" SizeC="2" SizeT="1"/>'
       '</Image></OME>')


def pages(n, shape=(50, 70), dtype=np.uint16, seed=0):
    rng = np.random.default_rng(seed)
    return [rng.integers(0, np.iinfo(dtype).max + 1, shape).astype(dtype) for _ in range(n)]


@pytest.mark.parametrize("dtype", [np.uint8, np.uint16])
@pytest.mark.parametrize("big", [False, True])
def test_read_files_written_by_pillow(tmp_path, dtype, big):
    a = pages(4, dtype=dtype)
    p = tmp_path / "pil.tif"
    ims = [PIL.fromarray(x) for x in a]
    ims[0].save(p, save_all=True, append_images=ims[1:], compression=None, big_tiff=big, description="hello")
    with tiffio.TiffFile(p) as tif:
        assert tif.bigtiff == big and len(tif.pages) == 4
        assert tif.series[0].shape == (4, 50, 70) and tif.series[0].dtype == np.dtype(dtype)
        assert tif.pages[0].description == "hello"
        for i, x in enumerate(a):
            assert np.array_equal(tif.series[0].pages[i].asarray(), x)
            out = np.empty((50, 70), dtype)
            assert np.array_equal(tif.pages[i].read_into(out), x)


def test_memmap_roundtrip_and_foreign_readers(tmp_path):
    a = np.stack(pages(6)).reshape(1, 2, 3, 50, 70)
    p = tmp_path / "stack.tif"
    mm = tiffio.memmap(p, a.shape, a.dtype, description=OME)
    mm[...] = a
    mm.flush()
    del mm
    with tiffio.TiffFile(p) as tif:
        assert tif.bigtiff and tif.ome_metadata == OME
        assert tif.series[0].axes == "CZYX" and tif.series[0].shape == (2, 3, 50, 70)
        assert all(pg.is_contiguous for pg in tif.pages)
        assert np.array_equal(tif.asarray(), a[0])
    ok, got = cv2.imreadmulti(str(p), flags=cv2.IMREAD_UNCHANGED)      # libtiff through OpenCV
    assert ok and len(got) == 6 and all(np.array_equal(g, x) for g, x in zip(got, a.reshape(6, 50, 70)))
    im = PIL.open(p)                                                   # libtiff through Pillow
    for i in range(6):
        im.seek(i)
        assert np.array_equal(np.array(im), a.reshape(6, 50, 70)[i])
    assert im.tag_v2[270] == OME


def test_big_endian_and_multi_strip(tmp_path):
    """A hand-built big-endian classic TIFF with two strips that are not adjacent in the file."""
    img = (np.arange(6 * 5) * 1000).astype(">u2").reshape(6, 5)
    s0, s1 = img[:4].tobytes(), img[4:].tobytes()
    data0_off, data1_off = 8, 8 + len(s0) + 6                # 6 bytes of padding between the strips
    ifd_off = data1_off + len(s1)
    ifd_off += ifd_off % 2
    entries = [(256, 3, 1, 5), (257, 3, 1, 6), (258, 3, 1, 16), (259, 3, 1, 1), (262, 3, 1, 1), (277, 3, 1, 1),
               (278, 3, 1, 4), (339, 3, 1, 1)]
    extra_off = ifd_off + 2 + 10 * 12 + 4
    blob = bytearray(b"MM" + struct.pack(">HI", 42, ifd_off))
    blob += s0 + b"\x00" * 6 + s1
    blob += b"\x00" * (ifd_off - len(blob))
    blob += struct.pack(">H", 10)
    for tag, typ, cnt, val in entries:
        blob += struct.pack(">HHI", tag, typ, cnt) + struct.pack(">H", val) + b"\x00\x00"
    blob += struct.pack(">HHII", 273, 4, 2, extra_off) + struct.pack(">HHII", 279, 4, 2, extra_off + 8)
    blob += struct.pack(">I", 0)
    blob += struct.pack(">II", data0_off, data1_off) + struct.pack(">II", len(s0), len(s1))
    p = tmp_path / "be.tif"
    p.write_bytes(bytes(blob))
    with tiffio.TiffFile(p) as tif:
        pg = tif.pages[0]
        assert tif.byteorder == ">" and not pg.is_contiguous
        assert np.array_equal(pg.asarray(), img.astype(np.uint16))
        out = np.empty((6, 5), np.uint16)
        assert np.array_equal(pg.read_into(out), img.astype(np.uint16))


@pytest.mark.parametrize("dtype", [np.uint8, np.uint16])
@pytest.mark.parametrize("comp", ["tiff_lzw", "tiff_adobe_deflate"])
@pytest.mark.parametrize("predictor", [False, True])
def test_compressed_strips_written_by_libtiff(tmp_path, dtype, comp, predictor):
    """LZW / deflate strips, with and without the horizontal predictor, decode to what libtiff encoded."""
    rng = np.random.default_rng(5)
    ramp = np.add.outer(np.arange(300), np.arange(410)) * (np.iinfo(dtype).max // 800)
    a = (ramp + rng.integers(0, 4, ramp.shape)).astype(dtype)          # smooth + noise: long and short LZW strings
    a[100:140] = rng.integers(0, np.iinfo(dtype).max, (40, 410))       # incompressible band: table resets
    p = tmp_path / "c.tif"
    PIL.fromarray(a).save(p, compression=comp, **({"tiffinfo": {317: 2}} if predictor else {}))
    with tiffio.TiffFile(p) as tif:
        pg = tif.pages[0]
        assert pg.compression in (5, 8) and pg.predictor == (2 if predictor else 1) and not pg.is_contiguous
        assert np.array_equal(pg.asarray(), a)
        out = np.empty(a.shape, dtype)
        assert np.array_equal(pg.read_into(out), a)


def _tiled_tiff(img, tile, compress):
    """Hand-built little-endian classic TIFF with TileWidth/TileLength and optional deflate."""
    import zlib
    h, w = img.shape
    tl, tw = tile
    blobs = []
    for y in range(0, h, tl):
        for x in range(0, w, tw):
            t = np.zeros((tl, tw), img.dtype)
            part = img[y:y + tl, x:x + tw]
            t[:part.shape[0], :part.shape[1]] = part
            raw = t.tobytes()
            blobs.append(zlib.compress(raw) if compress else raw)
    n = len(blobs)
    offs, pos = [], 8
    for b in blobs:
        offs.append(pos)
        pos += len(b) + (len(b) & 1)
    arr_off = pos
    ifd_off = arr_off + 8 * n
    body = bytearray(b"II" + struct.pack("<HI", 42, ifd_off))
    for b in blobs:
        body += b + b"\x00" * (len(b) & 1)
    body += struct.pack(f"<{n}I", *offs) + struct.pack(f"<{n}I", *[len(b) for b in blobs])
    bits = img.dtype.itemsize * 8
    entries = [(256, 3, 1, w), (257, 3, 1, h), (258, 3, 1, bits), (259, 3, 1, 8 if compress else 1), (262, 3, 1, 1),
               (277, 3, 1, 1), (322, 3, 1, tw), (323, 3, 1, tl), (324, 4, n, arr_off), (325, 4, n, arr_off + 4 * n),
               (339, 3, 1, 1)]
    body += struct.pack("<H", len(entries))
    for tag, typ, cnt, val in entries:
        if typ == 3 and cnt == 1:
            body += struct.pack("<HHIHH", tag, typ, cnt, val, 0)
        else:
            body += struct.pack("<HHII", tag, typ, cnt, val)
    body += struct.pack("<I", 0)
    return bytes(body)


@pytest.mark.parametrize("compress", [False, True])
def test_tiled_pages(tmp_path, compress):
    """Tiled pages (the layout pyramidal OME-TIFF writers use), ragged edge tiles included; cross-checked with libtiff."""
    a = np.random.default_rng(2).integers(0, 65535, (70, 100)).astype(np.uint16)
    p = tmp_path / "tiled.tif"
    p.write_bytes(_tiled_tiff(a, (32, 48), compress))
    assert np.array_equal(np.array(PIL.open(p)), a)                    # the hand-built file is a valid TIFF
    with tiffio.TiffFile(p) as tif:
        pg = tif.pages[0]
        assert pg.tiled and len(pg.segments) == 9 and not pg.is_contiguous
        assert np.array_equal(pg.asarray(), a)


def test_unsupported_files_fail_loudly(tmp_path):
    a = pages(1)[0].astype(np.uint8)
    p = tmp_path / "jpeg.tif"
    PIL.fromarray(a).save(p, compression="jpeg")
    with pytest.raises(tiffio.TiffFormatError, match="compressed"):
        tiffio.TiffFile(p)
    rgb = np.zeros((8, 8, 3), np.uint8)
    p2 = tmp_path / "rgb.tif"
    PIL.fromarray(rgb).save(p2, compression=None)
    with pytest.raises(tiffio.TiffFormatError, match="SamplesPerPixel"):
        tiffio.TiffFile(p2)
    p3 = tmp_path / "junk.tif"
    p3.write_bytes(b"not a tiff at all")
    with pytest.raises(tiffio.TiffFormatError):
        tiffio.TiffFile(p3)


def test_page_provider_and_sink(tmp_path):
    a = pages(4, seed=3)
    p = tmp_path / "cyc1.tif"
    tiffio.imwrite(p, np.stack(a))
    prov = tiffio.TiffPageProvider({1: {"DAPI": {0: (str(p), 0), 1: (str(p), 1)}, "CD3": {0: (str(p), 2), 1: (str(p), 3)}}},
                                   pinned=False)
    ds = prov.dataset()
    assert np.array_equal(ds[1]["CD3"][1](), a[3]) and np.array_equal(ds[1]["DAPI"][0](), a[0])
    sink = tiffio.TiffStackSink(tmp_path / "out.tif", {(1, "DAPI"): 0, (1, "CD3"): 1}, 2, (50, 70), np.uint16, OME)
    for ch in ("DAPI", "CD3"):
        for z in (0, 1):
            sink(1, ch, z, ds[1][ch][z]())
    sink.close()
    prov.close()
    with tiffio.TiffFile(tmp_path / "out.tif") as tif:
        assert np.array_equal(tif.asarray().reshape(4, 50, 70), np.stack(a))


def test_native_lzw_decoder_matches_python_and_rejects_garbage(tmp_path):
    """ma_tiff_lzw_decode (host code in the shared library) against the pure-Python decoder on a libtiff-written strip."""
    rng = np.random.default_rng(9)
    a = (np.add.outer(np.arange(64), np.arange(300)) + rng.integers(0, 3, (64, 300))).astype(np.uint16)
    p = tmp_path / "l.tif"
    PIL.fromarray(a).save(p, compression="tiff_lzw")
    with tiffio.TiffFile(p) as tif:
        pg = tif.pages[0]
        off, n, y0, x0, rows, cols = pg.segments[0]
        raw, want = bytes(tif._map[off:off + n]), rows * cols * 2
        assert tiffio._lzw_decode(raw, want) == tiffio._lzw_decode_py(raw, want)
        assert len(tiffio._lzw_decode(raw, want)) == want
        assert np.array_equal(pg.asarray(), a)
    # a stream that starts with a code beyond the table is rejected; a truncated one yields fewer bytes than asked
    with pytest.raises(tiffio.TiffFormatError):
        tiffio._lzw_decode(b"\xff\xff\xff\xff", 16)
    assert len(tiffio._lzw_decode(raw[:len(raw) // 2], want)) < want


def test_page_array_outlives_the_file(tmp_path):
    """The reference returns the page after its with-block (shared_modules/utils.py:69-72): the array must stay valid
    once the TiffFile is closed (a view over the file's own mmap dangled)."""
    a = np.stack(pages(3)).reshape(1, 1, 3, 50, 70)
    p = tmp_path / "stack.tif"
    tiffio.imwrite(p, a)
    with tiffio.TiffFile(p) as tif:
        first = tif.series[0].pages[0].asarray()
        whole = tif.asarray()
    assert np.array_equal(first, a[0, 0, 0]) and np.array_equal(whole.reshape(3, 50, 70), a[0, 0])
    single = tmp_path / "single.tif"
    tiffio.imwrite(single, a[0, 0, 1])
    with tiffio.TiffFile(single) as tif:
        one = tif.asarray()
    assert np.array_equal(one, a[0, 0, 1]) and int(one.sum()) == int(a[0, 0, 1].astype(np.int64).sum())


def test_threaded_copy_and_page_writer(tmp_path):
    """copy_rows (threaded row copy for large pages) and write_page (pwrite behind a memory-mapped stack) are plain copies."""
    rng = np.random.default_rng(3)
    big = rng.integers(0, 65535, (6000, 6000)).astype(np.uint16)          # 72 MB: takes the threaded path
    dst = np.zeros_like(big)
    tiffio.copy_rows(dst, big)
    assert np.array_equal(dst, big)
    p = tmp_path / "stack.tif"
    a = np.stack(pages(6)).reshape(1, 2, 3, 50, 70)
    mm = tiffio.memmap(p, a.shape, a.dtype)
    for c in range(2):
        for z in range(3):
            tiffio.write_page(mm, (0, c, z), a[0, c, z])
    assert np.array_equal(np.asarray(mm), a)          # the mapping sees what pwrite wrote
    mm.flush()
    del mm
    tiffio.close_writers()
    with tiffio.TiffFile(p) as tif:
        assert np.array_equal(tif.asarray().reshape(a.shape), a)
    plain = np.zeros((2, 50, 70), np.uint16)          # not a memmap: falls back to an assignment
    tiffio.write_page(plain, (1,), a[0, 0, 0])
    assert np.array_equal(plain[1], a[0, 0, 0])
