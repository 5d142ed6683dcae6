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


def test_unsupported_files_fail_loudly(tmp_path):
    a = pages(1)[0]
    p = tmp_path / "lzw.tif"
    PIL.fromarray(a).save(p, compression="tiff_lzw")
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
