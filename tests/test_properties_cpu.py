"""CPU property tests (hypothesis): tile geometry, partition and transfer-plan invariants over random shapes
within the YAML ranges of the reference (config_reader.py:93-97: TileSize >= 20, Overlap in [10, TileSize])."""
import numpy as np
from hypothesis import given, settings, strategies as st

from microaligner_b200 import parallel
from oracle import reference_flow as rf


@settings(max_examples=40, deadline=None, derandomize=True)
@given(h=st.integers(100, 900), w=st.integers(100, 900), T=st.integers(20, 400), data=st.data())
def test_split_stitch_roundtrip(h, w, T, data):
    ov = data.draw(st.integers(10, max(10, min(T, 60))))
    rng = np.random.default_rng(h * 1000 + w)
    a = rng.integers(0, 65535, (h, w)).astype(np.uint16)
    tiles = rf.split(a, T, ov)
    ny, nx = rf.tile_grid(h, w, T)
    assert len(tiles) == ny * nx and all(t.shape == (T + 2 * ov, T + 2 * ov) for t in tiles)
    assert np.array_equal(rf.stitch(tiles, h, w, T, ov), a)
    # the zero padding really is zero and the centre really is the image
    t0 = tiles[0]
    assert not t0[:ov].any() and not t0[:, :ov].any()
    assert np.array_equal(t0[ov:ov + min(T, h), ov:ov + min(T, w)], a[:min(T, h), :min(T, w)])


@settings(max_examples=60, deadline=None, derandomize=True)
@given(ny=st.integers(1, 40), nx=st.integers(1, 40), world=st.integers(1, 9), T=st.integers(20, 300), data=st.data())
def test_tile_and_band_partitions_cover_exactly_once(ny, nx, world, T, data):
    h = (ny - 1) * T + data.draw(st.integers(1, T))
    w = (nx - 1) * T + data.draw(st.integers(1, T))
    tiles = parallel.split_even(ny * nx, world)
    assert tiles[0][0] == 0 and tiles[-1][1] == ny * nx and all(a[1] == b[0] for a, b in zip(tiles, tiles[1:]))
    assert max(b - a for a, b in tiles) - min(b - a for a, b in tiles) <= 1          # balanced
    cover = np.zeros((h, w), np.int32)
    for t in tiles:
        rects = parallel.tile_range_rects(t, nx, T, h, w)
        assert len(rects) <= 3
        for (y0, y1, x0, x1) in rects:
            cover[y0:y1, x0:x1] += 1
    assert (cover == 1).all()


@settings(max_examples=40, deadline=None, derandomize=True)
@given(ny=st.integers(1, 12), nx=st.integers(1, 8), world=st.integers(2, 8), ov=st.integers(0, 40))
def test_rect_plan_delivers_every_needed_row_once(ny, nx, world, ov):
    T = 50
    h, w = ny * T - 7, nx * T - 3
    have = [parallel.tile_range_rects(t, nx, T, h, w) for t in parallel.split_even(ny * nx, world)]
    bands = [(a * T, min(b * T, h)) for a, b in parallel.split_even(ny, world)]
    need = [(max(a - ov, 0), min(b + ov, h)) if b > a else (a, a) for a, b in bands]
    plan = parallel.rect_plan(have, need)
    for dst, nd in enumerate(need):
        got = np.zeros((h, w), np.int32)
        for (y0, y1, x0, x1) in have[dst]:
            got[y0:y1, x0:x1] += 1
        for (s, d, (y0, y1, x0, x1)) in plan:
            assert s != d
            if d == dst:
                got[y0:y1, x0:x1] += 1
        assert (got[nd[0]:nd[1]] == 1).all()


@settings(max_examples=60, deadline=None, derandomize=True)
@given(h=st.integers(50, 5000), w=st.integers(50, 5000), T=st.integers(20, 1000), world=st.integers(1, 8))
def test_nmi_chunks_have_one_owner(h, w, T, world):
    ny = -(-h // T)
    bands = [(a * T, min(b * T, h)) for a, b in parallel.split_even(ny, world)]
    n, chunk = h * w, T * T
    nchunks = -(-n // chunk)
    seen = np.zeros(nchunks, np.int32)
    for b in bands:
        c0, c1 = parallel.chunk_range_of_band(b, w, chunk, n)
        seen[c0:c1] += 1
    assert (seen == 1).all()


@settings(max_examples=30, deadline=None, derandomize=True)
@given(n=st.integers(1, 60), world=st.integers(1, 6), data=st.data())
def test_row_transfer_plan(n, world, data):
    owned = parallel.split_even(n, world)
    need = []
    for a, b in owned:
        lo, hi = data.draw(st.integers(0, 5)), data.draw(st.integers(0, 5))
        need.append((max(a - lo, 0), min(b + hi, n)) if b > a else (a, a))
    plan = parallel.transfer_plan(owned, need)
    for dst, nd in enumerate(need):
        rows = np.zeros(n, np.int32)
        rows[owned[dst][0]:owned[dst][1]] += 1
        for s, d, (a, b) in plan:
            if d == dst:
                assert owned[s][0] <= a and b <= owned[s][1]
                rows[a:b] += 1
        assert (rows[nd[0]:nd[1]] == 1).all()
