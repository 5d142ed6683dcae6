"""CPU: row-band / tile-range layout of the sharded engine (microaligner_b200.engine.LevelLayout)."""
import pytest

from microaligner_b200 import parallel
from microaligner_b200.engine import LevelLayout


class FakeComm:
    def __init__(self, rank, world):
        self.rank, self.world = rank, world

    def tile_row_bands(self, ny):
        return parallel.split_even(ny, self.world)


@pytest.mark.parametrize("h,w,T,ov,world", [(20000, 20000, 1000, 100, 8), (2500, 2500, 1000, 100, 8), (5000, 5300, 1000, 100, 4),
                                            (700, 820, 150, 20, 5), (1250, 1250, 1000, 100, 8)])
def test_level_layout(h, w, T, ov, world):
    layouts = [LevelLayout(h, w, T, ov, FakeComm(r, world)) for r in range(world)]
    L = layouts[0]
    assert L.tiled == (not max(h, w) / T < 2)
    if not L.tiled:
        assert not L.sharded and L.bands == [(0, h)] * world and L.fb_tiles == [(0, L.ny * L.nx)] * world
        return
    # bands partition the rows on tile-row boundaries; Farneback tiles are balanced to within one tile
    assert L.bands[0][0] == 0 and L.bands[-1][1] <= h and max(b for _, b in L.bands) == h
    for (a0, a1), (b0, b1) in zip(L.bands, L.bands[1:]):
        assert a1 == b0 or b1 == b0            # contiguous (trailing bands may be empty)
    assert all(a % T == 0 for a, b in L.bands if b > a)
    sizes = [b - a for a, b in L.fb_tiles]
    assert sum(sizes) == L.ny * L.nx and max(sizes) - min(sizes) <= 1
    # the rows a rank's Farneback tiles read cover the windows of all its tiles
    for r, (t0, t1) in enumerate(L.fb_tiles):
        rows = L.fb_window_rows()[r]
        for t in range(t0, t1):
            i = t // L.nx
            assert rows[0] <= max(i * T - ov, 0) and rows[1] >= min((i + 1) * T + ov, h)
    # grow() clips to the image and leaves empty bands empty
    for (a, b), (ga, gb) in zip(L.bands, L.grow(ov, ov + 7)):
        if b > a:
            assert ga == max(a - ov, 0) and gb == min(b + ov + 7, h)
        else:
            assert gb == ga
    assert all(l.band == L.bands[r] for r, l in enumerate(layouts))


@pytest.mark.parametrize("shape,T,ov,world,use_dog", [((50000, 50000), 1000, 100, 8, False), ((20000, 20000), 1000, 100, 8, True),
                                                      ((5000, 5300), 1000, 100, 4, False), ((2300, 2100), 300, 40, 3, True),
                                                      ((1500, 1700), 150, 20, 5, False)])
def test_local_pyramid_requirements(shape, T, ov, world, use_dog):
    """Band-local pyramid (Engine.local_pyramid): what a rank computes of every level covers what it reads of that
    level and what the 5-tap pyrDown of the next coarser level reaches; unsharded levels are complete."""
    from microaligner_b200.engine import Engine
    for rank in range(world):
        comm = FakeComm(rank, world)
        eng = Engine.__new__(Engine)
        eng.num_pyr_lvl = 4
        shapes = eng.level_shapes(shape)
        assert shapes[0] == ((shape[0] + 1) // 2, (shape[1] + 1) // 2) and len(shapes) <= 4
        gen = [LevelLayout(h, w, T, ov, comm) for h, w in shapes]
        need = [L.input_rows(use_dog) for L in gen]
        req = Engine.pyramid_requirements(need, [h for h, _ in shapes])
        for k, L in enumerate(gen):
            a, b = req[k]
            na, nb = need[k]
            assert 0 <= a <= b <= L.h
            if nb > na:
                assert a <= na and b >= nb
            if not L.sharded:
                assert (a, b) == (0, L.h)
            else:   # the rows every stage of register() reads on a sharded level (engine.py)
                B = L.band
                over = -(-T * T // L.w) + 1
                reads = []
                if B[1] > B[0]:
                    reads += [(B[0] - ov, B[1] + ov), (B[0] - 20, B[1] + over + 20)]
                f = L.fb_window_rows(20 if use_dog else 0)[rank]
                if f[1] > f[0]:
                    reads.append(f)
                for ra, rb in reads:
                    assert a <= max(ra, 0) and b >= min(rb, L.h)
            if k + 1 < len(gen) and req[k + 1][1] > req[k + 1][0]:
                ca, cb = req[k + 1]
                assert a <= max(2 * ca - 2, 0) and b >= min(2 * cb + 2, L.h)


@pytest.mark.parametrize("shape,world,frac", [((50000, 50000), 8, 0.18), ((50000, 50000), 2, 0.55), ((20000, 20000), 4, 0.31)])
def test_pyramid_plan_keeps_the_upload_share_small(shape, world, frac):
    """Hybrid pyramid (Engine.pyramid_plan): large levels band-local, the first level lower than GATHER_BELOW rows
    gathered, coarser ones replicated -- so a rank reads about 1/world of the full-resolution images, not the third
    that the coarse levels' tile rows would map to."""
    from microaligner_b200.engine import Engine
    for rank in range(world):
        eng = Engine(1000, 100, 4, 3, True, False, comm=FakeComm(rank, world))
        eng.local_pyramid = True
        shapes, g, req, slices = eng.pyramid_plan(shape)
        assert g is not None and shapes[g][0] < Engine.GATHER_BELOW and (g == 0 or shapes[g - 1][0] >= Engine.GATHER_BELOW)
        assert len(req) == g + 1 and req[g] == slices[rank]
        for k in range(g):       # band-local levels: own reads + support of the next level's rows
            L = LevelLayout(shapes[k][0], shapes[k][1], 1000, 100, FakeComm(rank, world))
            na, nb = L.input_rows(False)
            assert req[k][0] <= na and req[k][1] >= nb
            assert req[k][0] <= max(2 * req[k + 1][0] - 2, 0) and req[k][1] >= min(2 * req[k + 1][1] + 2, L.h)
        r0, r1 = eng.full_input_rows(shape)
        assert (r1 - r0) <= frac * shape[0], (rank, r0, r1)
