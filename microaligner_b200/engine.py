"""Device engine of OptFlowRegistrator.register() / Warper.warp(): the reference's coarse-to-fine loop
(optflow_reg/optflow_registrator.py:93-173) on device tensors, written once for 1..N GPUs.

Sharding (parallel.Comm): at every tiled pyramid level rank r owns a contiguous band of
tile rows, and -- independently, for balance -- an equal share of the level's TILES for the Farneback
flow, which is 4/5 of the work.  Each stage computes only what it owns; what a later stage needs beyond
that (tile-window overlap, the 20-row DoG halo, the rows an NMI chunk runs past the band, the rows pyrUp
reads, the centres of off-band Farneback tiles) is fetched from the owner with point-to-point row /
rectangle exchanges.  Global scalars -- DoG min/max and the per-chunk NMI scores -- are all-reduced
(one collective per batch), so every rank takes the same Better/Worse decision.
Levels the reference computes untiled (max(shape)/tile_size < 2) are tiny and run replicated.
With one rank every exchange is a no-op and the code below is simply the single-GPU path."""
import contextlib
import os
import time
from collections import defaultdict
from math import log2
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import ops, parallel

Range = Tuple[int, int]


def _nvtx():
    """torch's NVTX binding (None when this build of torch has none)."""
    try:
        return torch.cuda.nvtx if torch.cuda.is_available() else None
    except Exception:  # noqa: BLE001
        return None


def _clip(a: int, b: int, h: int) -> Range:
    a = min(max(a, 0), h)
    return (a, min(max(b, a), h))


def _union(a: Range, b: Range) -> Range:
    return (min(a[0], b[0]), max(a[1], b[1]))


class LevelLayout:
    """Who owns which rows of one pyramid level."""

    def __init__(self, h: int, w: int, T: int, ov: int, comm: parallel.Comm):
        self.h, self.w, self.T, self.ov = h, w, T, ov
        self.tiled = not (max(h, w) / T < 2)          # flow_calc.py:60-64 and similarity_scoring.py:35
        self.ny, self.nx = -(-h // T), -(-w // T)
        self.sharded = self.tiled and comm.world > 1
        if self.sharded:
            self.tile_rows = comm.tile_row_bands(self.ny)
            self.bands = [_clip(a * T, b * T, h) for a, b in self.tile_rows]
            # Farneback (4/5 of the work) is balanced at TILE granularity, independently of the row bands
            self.fb_tiles = parallel.split_even(self.ny * self.nx, comm.world)
        else:
            self.tile_rows = [(0, self.ny)] * comm.world
            self.bands = [(0, h)] * comm.world
            self.fb_tiles = [(0, self.ny * self.nx)] * comm.world
        self.rank = comm.rank
        self.world = comm.world

    def grow(self, lo: int, hi: int) -> List[Range]:
        """Every rank's band extended by lo rows upwards and hi rows downwards (clipped)."""
        return [_clip(a - lo, b + hi, self.h) if b > a else (a, a) for a, b in self.bands]

    @property
    def band(self) -> Range:
        return self.bands[self.rank]

    def fb_window_rows(self, halo: int = 0) -> List[Range]:
        """Image rows covered by the tile windows of every rank's Farneback tiles (+ halo)."""
        out = []
        for t0, t1 in self.fb_tiles:
            if t1 <= t0:
                out.append((0, 0))
            else:
                i0, i1 = t0 // self.nx, (t1 - 1) // self.nx
                out.append(_clip(i0 * self.T - self.ov - halo, (i1 + 1) * self.T + self.ov + halo, self.h))
        return out

    def fb_rects(self):
        return [parallel.tile_range_rects(t, self.nx, self.T, self.h, self.w) for t in self.fb_tiles]

    def input_rows(self, use_dog: bool) -> Range:
        """Rows of this level's ref / mov pyramid images that THIS rank ever reads in Engine.register(): its band with
        the tile-window overlap (pre-/post-warp), the NMI chunk overrun and the 20-row DoG halo of the gate, and the
        tile windows of its Farneback tiles (+ DoG halo when the flow runs on DoG images); a few spare rows on top.
        An unsharded level is read completely."""
        if not self.sharded:
            return (0, self.h)
        over = -(-self.T * self.T // self.w) + 1
        a, b = self.band
        need = None
        if b > a:
            need = _clip(a - self.ov - 24, b + max(self.ov, over + 20) + 24, self.h)
        fa, fb = self.fb_window_rows(24 if use_dog else 4)[self.rank]
        if fb > fa:
            need = (fa, fb) if need is None else _union(need, (fa, fb))
        return need if need is not None else (0, 0)


class Engine:
    def __init__(self, tile_size=1000, overlap=100, num_pyr_lvl=4, num_iterations=3, use_full_res_img=False,
                 use_dog=False, comm: Optional[parallel.Comm] = None, log: Callable[[str], None] = None,
                 contract_fma: bool = False, corrected: bool = False):
        self.T, self.ov = int(tile_size), int(overlap)
        self.num_pyr_lvl, self.iters = int(num_pyr_lvl), int(num_iterations)
        self.full_res, self.use_dog = bool(use_full_res_img), bool(use_dog)
        self.comm = comm or parallel.get()
        self.win = self.ov - (1 - self.ov % 2)         # optflow_registrator.py:91
        self._log = log
        self.contract_fma = bool(contract_fma)   # opt-in fast window blur (not bit-identical to OpenCV)
        # opt-in, NOT the reference's behaviour: compose flows properly (m = f2 + f1(p - f2)) and scale the final
        # up-sampling by 2 -- removes quirks Q1 / Q2 that keep the reference's multi-level flow far from ground truth
        self.corrected = bool(corrected)
        self.decisions: List[dict] = []
        self.force_decisions = None  # test hook (list of bools per level), mirrors oracle.reference_flow.register
        self.gather_flow = True      # False: with several ranks the returned flow is valid on this rank's band only
        # several ranks only: every rank reduces just the rows of each pyramid level it will read
        # (LevelLayout.input_rows) from its copy of the input -- of which only Engine.full_input_rows() need to be valid
        # -- instead of computing a slice of every level and gathering the levels over NVLink.  Rows outside that range
        # stay uninitialised.  MA_LOCAL_PYRAMID=0 selects the gathering variant (A/B measurements).
        self.local_pyramid = os.environ.get("MA_LOCAL_PYRAMID", "1") not in ("", "0")
        self.GATHER_BELOW = int(os.environ.get("MA_PYRAMID_GATHER_BELOW", Engine.GATHER_BELOW))   # see pyramid_plan()
        # rows of the final flow are handed to a host sink as soon as they are final; on one GPU the last level runs its
        # Farneback tile row by tile row and merges / downloads speculatively behind it (see register())
        self.stream_groups = True
        # A level's accept / reject decision is read back from the device only after the NEXT level's warp and DoG images
        # have been enqueued on the assumption "accepted", so the device queue does not run dry at every level
        # (MA_DEFER_GATE=0: read it back at once).  Results do not depend on it.
        self.defer_gate = os.environ.get("MA_DEFER_GATE", "1") not in ("", "0")
        self.gather_pieces = 4       # pieces in which a sharded warp() sends its band to the other ranks
        self.group_tiles = 40        # tiles per Farneback launch of that streamed last level (enough CTAs to fill 148 SMs)
        self.flow_layout = None

    def log(self, *a):
        if self.comm.rank == 0:
            (self._log or print)(*a)

    # Every phase of register() / warp() is an NVTX range (visible in Nsight Systems / ncu --nvtx), nested in one range
    # per pyramid level.  Optional phase timing (Engine.trace = True): device-synchronised wall time per phase, summed
    # in Engine.times -- bench.py prints it as `phases_ms`.
    trace = False
    times = defaultdict(float)
    # Engine.timeline = [] records, without synchronising, when the host entered / left every phase and a device event at
    # both points (scripts/timeline.py lines them up: where the device waits for the host or for a peer).
    timeline = None

    @contextlib.contextmanager
    def phase(self, name):
        nvtx = _nvtx()
        if nvtx is not None:
            nvtx.range_push("ma:" + name)
        try:
            if Engine.timeline is not None:     # no synchronisation: CPU enqueue times and device events per phase
                rec = [name, time.perf_counter(), torch.cuda.Event(enable_timing=True), None, torch.cuda.Event(enable_timing=True)]
                rec[2].record()
                Engine.timeline.append(rec)
                try:
                    yield
                finally:
                    rec[3] = time.perf_counter()
                    rec[4].record()
                return
            if not Engine.trace:
                yield
                return
            torch.cuda.synchronize()
            t = time.perf_counter()
            yield
            torch.cuda.synchronize()
            Engine.times[name] += time.perf_counter() - t
        finally:
            if nvtx is not None:
                nvtx.range_pop()

    # ------------------------------------------------------------------ building blocks
    def pyramid(self, arr: torch.Tensor):
        if self.num_pyr_lvl < 0:
            raise ValueError("Number of pyramid levels cannot be less than 0")
        if self.num_pyr_lvl == 0 and not self.full_res:
            raise ValueError("Number of pyramid levels is 0 and use_full_res_img is False. "
                             "Please change one of the parameters")
        pyr, factors, cur = [], [], arr
        for lvl in range(self.num_pyr_lvl):
            factor = 2 ** (lvl + 1)
            if arr.shape[0] / factor < 100 or arr.shape[1] / factor < 100:
                break
            if self.comm.world > 1 and cur.shape[0] >= 16384:   # below that the gather latency outweighs the saving
                # every rank reduces its slice of rows, then the (4x smaller) level is gathered over NVLink
                oh, ow = (cur.shape[0] + 1) // 2, (cur.shape[1] + 1) // 2
                bands = parallel.split_even(oh, self.comm.world)
                nxt = torch.empty((oh, ow), dtype=cur.dtype, device=cur.device)
                ops.pyr_down_rows(cur, bands[self.comm.rank], nxt)
                cur = self.comm.gather_rows(nxt, bands)
            else:
                cur = ops.pyr_down(cur)
            pyr.append(cur)
            factors.append(factor)
        pyr.reverse()
        factors.reverse()
        if self.full_res:
            pyr.append(arr)
            factors.append(1)
        return pyr, factors

    def level_shapes(self, shape) -> List[Tuple[int, int]]:
        """Shapes of the pyrDown chain in generation order (fine -> coarse), with pyramid()'s stopping rule."""
        out, (h, w) = [], shape
        for lvl in range(self.num_pyr_lvl):
            factor = 2 ** (lvl + 1)
            if shape[0] / factor < 100 or shape[1] / factor < 100:
                break
            h, w = (h + 1) // 2, (w + 1) // 2
            out.append((h, w))
        return out

    @staticmethod
    def pyramid_requirements(need: Sequence[Range], heights: Sequence[int]) -> List[Range]:
        """need[k] = rows of level k (generation order, fine -> coarse) a rank reads itself; result[k] = rows it has to
        COMPUTE: its own plus the rows the 5-tap pyrDown of the next coarser level's requirement reaches (2a-2 .. 2b+1)."""
        req = list(need)
        for k in range(len(need) - 2, -1, -1):
            a, b = req[k + 1]
            if b > a:
                sup = _clip(2 * a - 2, 2 * b + 2, heights[k])
                req[k] = sup if req[k][1] <= req[k][0] else _union(req[k], sup)
        return req

    GATHER_BELOW = 16384      # pyramid levels lower than this are completed on every rank (slices gathered over NVLink)

    def pyramid_plan(self, shape):
        """How one rank of several builds its pyramids.  Tiles have the same size at every level, so a band of tile rows
        of a COARSE level maps to a large part of the full-resolution image (a third of it at the coarsest level of a
        50 000^2 slide on 8 ranks): building every level band-locally would make each rank read and reduce that much.
        Instead the first level lower than GATHER_BELOW rows is computed in equal slices and gathered (a few hundred MB
        over NVLink), the coarser ones are reduced from it, replicated, and only the large levels above it are band-local.
        Returns (shapes fine -> coarse, index of the gathered level or None, rows to compute per level, its slices)."""
        shapes = self.level_shapes(tuple(shape))
        gen = [LevelLayout(h, w, self.T, self.ov, self.comm) for h, w in shapes]
        g = next((k for k, (h, _) in enumerate(shapes) if h < self.GATHER_BELOW), None)
        top = len(shapes) if g is None else g + 1                      # levels [0, top) are computed by rows
        need = [L.input_rows(self.use_dog) for L in gen[:top]]
        slices = None
        if g is not None:
            slices = parallel.split_even(shapes[g][0], self.comm.world)
            need[g] = slices[self.comm.rank]                           # the gather completes the level anyway
        req = self.pyramid_requirements(need, [h for h, _ in shapes[:top]]) if need else []
        return shapes, g, req, slices

    def pyramid_local(self, arr: torch.Tensor, plan):
        """pyramid() for one rank of several (see pyramid_plan): band-local levels hold valid data on their rows only."""
        shapes, g, req, slices = plan
        pyr, cur = [], arr
        for k, (h, w) in enumerate(shapes):
            if g is not None and k > g:
                nxt = ops.pyr_down(cur)                                # small level, replicated
            else:
                nxt = torch.empty((h, w), dtype=arr.dtype, device=arr.device)
                if req[k][1] > req[k][0]:
                    ops.pyr_down_rows(cur, req[k], nxt)
                if k == g:
                    self.comm.gather_rows(nxt, slices)
            pyr.append(nxt)
            cur = nxt
        pyr.reverse()
        if self.full_res:
            pyr.append(arr)
        return pyr

    def dog_batch(self, items, L: Optional[LevelLayout] = None) -> List[torch.Tensor]:
        """uint8 DoG of several images: items = [(img, rows)] of level L or [(img, rows, level)], result i valid on
        rows_i.  Sharded levels need the GLOBAL min/max of every source and of every difference image: every rank
        scans its own band, and all pairs of the batch travel in ONE all-reduce per phase."""
        items = [(it[0], it[1], it[2] if len(it) > 2 else L) for it in items]
        k = len(items)
        dev = items[0][0].device
        sharded = any(Li.sharded for _, _, Li in items)
        mm = torch.empty((k, 2), dtype=torch.float32, device=dev)
        dmm = torch.empty((k, 2), dtype=torch.float32, device=dev)
        for i, (img, rows, Li) in enumerate(items):
            ops.minmax_rows(img, Li.band if Li.sharded else (0, Li.h), out=mm[i])
        if sharded:
            self.comm.allreduce_minmax(mm)
        diffs = [ops.dog_diff_rows(img, mm[i], rows, dmm=dmm[i])[0] for i, (img, rows, Li) in enumerate(items)]
        if sharded:
            self.comm.allreduce_minmax(dmm)
        outs = []
        for i, (img, rows, Li) in enumerate(items):
            out = torch.empty((Li.h, Li.w), dtype=torch.uint8, device=dev)
            outs.append(ops.dog_quantize_rows(diffs[i], Li.h, Li.w, dmm[i], rows, out))
            diffs[i] = None
        return outs

    def level_rows(self, L: LevelLayout):
        """Row ranges of a level on this rank: (gate_rows, fb_rows, ref_dog_rows, over) -- the rows the similarity gate
        reads (the band plus what its last NMI chunk runs past it), the rows this rank's Farneback tiles read, the rows
        of the reference DoG image that serve both, and the gate's overrun."""
        B = L.band
        over = -(-self.T * self.T // L.w) + 1 if L.tiled else 0
        gate_rows = _clip(B[0], B[1] + over, L.h) if L.sharded else (0, L.h)
        if B[1] <= B[0]:                                       # more ranks than tile rows: nothing to do here
            gate_rows = (B[0], B[0])
        fb_rows = L.fb_window_rows()[L.rank] if L.sharded else (0, L.h)
        ref_dog_rows = gate_rows
        if self.use_dog and fb_rows[1] > fb_rows[0]:
            ref_dog_rows = _union(gate_rows, fb_rows) if gate_rows[1] > gate_rows[0] else fb_rows
        return gate_rows, fb_rows, ref_dog_rows, over

    def warp_rows(self, img: torch.Tensor, flow: torch.Tensor, L: LevelLayout, rows: Range) -> torch.Tensor:
        out = torch.empty_like(img)
        return ops.warp_tiles_rows(img, flow, self.T, self.ov, rows, out)

    def mi_scores_async(self, a: torch.Tensor, bs: Sequence[torch.Tensor], L: LevelLayout):
        """mi_tiled(a, b) for every b in bs (similarity_scoring.py:27-50): per-chunk NMI on the device, one all-reduce
        and one read-back for the whole batch, np.mean on the host (rounds like the reference).  Everything is enqueued
        here; the returned callable waits for the read-back and returns the scores."""
        n = a.numel()
        if not L.tiled:
            scores = torch.stack([ops.nmi_chunks(a, b, n)[:1] for b in bs])
        else:
            chunk = self.T * self.T
            nchunks = -(-n // chunk)
            scores = torch.zeros((len(bs), nchunks), dtype=torch.float64, device=a.device)
            cr = parallel.chunk_range_of_band(L.band, L.w, chunk, n) if L.sharded else (0, nchunks)
            if len(bs) == 2:
                ops.nmi_chunk_range2(a, bs[0], bs[1], chunk, cr, scores[0], scores[1])
            else:
                for i, b in enumerate(bs):
                    ops.nmi_chunk_range(a, b, chunk, cr, scores[i])
            if L.sharded:
                self.comm.allreduce_sum(scores)
        done = None
        if scores.is_cuda:
            host = torch.empty(scores.shape, dtype=scores.dtype, pin_memory=True)
            host.copy_(scores, non_blocking=True)
            done = torch.cuda.Event()
            done.record()
        else:
            host = scores

        def result() -> List[float]:
            if done is not None:
                done.synchronize()
            arr = host.numpy()
            return [float(np.mean(arr[i])) for i in range(len(bs))]
        return result

    def mi_scores(self, a: torch.Tensor, bs: Sequence[torch.Tensor], L: LevelLayout) -> List[float]:
        return self.mi_scores_async(a, bs, L)()

    def pyr_up(self, flow: torch.Tensor, Ls: LevelLayout, Ld: LevelLayout, scale: float) -> torch.Tensor:
        """cv.pyrUp(flow*scale) from level Ls to level Ld; every rank produces the rows of its Ld band plus the `ov` rows
        on either side that its tile windows cover (what the merge at level Ld reads), from source rows fetched once."""
        out = torch.empty((Ld.h, Ld.w, 2), dtype=torch.float32, device=flow.device)
        rows = Ld.grow(self.ov, self.ov)
        if Ls.sharded:
            need = [_clip(a // 2 - 2, (b + 1) // 2 + 2, Ls.h) if b > a else (0, 0) for a, b in rows]
            self.comm.exchange_rows(flow, Ls.bands, need)
        return ops.pyr_up_flow_rows(flow, (Ld.h, Ld.w), scale, rows[Ld.rank], out)

    # ------------------------------------------------------------------ register()
    def register(self, ref: torch.Tensor, mov: torch.Tensor, sink=None, ready=(None, None)) -> torch.Tensor:
        """The coarse-to-fine loop.  `sink` (ops.HostSink) receives this rank's rows of the final flow while the device
        keeps working; `ready` = events to wait for before ref / mov are read (asynchronous uploads)."""
        comm, T, ov = self.comm, self.T, self.ov
        full = LevelLayout(ref.shape[0], ref.shape[1], T, ov, comm)
        with self.phase("pyramid"):
            if self.local_pyramid and comm.world > 1:
                plan = self.pyramid_plan(tuple(ref.shape))
                shapes = plan[0]
                if self.num_pyr_lvl < 0 or (not shapes and not self.full_res):
                    self.pyramid(ref)                                               # raises the reference's ValueErrors
                ops.wait_upload(ready[0])
                ref_pyr = self.pyramid_local(ref, plan)
                ops.wait_upload(ready[1])
                mov_pyr = self.pyramid_local(mov, plan)
                factors = [2 ** (k + 1) for k in range(len(shapes))][::-1] + ([1] if self.full_res else [])
            else:
                ops.wait_upload(ready[0])
                ref_pyr, factors = self.pyramid(ref)
                ops.wait_upload(ready[1])                # the pyramid of ref runs while mov is still arriving
                mov_pyr, _ = self.pyramid(mov)
        layouts = [LevelLayout(p.shape[0], p.shape[1], T, ov, comm) for p in ref_pyr]
        num_lvl = len(factors)
        self.decisions = []
        m_flow, m_layout = None, None

        # The DoG images of the similarity gate -- reference and unwarped moving image of every level -- depend on no flow:
        # all levels go through one batch here, i.e. two all-reduces for the whole pyramid instead of two per level.
        with self.phase("dog"):
            items = []
            for k, Lk in enumerate(layouts):
                g_rows, _, r_rows, _ = self.level_rows(Lk)
                items += [(ref_pyr[k], r_rows, Lk), (mov_pyr[k], g_rows, Lk)]
            gate_dogs = self.dog_batch(items) if items else []
            del items
        pending = None       # the previous level's gate: enqueued, assumed to accept, not read back yet
        for lvl, factor in enumerate(factors):
            nvtx = _nvtx()
            if nvtx is not None:
                if lvl > 0:
                    nvtx.range_pop()
                nvtx.range_push(f"ma:level factor {factor}")
            L = layouts[lvl]
            B = L.band
            halo = ov + (20 if self.use_dog else 0)
            gate_rows, fb_rows, _, over = self.level_rows(L)
            ref_dog, od = gate_dogs[2 * lvl], gate_dogs[2 * lvl + 1]
            gate_dogs[2 * lvl] = gate_dogs[2 * lvl + 1] = None

            def head(m_flow, lvl=lvl, L=L, B=B, halo=halo, fb_rows=fb_rows, ref_dog=ref_dog):
                """Everything of this level in front of Farneback: pre-warp by the accumulated flow, its DoG image."""
                mov_l = mov_pyr[lvl]
                if lvl > 0:           # m_flow: from pyr_up(), i.e. valid on the band and the overlap of its tile windows
                    with self.phase("warp"):
                        mov_l = self.warp_rows(mov_l, m_flow, L, B)
                    if L.sharded:
                        dh = 20 if self.use_dog else 0
                        need = [b if a[1] <= a[0] else (_union(a, b) if b[1] > b[0] else a)
                                for a, b in zip(L.grow(halo, halo), L.fb_window_rows(dh))]
                        with self.phase("exchange"):
                            comm.exchange_rows(mov_l, L.bands, need)
                if not self.use_dog:
                    return mov_l, ref_pyr[lvl], mov_l
                with self.phase("dog"):
                    return mov_l, ref_dog, self.dog_batch([(mov_l, fb_rows)], L)[0]

            # The head of this level is enqueued behind the previous level's gate BEFORE that gate is read back, on the
            # assumption that it accepted (it nearly always does): the host only blocks once the device has this much
            # work queued, so the device does not run dry at every level while the host catches up.  A rejected level
            # costs the head twice.
            state = head(m_flow)
            if pending is not None:
                gate, p_lvl, p_flow, p_layout = pending
                pending = None
                if not self._judge(gate, factors[p_lvl], p_lvl):
                    del state
                    with self.phase("merge+pyrup"):
                        m_flow, m_layout = self._rejected(p_lvl, p_flow, p_layout, L, ref.device), L
                    state = head(m_flow)
                del gate, p_flow
            self.log("Pyramid factor", factor)
            mov_l, fb_ref, fb_mov = state
            del state
            this_flow = torch.empty((L.h, L.w, 2), dtype=torch.float32, device=ref.device)
            # Speculative tail of the LAST level: if the gate accepts this level -- it nearly always does -- the result is
            # merge(m_flow, this_flow), which is per tile.  So this rank's Farneback tiles run one group of tile rows at a
            # time, every tile row of its band whose window it has completed itself is merged right behind it, and those rows
            # start their way to the host while the next group is computed.  The tile rows next to another rank's tiles wait
            # for the exchange, as before.  Should the gate reject the level, all rows are simply sent again from the flow
            # that is returned instead.
            speculate = (sink is not None and self.stream_groups and L.tiled and lvl == num_lvl - 1 and lvl > 0
                         and self.full_res and not self.corrected)
            merged, m_lo, m_hi = None, 0, 0                        # tile rows [m_lo, m_hi) of `merged` are done and sent
            with self.phase("farneback" if L.tiled else "farneback(untiled level)"):
                if speculate:
                    merged = torch.empty_like(this_flow)
                    t0, t1 = L.fb_tiles[L.rank]
                    ba, bb = L.tile_rows[L.rank]
                    ia, ib = -(-t0 // L.nx), t1 // L.nx            # tile rows [ia, ib) are computed entirely by this rank
                    fb = lambda a, b: b > a and ops.farneback_tiles(fb_mov, fb_ref, T, ov, self.win, self.iters, (a, b),   # noqa: E731
                                                                    out=this_flow, contract_fma=self.contract_fma)
                    if ib <= ia:                                   # less than one whole tile row: nothing to stream
                        fb(t0, t1)
                    else:
                        fb(t0, ia * L.nx)                          # tail of the previous rank's last tile row
                        # a tile row reads this_flow up to `ov` rows into both neighbouring tile rows
                        m_lo = m_hi = max(ba, ia + (1 if ia > 0 else 0))
                        step = max(1, -(-self.group_tiles // L.nx))     # tile rows per group
                        for i0 in range(ia, ib, step):
                            i1 = min(i0 + step, ib)
                            fb(i0 * L.nx, i1 * L.nx)
                            ready_hi = min(bb, i1 if i1 == L.ny else i1 - 1)
                            if ready_hi > m_hi:
                                ops.merge_flows_tile_rows(m_flow, this_flow, T, ov, (m_hi, ready_hi), merged)
                                sink.push(merged, _clip(m_hi * T, ready_hi * T, L.h))
                                m_hi = ready_hi
                        fb(ib * L.nx, t1)                          # head of the next rank's first tile row
                elif L.tiled:
                    ops.farneback_tiles(fb_mov, fb_ref, T, ov, self.win, self.iters, L.fb_tiles[L.rank], out=this_flow,
                                        contract_fma=self.contract_fma)
                else:
                    ops.farneback_tiles(fb_mov, fb_ref, 0, 0, self.win, self.iters, out=this_flow, contract_fma=self.contract_fma)
            del fb_mov, fb_ref
            if L.sharded:   # tile centres -> the row bands (+ overlap) the row-wise stages work on
                with self.phase("exchange"):
                    comm.exchange_rects(this_flow, L.fb_rects(), L.grow(ov, ov))

            with self.phase("warp"):
                warped = self.warp_rows(mov_l, this_flow, L, B)
            del mov_l
            if L.sharded:
                with self.phase("exchange"):
                    comm.exchange_rows(warped, L.bands, L.grow(20, over + 20))
            with self.phase("dog"):
                wd = self.dog_batch([(warped, gate_rows)], L)[0]
            del warped
            with self.phase("nmi gate"):
                gate = self.mi_scores_async(ref_dog, [wd, od], L)
            del ref_dog, wd, od
            Ln = layouts[lvl + 1] if lvl + 1 < num_lvl else None
            if self.defer_gate and Ln is not None:
                better = True                                      # read back behind the next level's head
                pending = (gate, lvl, m_flow, L)
            else:
                better = self._judge(gate, factor, lvl)
            del gate

            tail = self.phase("merge+pyrup")
            tail.__enter__()
            if better:
                if lvl == 0:
                    if num_lvl > 1:
                        m_flow, m_layout = self.pyr_up(this_flow, L, Ln, 2.0), Ln
                    else:
                        m_flow, m_layout = self._upscale_to_full(this_flow, L, full, factor)
                elif lvl == num_lvl - 1:
                    if merged is None:
                        merged = self._merge(m_flow, this_flow, L)
                    else:                                # the tile rows of the band that were not merged speculatively
                        ba, bb = L.tile_rows[L.rank]
                        for a, b in ((ba, min(m_lo, bb)), (max(m_hi, ba), bb)) if m_hi > m_lo else ((ba, bb),):
                            if b > a:
                                ops.merge_flows_tile_rows(m_flow, this_flow, T, ov, (a, b), merged)
                                sink.push(merged, _clip(a * T, b * T, L.h))
                        sink = None                      # every row of the result is on its way to the host
                    m_flow, m_layout = merged, L
                    if not self.full_res:
                        m_flow, m_layout = self._upscale_to_full(merged, L, full, factor)
                else:
                    merged = self._merge(m_flow, this_flow, L)
                    m_flow, m_layout = self.pyr_up(merged, L, Ln, 2.0), Ln
            elif Ln is not None:
                m_flow, m_layout = self._rejected(lvl, m_flow, L, Ln, ref.device), Ln
            elif lvl == 0:
                m_flow, m_layout = torch.zeros(tuple(mov.shape) + (2,), dtype=torch.float32, device=ref.device), full
            elif not self.full_res:
                m_flow, m_layout = self.pyr_up(m_flow, L, full, 2.0), full
            tail.__exit__(None, None, None)
            del this_flow

        if _nvtx() is not None and factors:
            _nvtx().range_pop()
        if m_flow is None:
            # no pyramid level at all: the reference fails on its unbound local (optflow_registrator.py:173)
            raise UnboundLocalError("cannot access local variable 'm_flow' where it is not associated with a value")
        self.flow_layout = m_layout
        if sink is not None:       # not (or wrongly) speculated: this rank's rows of the result go to the host now
            sink.push(m_flow, self.result_rows(m_layout, m_flow.shape[0]))
        if self.gather_flow and m_layout is not None and m_layout.sharded:
            with self.phase("gather flow"):
                comm.gather_rows(m_flow, m_layout.bands)
        return m_flow

    def _judge(self, gate, factor: int, lvl: int) -> bool:
        """Read a level's gate back and record the decision (optflow_registrator.py:150-171)."""
        after, before = gate()
        self.log("    MI score after:", after, "| MI score before:", before)
        better = after > before
        if self.force_decisions is not None:      # test hook: exercise the "Worse alignment" branches
            better = bool(self.force_decisions[lvl])
        self.decisions.append(dict(factor=factor, mi_after=after, mi_before=before, better=better))
        self.log("    Better alignment than before" if better else "    Worse alignment than before")
        return better

    def _rejected(self, lvl: int, m_flow, L: LevelLayout, Ln: LevelLayout, device) -> torch.Tensor:
        """The accumulated flow handed to level lvl+1 when level lvl is rejected (optflow_registrator.py:160-171)."""
        if lvl == 0:
            return torch.zeros((Ln.h, Ln.w, 2), dtype=torch.float32, device=device)
        return self.pyr_up(m_flow, L, Ln, 4.0)

    def _merge(self, m_flow: torch.Tensor, this_flow: torch.Tensor, L: LevelLayout) -> torch.Tensor:
        out = torch.empty_like(this_flow)
        if not self.corrected:   # merge_two_flows per tile (optflow_registrator.py:37-47, 217-240), quirks included
            return ops.merge_flows_tile_rows(m_flow, this_flow, self.T, self.ov, L.tile_rows[L.rank], out)
        if L.sharded:            # p - f2(p) may leave the band: the opt-in mode simply gathers the accumulated flow
            self.comm.gather_rows(m_flow, L.bands)
        return ops.compose_flows_rows(m_flow, this_flow, L.band, out)

    def _upscale_to_full(self, flow: torch.Tensor, L: LevelLayout, full: LevelLayout, factor: int):
        """_upscale_flow_to_full_res (optflow_registrator.py:204-215): NOT scaled by 2 (quirk Q2)."""
        if abs(flow.shape[0] - full.h) <= 1:
            return flow, L
        num_lvls = int(log2(factor))
        out, lay = flow, L
        for i in range(num_lvls):
            if i == num_lvls - 1:
                out, lay = self.pyr_up(flow, L, full, 2.0 if self.corrected else 1.0), full
            else:  # unreachable for contiguous factors; kept for parity with the reference's loop
                mid = LevelLayout(2 * flow.shape[0], 2 * flow.shape[1], self.T, self.ov, self.comm)
                out, lay = self.pyr_up(flow, L, mid, 1.0), mid
        return out, lay

    # ------------------------------------------------------------------ warp()
    def warp(self, img: torch.Tensor, flow: torch.Tensor, gather: bool = True) -> torch.Tensor:
        """Warper.warp (warper.py:37-76); with several ranks each warps its band and the image is gathered
        (gather=False: the result stays valid on this rank's band, Engine.warp_band(), only)."""
        L = LevelLayout(img.shape[0], img.shape[1], self.T, self.ov, self.comm)
        if self.comm.world > 1 and L.ny >= 2:
            tr = self.comm.tile_row_bands(L.ny)
            bands = [_clip(a * self.T, b * self.T, L.h) for a, b in tr]
            out = torch.empty_like(img)
            # the band is warped in a few pieces and every finished piece starts travelling while the next one is computed
            pieces = self.gather_pieces if gather else 1
            pending = []
            for k in range(pieces):
                sub = [(a + (b - a) * k // pieces, a + (b - a) * (k + 1) // pieces) for a, b in bands]
                with self.phase("warp"):
                    ops.warp_tiles_rows(img, flow, self.T, self.ov, sub[self.comm.rank], out)
                if gather:
                    pending.append(self.comm.gather_rows(out, sub, wait=False))
            with self.phase("gather image"):
                for p in pending:
                    p.wait()
            return out
        return ops.warp_tiles(img, flow, self.T, self.ov)

    # ------------------------------------------------------------------ host images in, host arrays out
    def result_rows(self, layout: Optional[LevelLayout], h: int) -> Range:
        """Rows of a result this rank delivers to the host: its band where the result is sharded, else an even share of
        the replicated result (every rank holds all of it; the download is still spread over all PCIe links)."""
        if layout is not None and layout.sharded:
            return layout.band
        return parallel.split_even(h, self.comm.world)[self.comm.rank]

    def warp_band(self, shape) -> Range:
        """Rows of the warped image this rank produces in warp(): its band of tile rows, or everything when warp() is
        not sharded."""
        L = LevelLayout(shape[0], shape[1], self.T, self.ov, self.comm)
        if self.comm.world > 1 and L.ny >= 2:
            a, b = self.comm.tile_row_bands(L.ny)[self.comm.rank]
            return _clip(a * self.T, b * self.T, L.h)
        return (0, L.h)

    def full_input_rows(self, shape) -> Range:
        """Rows of the full-resolution ref / mov images this rank reads in register() with the band-local pyramid: the
        support of its share of the first pyramid level, and its share of the full-resolution level if that is used."""
        full = LevelLayout(shape[0], shape[1], self.T, self.ov, self.comm)
        if self.comm.world == 1 or not self.local_pyramid:
            return (0, full.h)
        shapes, _, req, _ = self.pyramid_plan(tuple(shape))
        need = None
        if shapes:
            a, b = req[0]
            if b > a:
                need = _clip(2 * a - 2, 2 * b + 2, full.h)
        if self.full_res:
            r = full.input_rows(self.use_dog)
            if r[1] > r[0]:
                need = r if need is None else _union(need, r)
        return need if need is not None else (0, 0)

    def register_host(self, ref: np.ndarray, mov: np.ndarray, device=None):
        """register() for HOST images, the path behind the drop-in numpy API on 1..N GPUs.

        Every rank moves only its own share over its own PCIe link: it uploads the rows of ref / mov it reads
        (full_input_rows; everything on one GPU) on a copy stream -- the pyramid of ref is built while mov is still
        arriving -- keeps the flow sharded on the devices, and its rows of the result travel to the host on a side
        stream as soon as they are final (speculatively behind the last level's Farneback on one GPU).  The result is one
        full (H, W, 2) array: page-locked memory on one GPU, node-shared memory mapped by all ranks otherwise (complete
        on every rank when the call returns).  Returns (flow array, device flow valid on this rank's rows +- overlap)."""
        comm = self.comm
        self.gather_flow = False
        rows_in = self.full_input_rows(ref.shape)
        ref_d, ev_r = ops.upload_rows(ref, rows_in, device)
        mov_d, ev_m = ops.upload_rows(mov, rows_in, device)
        shape = tuple(ref.shape) + (2,)
        if comm.world == 1:
            out = ops.host_result(shape, np.float32)
        else:      # node-shared block, recycled from call to call: this rank's rows of it stay page-locked
            out = comm.shared_host_empty(shape, np.float32)
            ops.pin_rows(out, self.result_rows(LevelLayout(shape[0], shape[1], self.T, self.ov, comm), shape[0]))
        sink = ops.HostSink(out)
        m_flow = self.register(ref_d, mov_d, sink=sink, ready=(ev_r, ev_m))
        sink.wait()
        comm.barrier()            # node-shared result: complete once every rank's rows have landed
        return out, m_flow

    def warp_host(self, img: np.ndarray, flow: torch.Tensor):
        """warp() for a host image and a device flow valid on this rank's band +- overlap (register_host's, or an
        uploaded one): uploads the rows this rank's tile windows read, warps its band, downloads it into the shared
        result.  Several ranks only -- one GPU streams tile rows through ops.warp_tiles_host_streamed."""
        comm = self.comm
        band = self.warp_band(img.shape)
        L = LevelLayout(img.shape[0], img.shape[1], self.T, self.ov, comm)
        if L.sharded and L.ny < 2:      # a single row of tiles: warp() is replicated, so it needs the whole flow
            comm.gather_rows(flow, L.bands)
        sharded = comm.world > 1 and L.ny >= 2
        rows = band if sharded else self.result_rows(None, img.shape[0])
        if comm.world > 1:
            out = comm.shared_host_empty(img.shape, img.dtype)
            ops.pin_rows(out, rows)
        else:
            out = ops.host_result(img.shape, img.dtype)
        if sharded:      # tile rows of the band stream up / through the kernel / down on three streams
            ops.warp_tiles_host_streamed(img, flow, self.T, self.ov, rows=band, out=out)
        else:
            img_d, ev = ops.upload_rows(img, (0, img.shape[0]), flow.device)
            ops.wait_upload(ev)
            out_d = self.warp(img_d, flow, gather=False)
            sink = ops.HostSink(out)
            sink.push(out_d, rows)
            sink.wait()
        comm.barrier()
        return out

    def flow_rows_needed(self, shape) -> Range:
        """Rows of a host flow this rank has to upload for warp_host(): its warp band +- overlap."""
        band = self.warp_band(shape)
        return _clip(band[0] - self.ov, band[1] + self.ov, shape[0]) if band[1] > band[0] else (0, 0)
