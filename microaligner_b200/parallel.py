"""Row-band sharding of the registration path across the GPUs of one node.

One process per GPU (torchrun); `init(group)` installs the process group (NCCL over NVLink on the GPU
box, gloo in the CPU tests).  The path shards by tile rows: every stage computes the
image rows it owns, halo rows are exchanged point-to-point, the few global scalars (DoG min/max, NMI
chunk scores) are all-reduced, and the final flow / image is gathered band by band.  With a world of 1
every function here is a no-op, so the single-GPU path runs the very same engine code."""
import os
import tempfile
import weakref
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist

Range = Tuple[int, int]


def split_even(n: int, parts: int) -> List[Range]:
    """Contiguous, as-even-as-possible partition of range(n) into `parts` (possibly empty) ranges."""
    base, rem = divmod(n, parts)
    out, a = [], 0
    for r in range(parts):
        b = a + base + (1 if r < rem else 0)
        out.append((a, b))
        a = b
    return out


def intersect(a: Range, b: Range) -> Range:
    lo, hi = max(a[0], b[0]), min(a[1], b[1])
    return (lo, hi) if hi > lo else (lo, lo)


def subtract(a: Range, b: Range) -> List[Range]:
    """a minus b as up to two ranges."""
    out = []
    if a[1] <= a[0]:
        return out
    i = intersect(a, b)
    if i[1] <= i[0]:
        return [a]
    if a[0] < i[0]:
        out.append((a[0], i[0]))
    if i[1] < a[1]:
        out.append((i[1], a[1]))
    return out


def transfer_plan(owned: Sequence[Range], need: Sequence[Range]) -> List[Tuple[int, int, Range]]:
    """(src, dst, rows) for every block of rows rank `dst` needs but does not own. `owned` must be disjoint."""
    plan = []
    for dst, nd in enumerate(need):
        for miss in subtract(nd, owned[dst]):
            for src, ow in enumerate(owned):
                if src == dst:
                    continue
                rows = intersect(miss, ow)
                if rows[1] > rows[0]:
                    plan.append((src, dst, rows))
    return plan


Rect = Tuple[int, int, int, int]  # y0, y1, x0, x1


def tile_range_rects(tiles: Range, nx: int, T: int, h: int, w: int) -> List[Rect]:
    """Centres of the row-major tile range [tiles) as at most three rectangles (head of the first tile row,
    the full rows in between, tail of the last tile row)."""
    t0, t1 = tiles
    out = []
    while t0 < t1:
        i, j = divmod(t0, nx)
        if j == 0 and t1 - t0 >= nx:                       # whole tile rows
            rows = (t1 - t0) // nx
            out.append((i * T, min((i + rows) * T, h), 0, w))
            t0 += rows * nx
        else:                                              # partial tile row
            j1 = min(nx, j + (t1 - t0))
            out.append((i * T, min((i + 1) * T, h), j * T, min(j1 * T, w)))
            t0 += j1 - j
    return out


def rect_plan(have: Sequence[Sequence[Rect]], need_rows: Sequence[Range]):
    """(src, dst, rect): the part of every rectangle rank `src` holds that rank `dst` needs (full-width row
    range need_rows[dst]) and does not hold itself."""
    plan = []
    for dst, nd in enumerate(need_rows):
        if nd[1] <= nd[0]:
            continue
        for src, rects in enumerate(have):
            if src == dst:
                continue
            for (y0, y1, x0, x1) in rects:
                rows = intersect((y0, y1), nd)
                if rows[1] <= rows[0]:
                    continue
                # drop what dst computed itself
                pieces = [(rows[0], rows[1], x0, x1)]
                for (a0, a1, b0, b1) in have[dst]:
                    nxt = []
                    for (p0, p1, q0, q1) in pieces:
                        ry, rx = intersect((p0, p1), (a0, a1)), intersect((q0, q1), (b0, b1))
                        if ry[1] <= ry[0] or rx[1] <= rx[0]:
                            nxt.append((p0, p1, q0, q1))
                            continue
                        for (u0, u1) in subtract((p0, p1), ry):
                            nxt.append((u0, u1, q0, q1))
                        for (v0, v1) in subtract((q0, q1), rx):
                            nxt.append((ry[0], ry[1], v0, v1))
                    pieces = nxt
                plan.extend((src, dst, r) for r in pieces)
    return plan


def chunk_range_of_band(band: Range, w: int, chunk: int, n: int) -> Range:
    """NMI chunks (runs of `chunk` row-major elements of an image of width w, n elements in total)
    whose first element lies in image rows [band)."""
    c0 = -(-(band[0] * w) // chunk)
    c1 = -(-(min(band[1] * w, n)) // chunk)
    return (c0, max(c0, c1))


class _Done:
    def wait(self):
        pass


class _Pending:
    def __init__(self, reqs):
        self.reqs = reqs

    def wait(self):
        for r in self.reqs:
            r.wait()


class Comm:
    def __init__(self, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if group is not None else 1
        self.rank = dist.get_rank(group) if group is not None else 0
        self.backend = dist.get_backend(group) if group is not None else None
        # ranks in this module are GROUP ranks; P2POp / broadcast address peers by GLOBAL rank
        self._global = [dist.get_global_rank(group, r) for r in range(self.world)] if group is not None else [0]
        self._sign = {}      # device -> [-1, 1] (allreduce_minmax)

    # -- layout ---------------------------------------------------------------------------------
    def tile_row_bands(self, ny: int) -> List[Range]:
        """Tile rows per rank for the row-wise stages: boundaries at round(r * ny / world), i.e. as close as whole tile
        rows get to the equal shares of TILES the Farneback stage works on (LevelLayout.fb_tiles) -- what has to move
        between the two partitions after every Farneback call is at most half a tile row per boundary."""
        cuts = [(r * ny + self.world // 2) // self.world for r in range(self.world + 1)]
        return [(cuts[r], cuts[r + 1]) for r in range(self.world)]

    # -- data movement --------------------------------------------------------------------------
    def _wire(self, t: torch.Tensor) -> torch.Tensor:
        # NCCL (torch) has no 16-bit integer type: uint16 images travel as bytes, row slicing is unaffected
        return t.view(torch.uint8) if t.dtype == torch.uint16 else t

    def exchange_rows(self, t: torch.Tensor, owned: Sequence[Range], need: Sequence[Range], wait: bool = True):
        """Make rows need[rank] of `t` valid on this rank, given that rank q holds rows owned[q].
        wait=False returns a handle whose .wait() completes the exchange: the transfers run on the communication stream
        while the caller keeps launching kernels (that must not touch the rows in flight)."""
        if self.world == 1:
            return t if wait else _Done()
        plan = transfer_plan(owned, need)
        mine = [p for p in plan if self.rank in (p[0], p[1])]
        if not mine:
            return t if wait else _Done()
        tw = self._wire(t)
        stage_cpu = self.backend == "gloo" and t.is_cuda
        ops, recvs = [], []
        for src, dst, (a, b) in mine:
            view = tw[a:b]
            if src == self.rank:
                buf = view.cpu() if stage_cpu else view
                ops.append(dist.P2POp(dist.isend, buf, self._global[dst], group=self.group))
            else:
                buf = torch.empty(view.shape, dtype=view.dtype, device="cpu") if stage_cpu else view
                ops.append(dist.P2POp(dist.irecv, buf, self._global[src], group=self.group))
                if stage_cpu:
                    recvs.append((view, buf))
        reqs = dist.batch_isend_irecv(ops)
        if not wait and not recvs:
            return _Pending(reqs)
        for req in reqs:
            req.wait()
        for view, buf in recvs:
            view.copy_(buf)
        return t if wait else _Done()

    def exchange_rects(self, t: torch.Tensor, have: Sequence[Sequence[Rect]], need_rows: Sequence[Range]):
        """Rank q holds the rectangles have[q] of `t` (rows x columns); make rows need_rows[rank] valid here."""
        if self.world == 1:
            return t
        mine = [p for p in rect_plan(have, need_rows) if self.rank in (p[0], p[1])]
        if not mine:
            return t
        tw = self._wire(t)
        scale = tw.shape[1] // t.shape[1]                  # uint16 images travel as bytes: 2 columns per pixel
        cpu = self.backend == "gloo" and t.is_cuda
        ops, recvs = [], []
        for src, dst, (y0, y1, x0, x1) in mine:
            view = tw[y0:y1, x0 * scale:x1 * scale]
            if src == self.rank:
                buf = view.contiguous()
                ops.append(dist.P2POp(dist.isend, buf.cpu() if cpu else buf, self._global[dst], group=self.group))
            else:
                direct = view.is_contiguous() and not cpu
                buf = view if direct else torch.empty(view.shape, dtype=view.dtype, device="cpu" if cpu else view.device)
                ops.append(dist.P2POp(dist.irecv, buf, self._global[src], group=self.group))
                if not direct:
                    recvs.append((view, buf))
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        for view, buf in recvs:
            view.copy_(buf)
        return t

    def gather_rows(self, t: torch.Tensor, owned: Sequence[Range], wait: bool = True):
        """Every rank ends up with all rows of `t` (rank q contributes rows owned[q])."""
        if self.world == 1:
            return t if wait else _Done()
        full = (0, t.shape[0])
        return self.exchange_rows(t, owned, [full] * self.world, wait=wait)

    def allreduce_minmax(self, mm: torch.Tensor):
        """mm = (..., 2) float32 [min, max] pairs (device); reduced in place over the group, one collective."""
        if self.world == 1:
            return mm
        sign = self._sign.get(mm.device)
        if sign is None:
            sign = self._sign[mm.device] = torch.tensor([-1.0, 1.0], dtype=torch.float32, device=mm.device)
        v = self._allreduce(mm * sign, dist.ReduceOp.MAX)          # max(-min) = -min(min): one collective for both
        torch.mul(v, sign, out=mm)
        return mm

    def _allreduce(self, t: torch.Tensor, op):
        if self.backend == "gloo" and t.is_cuda:      # CPU-staged: only the 1-GPU multi-rank tests take this path
            c = t.cpu()
            dist.all_reduce(c, op=op, group=self.group)
            t.copy_(c)
        else:
            dist.all_reduce(t, op=op, group=self.group)
        return t

    def allreduce_sum(self, t: torch.Tensor):
        if self.world > 1:
            self._allreduce(t, dist.ReduceOp.SUM)
        return t

    def broadcast(self, t: torch.Tensor, src: int = 0):
        if self.world > 1:
            dist.broadcast(self._wire(t), src=self._global[src], group=self.group)
        return t


    # -- node-shared host memory ----------------------------------------------------------------
    def barrier(self):
        if self.world > 1:
            dist.barrier(group=self.group)

    def shared_host_empty(self, shape, dtype) -> np.ndarray:
        """Host array of `shape` that every rank of the (single-node) group maps to the SAME pages (a file in /dev/shm
        created by rank 0 and unlinked once everybody has mapped it).  This is how the drop-in numpy API returns a full
        (H, W, 2) flow / (H, W) image on every rank although each rank downloads only its own band over its own PCIe
        link.  Collective: every rank must call it with the same arguments in the same order.  Blocks are recycled once
        the array handed out earlier is dead on every rank."""
        shape, dtype = tuple(int(v) for v in shape), np.dtype(dtype)
        if self.world == 1:
            return np.empty(shape, dtype)
        key = (shape, dtype.str)
        pool = self.__dict__.setdefault("_shared_pool", {}).setdefault(key, [])
        if pool:
            free = torch.tensor([1 if e["ref"] is None or e["ref"]() is None else 0 for e in pool], dtype=torch.int32)
            if self.backend == "nccl":
                free = free.cuda()
            dist.all_reduce(free, op=dist.ReduceOp.MIN, group=self.group)
            for e, ok in zip(pool, free.tolist()):
                if ok:
                    return self._hand_out(e)
        # rank 0 creates an anonymous memory file (memfd: RAM-backed like /dev/shm but not limited by the size of that
        # mount) and the other ranks open it through /proc; without memfd_create a file in /dev/shm (or the temp dir) does
        nbytes = max(int(np.prod(shape)) * dtype.itemsize, 1)
        name, fd = [None], None
        if self.rank == 0:
            seq = self.__dict__["_shared_seq"] = self.__dict__.get("_shared_seq", 0) + 1
            try:
                fd = os.memfd_create(f"microaligner_b200_{seq}")
                os.ftruncate(fd, nbytes)
                name[0] = f"/proc/{os.getpid()}/fd/{fd}"
            except (AttributeError, OSError):
                fd = None
                root = "/dev/shm" if os.path.isdir("/dev/shm") and _free_bytes("/dev/shm") > nbytes + (1 << 30) else tempfile.gettempdir()
                name[0] = os.path.join(root, f"microaligner_b200_{os.getpid()}_{seq}.shared")
                with open(name[0], "wb") as f:
                    f.truncate(nbytes)
        dist.broadcast_object_list(name, src=self._global[0], group=self.group)
        mm = np.memmap(name[0], dtype=dtype, mode="r+", shape=shape)
        self.barrier()
        if self.rank == 0:          # the mappings keep the pages alive
            if fd is not None:
                os.close(fd)
            else:
                os.unlink(name[0])
        entry = {"mm": mm, "ref": None}
        pool.append(entry)
        return self._hand_out(entry)

    @staticmethod
    def _hand_out(entry) -> np.ndarray:
        arr = np.asarray(entry["mm"]).view(np.ndarray)
        entry["ref"] = weakref.ref(arr)
        return arr


def _free_bytes(path: str) -> int:
    import shutil
    return shutil.disk_usage(path).free


_COMM = Comm(None)


def init(group=None) -> Comm:
    """Install the process group the engine shards over (None = single process)."""
    global _COMM
    _COMM = Comm(group)
    return _COMM


def get() -> Comm:
    return _COMM
