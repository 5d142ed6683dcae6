"""CPU: the row-band partition / halo-exchange logic of microaligner_b200.parallel on a world of 2-3 gloo ranks."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from microaligner_b200 import parallel


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_split_and_plan():
    assert parallel.split_even(20, 8) == [(0, 3), (3, 6), (6, 9), (9, 12), (12, 14), (14, 16), (16, 18), (18, 20)]
    assert parallel.split_even(3, 4) == [(0, 1), (1, 2), (2, 3), (3, 3)]
    owned = [(0, 10), (10, 20), (20, 25)]
    need = [(0, 13), (7, 22), (18, 25)]
    plan = parallel.transfer_plan(owned, need)
    assert sorted(plan) == sorted([(1, 0, (10, 13)), (0, 1, (7, 10)), (2, 1, (20, 22)), (1, 2, (18, 20))])
    assert parallel.transfer_plan([(0, 9)] * 3, [(0, 9)] * 3) == []     # replicated: nothing to move


@pytest.mark.parametrize("h,w,T,world", [(20000, 20000, 1000, 8), (2500, 3100, 1000, 2), (4321, 777, 500, 3), (5000, 5000, 1000, 4)])
def test_chunk_ownership_is_a_partition(h, w, T, world):
    """Every NMI chunk is owned by exactly one band and never runs more than ceil(T^2/w) rows past it."""
    ny = -(-h // T)
    bands = [(a * T, min(b * T, h)) for a, b in parallel.split_even(ny, world)]
    n, chunk = h * w, T * T
    nchunks = -(-n // chunk)
    seen = np.zeros(nchunks, int)
    over = -(-chunk // w) + 1
    for b in bands:
        c0, c1 = parallel.chunk_range_of_band(b, w, chunk, n)
        seen[c0:c1] += 1
        for c in range(c0, c1):
            assert c * chunk // w >= b[0]
            assert (min((c + 1) * chunk, n) - 1) // w < b[1] + over
    assert (seen == 1).all()


def _worker(rank, world, port, h, tmp):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    comm = parallel.init(dist.group.WORLD)
    try:
        truth = torch.arange(h * 6, dtype=torch.float32).reshape(h, 3, 2)
        bands = [(a * 10, min(b * 10, h)) for a, b in comm.tile_row_bands(-(-h // 10))]
        mine = bands[rank]
        t = torch.full_like(truth, -1.0)
        t[mine[0]:mine[1]] = truth[mine[0]:mine[1]]
        need = [(max(a - 4, 0), min(b + 7, h)) for a, b in bands]
        comm.exchange_rows(t, bands, need)
        a, b = need[rank]
        assert torch.equal(t[a:b], truth[a:b])
        if a > 0:
            assert (t[:a] == -1).all()
        # uint16 rides the wire as bytes
        u = torch.zeros((h, 5), dtype=torch.uint16)
        u[mine[0]:mine[1]] = 40000 + rank
        comm.gather_rows(u, bands)
        for q, (x, y) in enumerate(bands):
            assert (u[x:y].to(torch.int32) == 40000 + q).all()
        mm = torch.tensor([float(rank + 1), float(10 * (rank + 1))])
        comm.allreduce_minmax(mm)
        assert mm.tolist() == [1.0, 10.0 * world]
        s = torch.zeros(world, dtype=torch.float64)
        s[rank] = 0.25 * (rank + 1)
        comm.allreduce_sum(s)
        assert s.tolist() == [0.25 * (q + 1) for q in range(world)]
        open(os.path.join(tmp, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()
        parallel.init(None)


@pytest.mark.parametrize("world,h", [(2, 57), (3, 95)])
def test_exchange_rows_gloo(tmp_path, world, h):
    mp.spawn(_worker, args=(world, free_port(), h, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))
