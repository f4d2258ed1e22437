"""Host-side multi-rank logic on CPU: world_size 2 over gloo (rendezvous on 127.0.0.1)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from you_can_not_recommend_b200 import dist as ydist


def test_balanced_cuts_properties():
    rng = np.random.default_rng(0)
    for world in (1, 2, 4, 8):
        sizes = rng.integers(1, 1000, 97)
        ends = np.cumsum(sizes)
        cuts = ydist.balanced_cuts(ends, world)
        assert cuts[0] == 0 and cuts[-1] == len(ends) and (np.diff(cuts) >= 0).all()
        per = [int(sizes[cuts[r]:cuts[r + 1]].sum()) for r in range(world)]
        assert sum(per) == int(ends[-1])
        assert max(per) - min(per) <= 2 * sizes.max()
    assert list(ydist.balanced_cuts(np.asarray([5]), 4)) in ([0, 0, 0, 1, 1], [0, 1, 1, 1, 1], [0, 0, 1, 1, 1])
    assert list(ydist.balanced_cuts(np.zeros(0, np.int64), 2)) == [0, 0, 0]


def test_cost_model_cuts():
    """Row cost model behind the rank cuts: a row is not worth its nnz — a 1-rating row still costs
    a launch slot, the dual cost grows with n^3, every long row carries one k x k factorisation — and cuts made on
    the cumulative cost balance the cost, not the nnz."""
    c = ydist.row_cost(np.asarray([0, 1, 4, 48, 96, 97, 1000, 100000]), 100)
    assert c[0] == 0 and (np.diff(c[1:5]) > 0).all() and c[5] < c[4] * 1.2 and c[6] > c[5] and c[7] > 100 * c[5]
    assert c[1] * 96 > 2 * c[4] and c[4] > 4 * c[3]
    assert ydist.row_cost(np.asarray([200]), 256)[0] > 10 * ydist.row_cost(np.asarray([200]), 100)[0]
    rng = np.random.default_rng(1)
    counts = np.concatenate([rng.integers(1, 8, 5000), rng.integers(300, 900, 300)])      # short rows first, long last
    pto = np.arange(10, len(counts) + 1, 10)
    ends = ydist.cost_ends(counts, pto, 100)
    cuts = ydist.balanced_cuts(ends, 2)
    cost = ydist.row_cost(counts, 100)
    a = cost[:pto[cuts[1] - 1]].sum()
    assert abs(a - cost.sum() / 2) < 0.05 * cost.sum()
    nnz_cuts = ydist.balanced_cuts(np.cumsum(counts)[pto - 1], 2)
    assert nnz_cuts[1] != cuts[1]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    r, _, w = ydist.init_from_env("gloo")
    assert (r, w) == (rank, world)
    rows, k = 10, 4
    # every rank starts from the same replica and "solves" its own contiguous row range
    mat = torch.arange(rows * k, dtype=torch.float32).reshape(rows, k).clone()
    cuts = ydist.balanced_cuts(np.asarray([3, 4, 8, 10]), world)       # 4 portions -> row ends 3,4,8,10
    pto = [3, 4, 8, 10]
    lo, hi = int(cuts[rank]), int(cuts[rank + 1])
    a = 0 if lo == 0 else pto[lo - 1]
    b = a if hi == lo else pto[hi - 1]
    mat[a:b] = -(rank + 1)
    ranges = ydist.all_ranges((a, b), world)
    ydist.broadcast_ranges(mat, ranges)
    sums, last = ydist.reduce_rmse((1.0 + rank, 10.0 * (rank + 1), 0.5), {"rSum": float(rank), "rCnt": 2.0} if rank == 0 or hi > lo else None)
    gathered = ydist.all_gather_doubles([float(rank), 0.5, 10.0 * (rank + 1)])
    assert gathered == [[0.0, 0.5, 10.0], [1.0, 0.5, 20.0]]
    q.put((rank, ranges, mat.numpy().copy(), sums, last))
    dist.barrier()
    dist.destroy_process_group()


def _shared_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    ydist.init_from_env("gloo")

    def init(mats):
        mats[0][...] = 1.0
        mats[1][...] = 2.0

    U, V = ydist.node_shared_matrices("test_%d" % port, [(6, 3), (4, 3)], rank, init)
    first = (float(U.sum()), float(V.sum()))               # rank 0's initialisation is visible everywhere
    dist.barrier()                                          # (nobody writes before everybody has looked)
    a, b = (0, 3) if rank == 0 else (3, 6)
    U[a:b] = 10.0 * (rank + 1)                              # every rank writes back only the rows it solved
    dist.barrier()
    q.put((rank, first, np.array(U).copy(), os.path.exists("/dev/shm/ycnr_test_%d_0" % port)))
    dist.barrier()
    dist.destroy_process_group()


def test_node_shared_host_segments_gloo():
    """Per-portion mode with several ranks on one box: ONE pair of host factor segments (upstream: every worker
    of a node maps the same SysV segment, EmfBase.js:403-450); a rank's writes are seen by the others without
    any copy, and the /dev/shm names are gone once everybody has mapped them."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_shared_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=120) for _ in range(world)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, first, U, exists in res:
        assert first == (18.0, 24.0) and not exists
        assert (U[:3] == 10.0).all() and (U[3:] == 20.0).all()


def test_two_rank_exchange_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=120) for _ in range(world)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ranges = res[0][1]
    assert ranges == res[1][1] and ranges[0][0] == 0 and ranges[0][1] == ranges[1][0] and ranges[1][1] == 10
    assert (res[0][2] == res[1][2]).all()                       # replicas identical after the exchange
    for r, (a, b) in enumerate(ranges):
        assert (res[0][2][a:b] == -(r + 1)).all()
    assert res[0][3] == res[1][3] == (3.0, 30.0, 1.0)
    assert res[0][4] == res[1][4] == {"rSum": 1.0, "rCnt": 2.0}   # last portion lives on the highest rank
