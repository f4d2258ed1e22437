"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink/NVSwitch).

The reference clusters over TCP: every node holds full replicas of U and V, solved row
ranges are streamed to all peers ('alsSaveCalcedFactors', EmfMaster.js:711-723) and RMSE
partial sums are reduced on the lord ('rmseSaveCalcs', EmfMaster.js:726-736).  Here:

  * portions are split into contiguous nnz-balanced slices per rank (balanced_cuts),
  * after a half-step every rank broadcasts its solved row range of the replica
    (broadcast_ranges: an all-gather with unequal counts), or — fused variant — the solve
    kernels store rows straight into peer replicas (connect_peers + barrier),
  * RMSE sums are combined in rank order, the last portion's partials travel with them (Q7).

The same functions run on CPU tensors with the gloo backend (tests/test_dist_gloo.py).
"""
import os

import numpy as np


def balanced_cuts(ends, world):
    """ends[p] = cumulative ratings up to and including portion p (non-decreasing).
    Returns cuts[world+1] (portion indices) so that rank r takes [cuts[r], cuts[r+1])."""
    ends = np.asarray(ends, np.int64)
    n = len(ends)
    cuts = np.zeros(world + 1, np.int64)
    cuts[world] = n
    total = int(ends[-1]) if n else 0
    for g in range(1, world):
        target = (total * g) // world
        c = int(np.searchsorted(ends, target, side="left")) + 1 if total else 0
        # choose the boundary closer to the target
        if c - 1 >= 1 and abs(int(ends[c - 2]) - target) <= abs(int(ends[min(c, n) - 1]) - target):
            c -= 1
        cuts[g] = min(max(c, cuts[g - 1]), n)
    return cuts


def row_cost(n, k, dual_max=96):
    """Estimated device time of one solved row with n ratings (arbitrary units: ns on one B200 at k = 100),
    fitted to the measured per-class times: rows of up to dual_max ratings solve the n x n dual
    system (gather + Gram ~ n k, factorisation ~ n^3), longer rows pay the tensor-core Gram per rating plus one
    k x k factorisation — SURVEY.md §8(e): balance on n k^2 + k^3 / 3 rather than on raw nnz when rows are short."""
    n = np.asarray(n, np.float64)
    s = k / 100.0
    dual = 2.0 + 0.20 * s * n + 8.5e-5 * n ** 3
    primal = 46.0 * s ** 3 + 0.107 * s * s * n
    return np.where(n <= 0, 0.0, np.where(n <= dual_max, dual, primal))


def cost_ends(row_counts, portions_row_id_to, k, dual_max=96):
    """Cumulative row_cost at every portion end (int64, for balanced_cuts)."""
    c = np.cumsum(row_cost(row_counts, k, dual_max))
    pto = np.asarray(portions_row_id_to, np.int64)
    return np.round(c[pto - 1] * 16.0).astype(np.int64) if len(pto) else np.zeros(0, np.int64)


def init_from_env(backend=None):
    """torchrun contract: RANK / LOCAL_RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT."""
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, local, world


def all_ranges(my_range, world, group=None):
    """Every rank's solved row-id range [a, b)."""
    if world == 1:
        return [tuple(my_range)]
    import torch.distributed as dist
    out = [None] * world
    dist.all_gather_object(out, tuple(int(x) for x in my_range), group=group)
    return out


_grouped_ok = True


def broadcast_ranges(mat, ranges, group=None):
    """mat: 2-D torch tensor replica (rows x k) on every rank; rank r owns rows ranges[r].
    All-gather with unequal counts, in place: the output list is views of the replica itself (NCCL runs
    it as one group of broadcasts); ranks with an empty range fall back to one broadcast per range."""
    import torch.distributed as dist
    rank = dist.get_rank(group)
    global _grouped_ok
    if _grouped_ok and all(b > a for a, b in ranges) and dist.get_backend(group) == "nccl":
        views = [mat[a:b] for a, b in ranges]
        try:
            dist.all_gather(views, views[rank], group=group)
            return
        except (RuntimeError, ValueError, TypeError):   # argument check of an older torch: nothing was enqueued
            _grouped_ok = False
    works = []
    for r, (a, b) in enumerate(ranges):
        if b > a:
            works.append(dist.broadcast(mat[a:b], src=r, group=group, async_op=True))
    for w in works:
        w.wait()


class _DevArray:
    """__cuda_array_interface__ view of a raw device pointer (no ownership)."""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (int(ptr), False),
                                         "version": 3, "strides": None}


def device_tensor(ctx, which, rows, k, device):
    import torch
    return torch.as_tensor(_DevArray(ctx.device_factors_ptr(which), (rows, k)), device=torch.device("cuda", device))


def refresh_replicas(ctx, which, k, ranges, rank, group=None, host=None):
    """After a half-step: make replica `which` identical on all ranks."""
    import torch
    rows = ctx.total_users if which == 0 else ctx.total_items
    ctx.synchronize()                                  # library stream -> done
    dev = torch.cuda.current_device()
    t = device_tensor(ctx, which, rows, k, dev)
    broadcast_ranges(t, ranges, group)
    torch.cuda.synchronize()
    if host is not None:                               # per-portion mode: host segment is the truth
        for r, (a, b) in enumerate(ranges):
            if r != rank and b > a:
                ctx.download_factors(which, a, b - a)


def connect_peers(ctx, which, rank, world, group=None):
    """Fused all-gather: exchange CUDA IPC handles of replica `which` and register the peers,
    so the solve kernels store each solved row into every replica (NVLink peer stores)."""
    import torch.distributed as dist
    handles = [None] * world
    dist.all_gather_object(handles, ctx.ipc_export(which), group=group)
    ptrs = [ctx.ipc_import(h) for r, h in enumerate(handles) if r != rank]
    ctx.set_peers(which, ptrs)
    return ptrs


def node_shared_matrices(tag, shapes, rank, init_fn=None, group=None):
    """One set of float32 matrices shared by all ranks of the box through /dev/shm, the way every worker of a
    reference node maps the same SysV segments (EmfBase.js:403-412 create, 430-450 open): rank 0 creates the
    files and runs init_fn(mats), everybody maps them after a barrier, then the names are unlinked (the
    mappings stay).  Returns the list of np.memmap arrays."""
    import numpy as np
    paths = ["/dev/shm/ycnr_%s_%d" % (tag, i) for i in range(len(shapes))]
    mats = None
    if rank == 0:
        mats = [np.memmap(p, np.float32, "w+", shape=tuple(sh)) for p, sh in zip(paths, shapes)]
        if init_fn is not None:
            init_fn(mats)
        for m_ in mats:
            m_.flush()
    barrier(group)
    if rank != 0:
        mats = [np.memmap(p, np.float32, "r+", shape=tuple(sh)) for p, sh in zip(paths, shapes)]
    barrier(group)
    if rank == 0:            # every rank holds its mapping now: the names can go
        for p in paths:
            os.unlink(p)
    return mats


def barrier(group=None):
    import torch.distributed as dist
    dist.barrier(group=group)


_stream_cache = {}


def stream_barrier(ctx, group=None):
    """Device-side barrier on the library stream: a one-element NCCL all-reduce enqueued behind everything the
    context has launched; the stream continues when every rank has got there.  The host does not wait, so it
    queues the next half-step while the barrier is in flight (a host synchronize + dist.barrier() pair cost
    two host round trips per half-step)."""
    import torch
    import torch.distributed as dist
    dev = torch.cuda.current_device()
    key = (id(ctx), dev)
    if key not in _stream_cache:
        _stream_cache.clear()
        _stream_cache[key] = (torch.cuda.ExternalStream(ctx.stream_ptr(), device=torch.device("cuda", dev)),
                              torch.zeros(1, device=torch.device("cuda", dev)))
    st, t = _stream_cache[key]
    with torch.cuda.stream(st):
        dist.all_reduce(t, group=group)


def all_gather_doubles(values, group=None):
    """Every rank's list of doubles, in rank order (one all-gather; device tensors under NCCL, CPU under gloo)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    mine = torch.tensor([float(v) for v in values], dtype=torch.float64, device=dev)
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine, group=group)
    return torch.stack(parts).cpu().tolist()


def reduce_rmse(sums, last, group=None):
    """Combine (rSumDiff2, rCnt, rSum) over ranks in rank order; `last` = partials of the globally
    last portion = the highest rank that had any portion ('rmseSaveCalcs', EmfMaster.js:726-736,770-783).
    One all-gather of six doubles per rank (device tensors under NCCL, CPU tensors under gloo)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    mine = torch.tensor([float(sums[0]), float(sums[1]), float(sums[2]),
                         0.0 if last is None else float(last["rSum"]),
                         0.0 if last is None else float(last["rCnt"]),
                         0.0 if last is None else 1.0], dtype=torch.float64, device=dev)
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine, group=group)
    rows = torch.stack(parts).cpu().tolist()
    tot = [0.0, 0.0, 0.0]
    glast = None
    for r in rows:                      # rank order: deterministic sums
        for i in range(3):
            tot[i] += r[i]
        if r[5] != 0.0:
            glast = {"rSum": r[3], "rCnt": r[4]}
    return tuple(tot), glast
