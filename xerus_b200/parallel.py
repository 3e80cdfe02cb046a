"""Multi-GPU partitioning of the path — only where it shards naturally.

* A single TTTensor::round / move_core / one-site ALS is a strictly sequential chain: replicas only (bench.py --gpus N).
* Batches of independent TT operations (BASELINE config 5: y_b = A x_b ; y_b.round(r)) shard by item index,
  item b -> rank b mod world, with NO collective on the data path; only the per-item summaries (ranks, norms) are
  gathered at the end.  One process per GPU, torch.distributed for the plumbing (nccl on GPUs, gloo in the CPU tests).
"""
import numpy as np


def shard_items(n_items, rank, world):
    """Indices of the items rank `rank` owns: b mod world == rank (round robin keeps the shards balanced to +-1)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return list(range(rank, n_items, world))


def gather_by_item(local, n_items, group=None):
    """All-gathers {item index: summary} dictionaries and returns the list ordered by item index on every rank.
    Checks that every item was produced exactly once."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        parts = [None] * dist.get_world_size(group)
        dist.all_gather_object(parts, local, group=group)
    else:
        parts = [local]
    merged = {}
    for p in parts:
        for k, v in p.items():
            if k in merged:
                raise RuntimeError("item %d was processed by more than one rank" % k)
            merged[k] = v
    missing = [b for b in range(n_items) if b not in merged]
    if missing:
        raise RuntimeError("items not processed by any rank: %s" % missing[:8])
    return [merged[b] for b in range(n_items)]


def matvec_round_batch(A, make_x, n_items, max_rank, rank=0, world=1, process=None, workers=1):
    """BASELINE config 5 on this rank's shard: for every owned item b, y_b = A x_b (operator application) followed by
    y_b.round(max_rank).  `make_x(b)` returns the item's TTTensor (so inputs are a pure function of the item index and
    do not depend on the partitioning); `process` may replace the default apply+round (used by the CPU tests).
    With workers > 1 the items are driven by that many host threads, each bound to its own library worker (CUDA stream),
    so several sweeps are in flight on the GPU at once.  Returns {b: (ranks, frob_norm)} for the owned items."""
    mine = shard_items(n_items, rank, world)

    def one(b):
        x = make_x(b)
        if process is not None:
            return b, process(A, x, max_rank)
        y = A.apply(x)
        y.round(int(max_rank))
        return b, (tuple(y.ranks()), float(y.frob_norm()))

    if workers <= 1 or process is not None:
        return dict(one(b) for b in mine)
    return dict(run_on_workers(one, mine, workers))


def run_on_workers(fn, items, workers):
    """Maps fn over items with `workers` host threads, thread i bound to library worker i+1 (worker 0 stays with the caller)."""
    import concurrent.futures
    import itertools
    import threading
    import xerus_b200 as xb
    xb.synchronize()                       # inputs produced on the caller's worker are complete before hand-over
    ids, lock = itertools.count(1), threading.Lock()

    def bind():
        with lock:
            w = next(ids)
        xb.worker_select(w)

    with concurrent.futures.ThreadPoolExecutor(max_workers=workers, initializer=bind) as ex:
        out = list(ex.map(fn, items))
    xb.synchronize_all()
    return out


def item_rng(seed, b):
    """Per-item random stream: independent of how items are distributed over ranks."""
    return np.random.default_rng([int(seed), int(b)])


# ---- bond-split environment application (BASELINE config 4) --------------------------------------------------------------
def slab_range(r, rank, world):
    """Contiguous slab of the right bond index owned by `rank`: balanced to +-1, disjoint, covering [0, r)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, extra = divmod(r, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def env_apply(L, A_cores, R, v, slab=None, out=None):
    """y = {L, A_1..A_s, R} v (matrix-free local operator of ALS/DMRG, SPD environments) on CUDA float64 torch tensors:
    L (l, a, l), A_p (a, m, n, b), R (r, b, r), v (l, n_1..n_s, r) -> y (l, m_1..m_s, r).  `slab` = (begin, end) restricts the
    contraction over the right bond to that slab (partial result of full size).  Enqueued on the library stream."""
    import ctypes as C
    import torch
    from ._lib import call
    for t in [L, R, v] + list(A_cores):
        if not (t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()):
            raise ValueError("env_apply needs contiguous CUDA float64 tensors")
    s = len(A_cores)
    l, r = L.shape[0], R.shape[0]
    if out is None:
        out = torch.empty((l,) + tuple(a.shape[1] for a in A_cores) + (r,), dtype=torch.float64, device=v.device)
    begin, end = slab if slab is not None else (0, r)
    ptrs = (C.c_void_p * s)(*[a.data_ptr() for a in A_cores])
    dims = (C.c_size_t * (4 * s))(*[int(x) for a in A_cores for x in a.shape])
    call("xb_env_apply", out.data_ptr(), L.data_ptr(), l, L.shape[1], ptrs, dims, s, R.data_ptr(), r, R.shape[1],
         v.data_ptr(), begin, end)
    return out


def bond_split_apply(L, A_cores, R, v, rank, world, out=None, group=None):
    """The application split along the right bond index over `world` GPUs: every rank contracts its slab, then ONE sum
    all-reduce (NCCL over NVLink) assembles y on all ranks.  Must run with the library stream current
    (torch.cuda.stream(ExternalStream(xb.stream_handle()))) so the collective is ordered after the kernels."""
    import torch.distributed as dist
    y = env_apply(L, A_cores, R, v, slab=slab_range(R.shape[0], rank, world), out=out)
    if world > 1:
        dist.all_reduce(y, op=dist.ReduceOp.SUM, group=group)
    return y


def row_split_apply(L, A_cores, R, v, rank, world, out=None, group=None):
    """The application split along the LEFT bond index (rows of the result) over `world` GPUs: every rank computes its row block
    (1/world of all three stages), then ONE all-gather (NCCL over NVLink) assembles y on all ranks — nothing is summed.  Must run
    with the library stream current, as bond_split_apply."""
    import ctypes as C
    import torch
    import torch.distributed as dist
    from ._lib import call
    s = len(A_cores)
    l, r = L.shape[0], R.shape[0]
    if l % world:
        raise ValueError("the left bond must split evenly over the ranks")
    if out is None:
        out = torch.empty((l,) + tuple(a.shape[1] for a in A_cores) + (r,), dtype=torch.float64, device=v.device)
    begin, end = slab_range(l, rank, world)
    ptrs = (C.c_void_p * s)(*[a.data_ptr() for a in A_cores])
    dims = (C.c_size_t * (4 * s))(*[int(x) for a in A_cores for x in a.shape])
    call("xb_env_apply_rows", out[begin:end].data_ptr(), L.data_ptr(), l, L.shape[1], ptrs, dims, s, R.data_ptr(), r, R.shape[1],
         v.data_ptr(), begin, end)
    if world > 1:
        dist.all_gather_into_tensor(out, out[begin:end], group=group)
    return out


def row_split_apply_fused(L, A_cores, R, v, px):
    """row_split_apply with the all-gather fused into the application over peer memory (xb_env_apply_rows_fused): the epilogue of the
    last GEMM stores this rank's row block into the result area of every rank over NVLink.  Returns a torch view of this rank's copy
    of y (valid in the library stream's order; overwritten by the next call on `px`)."""
    import ctypes as C
    import torch
    from ._lib import call
    s = len(A_cores)
    l, r = L.shape[0], R.shape[0]
    if l % px.world:
        raise ValueError("the left bond must split evenly over the ranks")
    begin, end = slab_range(l, px.rank, px.world)
    ptrs = (C.c_void_p * s)(*[a.data_ptr() for a in A_cores])
    dims = (C.c_size_t * (4 * s))(*[int(x) for a in A_cores for x in a.shape])
    sym = (C.c_void_p * px.world)(*px.ptrs)
    px.epoch_rows += 1
    y = C.c_void_p()
    call("xb_env_apply_rows_fused", L.data_ptr(), l, L.shape[1], ptrs, dims, s, R.data_ptr(), r, R.shape[1], v.data_ptr(), begin, end,
         px.rank, px.world, sym, px.epoch_rows, C.byref(y))
    shape = (l,) + tuple(a.shape[1] for a in A_cores) + (r,)

    class _Arr:                                              # __cuda_array_interface__ view of the result area (no copy)
        __cuda_array_interface__ = {"shape": shape, "typestr": "<f8", "data": (y.value, False), "version": 3, "strides": None}
    return torch.as_tensor(_Arr(), device=v.device)


class PeerExchange:
    """Symmetric peer-memory buffers for bond_split_apply_fused: every rank allocates one buffer, exports its CUDA IPC handle,
    and maps the buffers of the other ranks of the box (handles travel through torch.distributed).  With `group_size` ranks in
    ONE process (tests: the "ranks" are host threads on library workers of the same device) no IPC is involved."""

    def __init__(self, rows, cols, rank, world, dist=None, local_buffers=None):
        import ctypes as C
        from ._lib import call
        self.rank, self.world, self.epoch, self.epoch_rows = rank, world, 0, 0
        self.rows, self.cols = rows, cols
        nbytes = C.c_size_t()
        call("xb_peer_buffer_bytes", rows, cols, world, C.byref(nbytes))
        self._opened = []
        if local_buffers is not None:                     # same process: pointers are shared as they are
            self.mine = None
            self.ptrs = list(local_buffers)
        else:
            mine, handle = C.c_void_p(), C.create_string_buffer(64)
            call("xb_peer_buffer_create", nbytes.value, C.byref(mine), handle)
            self.mine = mine
            handles = [handle.raw]
            if world > 1:
                handles = [None] * world
                dist.all_gather_object(handles, handle.raw)
            self.ptrs = []
            for p in range(world):
                if p == rank:
                    self.ptrs.append(mine.value)
                else:
                    q = C.c_void_p()
                    call("xb_peer_buffer_open", handles[p], C.byref(q))
                    self._opened.append(q)
                    self.ptrs.append(q.value)
            if world > 1:
                dist.barrier()                            # every buffer is zeroed and mapped before the first remote store

    @staticmethod
    def allocate_local(rows, cols, world):
        """`world` buffers in this process (for the single-device emulation); returns the list of device pointers."""
        import ctypes as C
        from ._lib import call
        nbytes = C.c_size_t()
        call("xb_peer_buffer_bytes", rows, cols, world, C.byref(nbytes))
        out = []
        for _ in range(world):
            q, h = C.c_void_p(), C.create_string_buffer(64)
            call("xb_peer_buffer_create", nbytes.value, C.byref(q), h)
            out.append(q.value)
        return out

    def check(self):
        """Synchronises this worker's stream and raises XerusError if a bounded wait of the fused exchange gave up on this rank
        (a peer lagged by more than the wait or died): results since then are not valid (xb_peer_buffer_check)."""
        from ._lib import call
        call("xb_peer_buffer_check", self.ptrs[self.rank])

    def close(self, dist=None):
        from ._lib import call
        import ctypes as C
        if dist is not None and self.world > 1:
            dist.barrier()
        for q in self._opened:
            call("xb_peer_buffer_close", q)
        self._opened = []
        if self.mine is not None:
            call("xb_peer_buffer_destroy", self.mine)
            self.mine = None


def bond_split_apply_fused(L, A_cores, R, v, px):
    """bond_split_apply with the reduction fused into the application over peer memory (xb_env_apply_fused): the last GEMM's
    epilogue writes each rank's row block into that rank's buffer over NVLink, a reduce kernel sums the blocks in rank order and
    writes the sum to every rank.  Returns a torch view of this rank's copy of y (valid in the library stream's order; it is
    overwritten by the next call on `px`)."""
    import ctypes as C
    import torch
    from ._lib import call
    s = len(A_cores)
    l, r = L.shape[0], R.shape[0]
    begin, end = slab_range(r, px.rank, px.world)
    ptrs = (C.c_void_p * s)(*[a.data_ptr() for a in A_cores])
    dims = (C.c_size_t * (4 * s))(*[int(x) for a in A_cores for x in a.shape])
    sym = (C.c_void_p * px.world)(*px.ptrs)
    px.epoch += 1
    y = C.c_void_p()
    call("xb_env_apply_fused", L.data_ptr(), l, L.shape[1], ptrs, dims, s, R.data_ptr(), r, R.shape[1], v.data_ptr(), begin, end,
         px.rank, px.world, sym, px.epoch, C.byref(y))
    shape = (l,) + tuple(a.shape[1] for a in A_cores) + (r,)
    n = 1
    for d in shape:
        n *= d

    class _Arr:                                              # __cuda_array_interface__ view of the result area (no copy)
        __cuda_array_interface__ = {"shape": shape, "typestr": "<f8", "data": (y.value, False), "version": 3, "strides": None}
    return torch.as_tensor(_Arr(), device=v.device)
