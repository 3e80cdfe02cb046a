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
