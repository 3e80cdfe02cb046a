"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: item sharding and result gathering of the batch workload
(BASELINE config 5).  The per-item arithmetic is replaced by the CPU oracle here — this tests the partitioning, not the
kernels; on GPUs the same functions run with the device path (tests/test_gpu_tt.py::test_batch_sharding_single_rank)."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import tt_oracle as O
from xerus_b200 import parallel


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


N_ITEMS, D, N, R, TARGET = 7, 5, 3, 4, 3


def _make_x(b):
    return O.tt_random([N] * D, R, parallel.item_rng(1234, b))


def _process(A, x, max_rank):
    y = O.tt_apply(A, x)
    y.round(max_rank)
    return (tuple(y.ranks()), float(y.frob_norm()))


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        A = O.laplace_operator(D, N)
        local = parallel.matvec_round_batch(A, _make_x, N_ITEMS, TARGET, rank, world, process=_process)
        assert sorted(local) == parallel.shard_items(N_ITEMS, rank, world)
        full = parallel.gather_by_item(local, N_ITEMS)
        q.put((rank, full))
    finally:
        dist.destroy_process_group()


def test_shard_items_partition():
    for world in [1, 2, 3, 8]:
        seen = sorted(b for r in range(world) for b in parallel.shard_items(4096, r, world))
        assert seen == list(range(4096))
        sizes = [len(parallel.shard_items(4096, r, world)) for r in range(world)]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        parallel.shard_items(10, 2, 2)


def test_batch_world2_matches_single_process():
    A = O.laplace_operator(D, N)
    single = parallel.gather_by_item(parallel.matvec_round_batch(A, _make_x, N_ITEMS, TARGET, process=_process), N_ITEMS)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank in (0, 1):
        assert len(results[rank]) == N_ITEMS
        for (rk_a, nrm_a), (rk_b, nrm_b) in zip(results[rank], single):
            assert rk_a == rk_b and abs(nrm_a - nrm_b) <= 1e-12 * abs(nrm_b)


def test_gather_detects_missing_and_duplicate_items():
    with pytest.raises(RuntimeError):
        parallel.gather_by_item({0: 1, 2: 3}, 3)


def test_slab_range_partition():
    for r in [1, 7, 512, 513]:
        for world in [1, 2, 3, 8]:
            slabs = [parallel.slab_range(r, k, world) for k in range(world)]
            assert slabs[0][0] == 0 and slabs[-1][1] == r
            assert all(a[1] == b[0] for a, b in zip(slabs, slabs[1:]))
            sizes = [e - b for b, e in slabs]
            assert max(sizes) - min(sizes) <= 1


def _bond_worker(rank, world, port, q):
    """Bond-split local apply with a numpy stand-in for the device kernel: the slabs' partial results all-reduce (gloo)
    to the full application on every rank."""
    import torch
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(3)
        l, r, a, n = 5, 7, 2, 3
        L, R = rng.standard_normal((l, a, l)), rng.standard_normal((r, a, r))
        A1, v = rng.standard_normal((a, n, n, a)), rng.standard_normal((l, n, r))
        b, e = parallel.slab_range(r, rank, world)
        part = np.einsum("xay,ainb,zbw,ynw->xiz", L, A1, R[:, :, b:e], v[:, :, b:e])
        y = torch.from_numpy(part.copy())
        dist.all_reduce(y)
        full = np.einsum("xay,ainb,zbw,ynw->xiz", L, A1, R, v)
        q.put((rank, float(np.linalg.norm(y.numpy() - full) / np.linalg.norm(full))))
    finally:
        dist.destroy_process_group()


def test_bond_split_allreduce_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_bond_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert results[0] < 1e-13 and results[1] < 1e-13


def _row_worker(rank, world, port, q):
    """Row-split local apply (split along the LEFT bond: xb_env_apply_rows + all-gather) with a numpy stand-in for the device
    kernel: the ranks' row blocks all-gather (gloo) to the full application on every rank, bit for bit."""
    import torch
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(4)
        l, r, a, n = 6, 7, 2, 3
        L, R = rng.standard_normal((l, a, l)), rng.standard_normal((r, a, r))
        A1, v = rng.standard_normal((a, n, n, a)), rng.standard_normal((l, n, r))
        b, e = parallel.slab_range(l, rank, world)
        block = np.einsum("xay,ainb,zbw,ynw->xiz", L[b:e], A1, R, v)                  # rows [b, e) of the result: nothing to sum
        y = torch.empty(l, n, r, dtype=torch.float64)
        dist.all_gather_into_tensor(y, torch.from_numpy(np.ascontiguousarray(block)))
        full = np.einsum("xay,ainb,zbw,ynw->xiz", L, A1, R, v)
        q.put((rank, float(np.linalg.norm(y.numpy() - full) / np.linalg.norm(full)), bool(np.array_equal(y.numpy()[b:e], block))))
    finally:
        dist.destroy_process_group()


def test_row_split_allgather_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_row_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = {k: (err, same) for k, err, same in (q.get(timeout=120) for _ in range(2))}
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for k in (0, 1):
        assert results[k][0] < 1e-13 and results[k][1]
