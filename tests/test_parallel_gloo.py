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
