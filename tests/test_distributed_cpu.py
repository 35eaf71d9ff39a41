"""CPU, world_size 2 over gloo: the host-side sharding / gathering logic of the multi-GPU LRT
(the data path has no collective; only the per-simulation results are all-gathered)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mind_the_gaps_b200 import lrt


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 10000, 10001):
        for world in (1, 2, 3, 8):
            spans = [lrt.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        a, b = lrt.shard_range(n_total, rank, world)
        rows = np.stack([np.arange(a, b, dtype=float), 10.0 * np.arange(a, b), np.full(b - a, rank, dtype=float)], axis=1)
        full = lrt.all_gather_rows(dist, rows, n_total)
        q.put((rank, full))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [11, 64])
def test_all_gather_rows_world2_gloo(n_total):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in range(world):
        full = got[r]
        assert full.shape == (n_total, 3)
        assert np.array_equal(full[:, 0], np.arange(n_total)) and np.array_equal(full[:, 1], 10.0 * np.arange(n_total))
        owner = np.concatenate([np.full(b - a, k) for k, (a, b) in enumerate(lrt.shard_range(n_total, k, world) for k in range(world))])
        assert np.array_equal(full[:, 2], owner)
