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


def test_default_chunk_policy():
    """One chunk per shard up to 16384 simulations; refined runs above two waves of ensembles keep the pipeline."""
    for world in (1, 2, 4, 8):
        a, b = lrt.shard_range(10000, 0, world)
        assert lrt.default_chunk(b - a) == b - a                       # plain LRT: the whole shard in one chunk
    assert lrt.default_chunk(0) == 1 and lrt.default_chunk(100000) == 16384
    assert lrt.default_chunk(10000, pipeline=3, refine=True) == 3334   # refined, > 8192: ceil(n / pipeline) in 2048..8192
    assert lrt.default_chunk(40000, pipeline=3, refine=True) == 8192
    assert lrt.default_chunk(5000, pipeline=3, refine=True) == 5000    # refined, small shard: one chunk
    chunks = lambda n, c: len(range(0, n, c))
    assert chunks(10000, lrt.default_chunk(10000)) == 1 and chunks(10000, lrt.default_chunk(10000, 3, True)) == 3


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


def _worker_varlen(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mind_the_gaps_b200.gpmodelling import GPModelling
        n = 5 + 3 * rank                                         # replicas keep different numbers of draws
        g = GPModelling.__new__(GPModelling)                     # result container only: no device needed
        g._mcmc_samples = np.arange(n * 2, dtype=float).reshape(n, 2) + 100 * rank
        g._loglikelihoods = -np.arange(n, dtype=float) - 1000 * rank
        counts = g.gather_posteriors()
        q.put((rank, counts, g.mcmc_samples, g.loglikelihoods, g.max_loglikelihood, g.max_parameters))
    finally:
        dist.destroy_process_group()


def test_replica_posteriors_are_pooled_world2_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_varlen, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = {r[0]: r[1:] for r in (q.get(timeout=120) for _ in range(world))}
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    exp_s = np.concatenate([np.arange(10, dtype=float).reshape(5, 2), np.arange(16, dtype=float).reshape(8, 2) + 100])
    exp_l = np.concatenate([-np.arange(5, dtype=float), -np.arange(8, dtype=float) - 1000])
    for r in range(world):
        counts, samples, ll, best, best_p = got[r]
        assert counts == [5, 8]
        assert np.array_equal(samples, exp_s) and np.array_equal(ll, exp_l)
        assert best == 0.0 and np.array_equal(best_p, [0.0, 1.0])
