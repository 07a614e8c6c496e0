"""CPU tests of the N>1 host logic with world_size 2 over gloo: pairs are partitioned over ranks, each rank produces the
expectation counts of its shard, one all-reduce gives every rank the total (the replacement of cPecanEm.py:182-188).
No GPU here, so the per-shard counts come from the oracle (the checker) -- what is under test is the partitioning,
the shard extraction and the collective."""
import os
import socket

import numpy as np
import pytest

import helpers
from cpecan_b200 import sharding, synth


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_contiguous_shards_balance_and_cover():
    rng = np.random.default_rng(0)
    for world in (1, 2, 4, 8):
        costs = rng.integers(1, 1000, size=int(rng.integers(world, 400)))
        b = sharding.contiguous_shards(costs, world)
        assert b[0] == 0 and b[-1] == costs.size and all(b[i] <= b[i + 1] for i in range(world))
        loads = [costs[b[i]:b[i + 1]].sum() for i in range(world)]
        assert max(loads) <= costs.sum() / world + costs.max()
    assert sharding.contiguous_shards([], 4) == [0, 0, 0, 0, 0]


def test_shard_packed_round_trip():
    pk = synth.evolved_pairs(9, 120, seed=4, trim=2, expansion=6)
    b = sharding.contiguous_shards(sharding.estimate_cost(pk, 6), 3)
    k = 0
    for r in range(3):
        sh = sharding.shard_packed(pk, b[r], b[r + 1])
        for i in range(b[r + 1] - b[r]):
            a, c = synth.unpack(sh, i), synth.unpack(pk, k)
            assert a[0] == c[0] and a[1] == c[1] and np.array_equal(a[2], c[2])
            k += 1
    assert k == 9


def _worker(rank, world, port, q):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    orc = helpers.port_oracle()
    pk = synth.evolved_pairs(10, 150, seed=8, trim=0, expansion=10)
    p = orc.default_params()
    p.diagonalExpansion = 10
    p.minDiagsBetweenTraceBack = 100
    spec = helpers.ModelSpec(2)
    bounds = sharding.contiguous_shards(sharding.estimate_cost(pk, 10), world)
    shard = sharding.shard_packed(pk, bounds[rank], bounds[rank + 1])
    local = np.zeros(3 * 3 + 3 * 16 + 1)
    for i in range(bounds[rank + 1] - bounds[rank]):
        sx, sy, a = synth.unpack(shard, i)
        local += orc.expectations(spec.orc(), p, sx, sy, a)
    total = sharding.allreduce_expectations(local)
    q.put((rank, bounds, total))
    dist.barrier()
    dist.destroy_process_group()


def test_expectations_all_reduce_world_size_2():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    results = [q.get(timeout=120) for _ in range(2)]
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    # single-process answer
    orc = helpers.port_oracle()
    pk = synth.evolved_pairs(10, 150, seed=8, trim=0, expansion=10)
    p = orc.default_params()
    p.diagonalExpansion = 10
    p.minDiagsBetweenTraceBack = 100
    want = np.zeros(58)
    for i in range(10):
        sx, sy, a = synth.unpack(pk, i)
        want += orc.expectations(helpers.ModelSpec(2).orc(), p, sx, sy, a)
    for rank, bounds, total in results:
        assert bounds[0] == 0 and bounds[-1] == 10 and 0 < bounds[1] < 10
        np.testing.assert_allclose(total, want, rtol=1e-12)
    n = sharding.normalise_hmm(want, 3)
    np.testing.assert_allclose(n[:9].reshape(3, 3).sum(1), 1.0)
    np.testing.assert_allclose(n[9:57].reshape(3, 16).sum(1), 1.0)
