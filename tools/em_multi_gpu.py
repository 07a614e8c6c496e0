"""Config 4 across GPUs: every rank runs the expectation pass on its shard of the pairs; the per-rank Hmm vectors
(S*S + 16*S + 1 doubles, left in HBM by the engine) are summed with ONE NCCL all-reduce; every rank normalises and
re-estimates the model for the next EM iteration (the replacement of cPecanEm.py:182-209).
Launch: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/em_multi_gpu.py [--pairs P]
Rank 0 checks the all-reduced first iteration against the whole batch run on one GPU (rtol 1e-9) and prints one JSON line."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import cpecan_b200 as cp  # noqa: E402
from cpecan_b200 import sharding, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=4000)
    ap.add_argument("--iterations", type=int, default=3)
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    p = cp.pairwiseAlignmentBandingParameters_construct()
    p.diagonalExpansion, p.constraintDiagonalTrim, p.splitMatrixBiggerThanThis = 10, 0, 3000 * 3000  # cPecanEm's realign options
    packed = synth.evolved_pairs(args.pairs, 2000, seed=4, trim=0, expansion=10)  # the same batch on every rank
    bounds = sharding.contiguous_shards(sharding.estimate_cost(packed, 10), world)
    mine = sharding.shard_packed(packed, bounds[rank], bounds[rank + 1])
    ctx = cp.Context(local, stream=torch.cuda.current_stream().cuda_stream)
    batch = cp.Batch(ctx, None, None, packed=mine)
    S, L = 5, cp.hmm_len(5)
    model = cp.stateMachine5_construct(cp.fiveState)
    first, times, likelihoods = None, [], []
    for it in range(args.iterations):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        batch.run(model, p, cp.MODE_EXPECTATIONS)
        tot = sharding.device_vector(batch.device_expectation_total_ptr(), L).clone()
        if world > 1:
            dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        summed = tot.cpu().numpy()
        times.append(time.perf_counter() - t0)
        if first is None:
            first = summed.copy()
        likelihoods.append(float(summed[-1]))
        hmm = sharding.normalise_hmm(summed, S)
        model = cp.hmm_getStateMachine(cp.fiveState, hmm[:S * S].reshape(S, S), hmm[S * S:S * S + 16 * S].reshape(S, 16))
    cells = torch.tensor([float(batch.stats().cells)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(cells)
    if rank == 0:
        whole = cp.Batch(ctx, None, None, packed=packed)
        whole.run(cp.stateMachine5_construct(cp.fiveState), p, cp.MODE_EXPECTATIONS)
        want = whole.fetch_expectations(per_pair=False)[1]
        np.testing.assert_allclose(first, want, rtol=1e-9, atol=1e-9)
        best = min(times)
        print(json.dumps({"config": "C4: %d x 2 kb expectations SM5 over %d GPU(s), one NCCL all-reduce of %d doubles per EM iteration" % (args.pairs, world, L),
                          "n_gpus": world, "cells": float(cells[0]), "s_per_em_iteration": best, "gcups": float(cells[0]) / best / 1e9,
                          "allreduce_matches_single_gpu": True, "likelihood_by_iteration": likelihoods}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
