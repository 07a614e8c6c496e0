"""Device-resident throughput of the other BASELINE.json configs (the headline bench is configs[1]).

  C1  1 pair, 1 kb, library defaults                         (latency of one call)
  C1c 1 pair, 1 kb, cPecanRealign defaults (e=4, trim 0, split 10)
  C3  N x 100 kb pairs, anchors + split points, library defaults
  C4  N x 2 kb pairs, expectations, SM3 and SM5, cPecanEm realign options (e=10, split 3000^2)
  C5  all pairs of 64 x 10 kb sequences (2016 pairs)
One JSON line per config: cells, ms per pass, GCUPS, pairs/s.  python tools/measure_configs.py [--scale 0.1]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import cpecan_b200 as cp  # noqa: E402
from cpecan_b200 import synth  # noqa: E402


def timed(batch, model, p, mode, steps=2, keep_plan=False):
    # keep_plan: the batch keeps its regions / bands / chunks between runs (EM iterations of a resident batch); otherwise every pass
    # plans anew, as the one call on a new batch does
    if keep_plan:
        os.environ.pop("CPB_NO_PLAN_CACHE", None)
    else:
        os.environ["CPB_NO_PLAN_CACHE"] = "1"
    batch.run(model, p, mode)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        batch.run(model, p, mode)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / steps


def line(name, n, st, dt, extra=None):
    out = {"config": name, "pairs": n, "cells": int(st.cells), "regions": int(st.nRegions), "blocks": int(st.nBlocks),
           "max_band_width": int(st.maxWidth), "ms_per_pass": 1e3 * dt, "gcups": st.cells / dt / 1e9, "pairs_per_s": n / dt,
           "phase_ms": {"band": st.msBand, "forward": st.msForward, "forward_checkpoint_pass": st.msCheckpoint, "backward": st.msBackward, "totals": st.msTotals, "posterior": st.msPosterior}}
    if extra:
        out.update(extra)
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=0.05, help="fraction of the BASELINE pair counts")
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    torch.cuda.set_device(0)
    ctx = cp.Context(0, stream=torch.cuda.current_stream().cuda_stream)
    sm5, sm3 = cp.stateMachine5_construct(cp.fiveState), cp.stateMachine3_construct(cp.threeState)
    want = set(args.only.split(",")) if args.only else None

    def on(k):
        return want is None or k in want

    p = cp.pairwiseAlignmentBandingParameters_construct()
    if on("C1"):
        packed = synth.evolved_pairs(1, 1000, seed=1, trim=int(p.constraintDiagonalTrim), expansion=int(p.diagonalExpansion))
        b = cp.Batch(ctx, None, None, packed=packed)
        dt = timed(b, sm5, p, cp.MODE_ALIGNED_PAIRS, steps=20)
        line("C1: 1 x 1 kb, library defaults", 1, b.stats(), dt)
        b.close()
        pc = cp.pairwiseAlignmentBandingParameters_construct()
        pc.constraintDiagonalTrim, pc.diagonalExpansion, pc.splitMatrixBiggerThanThis = 0, 4, 10
        packed = synth.evolved_pairs(1, 1000, seed=1, trim=0, expansion=4)
        b = cp.Batch(ctx, None, None, packed=packed)
        dt = timed(b, sm5, pc, cp.MODE_ALIGNED_PAIRS, steps=20)
        line("C1c: 1 x 1 kb, cPecanRealign defaults", 1, b.stats(), dt)
        b.close()
        n = max(1, int(100000 * args.scale))
        packed = synth.evolved_pairs(n, 1000, seed=2, trim=0, expansion=4)
        b = cp.Batch(ctx, None, None, packed=packed)
        dt = timed(b, sm5, pc, cp.MODE_ALIGNED_PAIRS)
        line("C2c: %d x 1 kb, cPecanRealign defaults" % n, n, b.stats(), dt)
        b.close()
    if on("C3"):
        n = max(1, int(10000 * args.scale))
        t0 = time.perf_counter()
        packed = synth.evolved_pairs(n, 100000, seed=3, trim=int(p.constraintDiagonalTrim), expansion=int(p.diagonalExpansion), batch=64)
        gen = time.perf_counter() - t0
        b = cp.Batch(ctx, None, None, packed=packed)
        dt = timed(b, sm5, p, cp.MODE_ALIGNED_PAIRS, steps=1)
        line("C3: %d x 100 kb, anchors + splits, library defaults" % n, n, b.stats(), dt, {"datagen_s": round(gen, 1), "chunks": int(b.stats().nChunks)})
        b.close()
    if on("C4"):
        n = max(1, int(50000 * args.scale))
        pe = cp.pairwiseAlignmentBandingParameters_construct()
        pe.diagonalExpansion, pe.constraintDiagonalTrim, pe.splitMatrixBiggerThanThis = 10, 0, 3000 * 3000
        packed = synth.evolved_pairs(n, 2000, seed=4, trim=0, expansion=10)
        b = cp.Batch(ctx, None, None, packed=packed)
        for name, m in (("SM5", sm5), ("SM3", sm3)):
            dt = timed(b, m, pe, cp.MODE_EXPECTATIONS)
            line("C4: %d x 2 kb expectations %s (e=10)" % (n, name), n, b.stats(), dt)
            dt = timed(b, m, pe, cp.MODE_EXPECTATIONS, keep_plan=True)
            line("C4: %d x 2 kb expectations %s (e=10), plan kept between EM iterations" % (n, name), n, b.stats(), dt)
        b.close()
    if on("C5"):
        k = max(4, int(round(64 * args.scale ** 0.5)))
        rng = np.random.default_rng(5)
        anc = synth.random_sequence(rng, 10000, acgt_only=True)
        members = [synth.evolve_with_alignment(rng, anc) for _ in range(k)]
        sx, sy, an = [], [], []
        for i in range(k):
            for j in range(i + 1, k):
                sx.append(members[i][0])
                sy.append(members[j][0])
                an.append(synth.anchors_between(members[i], members[j], trim=int(p.constraintDiagonalTrim), expansion=int(p.diagonalExpansion)))
        b = cp.Batch(ctx, sx, sy, an)
        dt = timed(b, sm5, p, cp.MODE_ALIGNED_PAIRS)
        line("C5: all pairs of %d x 10 kb (%d pairs)" % (k, len(sx)), len(sx), b.stats(), dt)
        b.close()
    ctx.close()


if __name__ == "__main__":
    main()
