#!/usr/bin/env python
"""Generate tests/golden/*.json from the reference's own code (oracle/_ref, built from /root/reference by
`make -f oracle/Makefile ref`).  Run in the build container only; the fixtures are committed so that the oracle
restatement and the CUDA path can be pinned where /root/reference does not exist.

  python tools/make_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import helpers  # noqa: E402
from cpecan_b200 import synth  # noqa: E402


def params(o, **kw):
    p = o.default_params()
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def pdict(p):
    return {f: getattr(p, f) for f, _ in p._fields_}


def main():
    ref = helpers.ref_oracle()
    assert ref is not None and ref.identity == "reference", "build oracle/_ref first (needs /root/reference)"
    rng = np.random.default_rng(20261018)
    cases = []

    def add(name, spec, p, sX, sY, anchors, rl, rr):
        a = np.asarray(anchors, dtype=np.int64).reshape(-1, 3)
        om = spec.orc()
        c = {
            "name": name, "type": spec.type,
            "transitions": None if spec.transitions is None else spec.transitions.ravel().tolist(),
            "emissions": None if spec.emissions is None else spec.emissions.ravel().tolist(),
            "params": pdict(p), "sX": sX, "sY": sY, "anchors": a.tolist(), "raggedLeft": bool(rl), "raggedRight": bool(rr),
            "alignedPairs": ref.aligned_pairs(om, p, sX, sY, a, rl, rr).tolist(),
            "expectations": [float.hex(v) for v in ref.expectations(om, p, sX, sY, a, rl, rr)],
            "band": ref.band(a, len(sX), len(sY), p.diagonalExpansion, bool(p.dynamicAnchorExpansion)).tolist(),
            "splitPoints": ref.split_points(a, len(sX), len(sY), p.splitMatrixBiggerThanThis, rl, rr),
            "modelDump": [float.hex(v) for v in ref.model_dump(om)],
        }
        if not p.dynamicAnchorExpansion:
            c["forwardLogProb"] = float.hex(ref.forward_prob(om, p, sX, sY, a, rl, rr))
        ind = ref.aligned_pairs_with_indels(om, p, sX, sY, a, rl, rr)
        c["gapXPairs"] = ind[1].tolist()
        c["gapYPairs"] = ind[2].tolist()
        cases.append(c)

    # the reference's own known-answer test (tests/pairwiseAlignerTest.c:242-324)
    add("kat_agcg", helpers.ModelSpec(0), params(ref, threshold=0.2), "AGCG", "AGTTCG", [], False, False)
    # small random cases in the shape of test_getAlignedPairsWithBanding (:403-438), all model types
    for t in range(4):
        for rep in range(3):
            sX = synth.random_sequence(rng, int(rng.integers(5, 90)))
            sY = synth.evolve_like_reference(rng, sX)
            a = synth.random_anchor_pairs(rng, len(sX), len(sY))
            tb = int(rng.integers(1, 10))
            p = params(ref, traceBackDiagonals=tb, minDiagsBetweenTraceBack=tb + int(rng.integers(2, 10)),
                       diagonalExpansion=2 * int(rng.integers(0, 10)), dynamicAnchorExpansion=int(rep == 2),
                       threshold=[0.01, 0.2, 0.0][rep], splitMatrixBiggerThanThis=[9000000, 9000000, 60][rep])
            spec = helpers.ModelSpec(t) if rep == 0 else helpers.ModelSpec.random(rng, t)
            add("small_t%d_r%d" % (t, rep), spec, p, sX, sY, a, rep == 1, rep == 2)
    # evolved pairs with lastz-style anchors: library defaults and cPecanRealign defaults (cPecanRealign.c:355-357)
    pk = synth.evolved_pairs(2, 400, seed=5, trim=14, expansion=20)
    for i in range(2):
        sx, sy, a = synth.unpack(pk, i)
        add("evolved400_lib_%d" % i, helpers.ModelSpec(0), params(ref, minDiagsBetweenTraceBack=300), sx.decode(), sy.decode(), a, False, False)
    pk = synth.evolved_pairs(1, 400, seed=6, trim=0, expansion=4)
    sx, sy, a = synth.unpack(pk, 0)
    add("evolved400_cli", helpers.ModelSpec(2), params(ref, constraintDiagonalTrim=0, diagonalExpansion=4, splitMatrixBiggerThanThis=10),
        sx.decode(), sy.decode(), a, False, False)

    extra = {
        # tests/pairwiseAlignerTest.c:69-93
        "band_kat": {"anchors": [[1, 0, 2], [2, 1, 2], [3, 3, 2]], "lX": 6, "lY": 5, "expansion": 2,
                     "band": ref.band(np.array([[1, 0, 2], [2, 1, 2], [3, 3, 2]]), 6, 5, 2).tolist()},
        # tests/pairwiseAlignerTest.c:578-647
        "split_kat": [
            {"anchors": [], "lX": 3000, "lY": 1000, "split": 4000000, "rl": 0, "rr": 0, "out": ref.split_points([], 3000, 1000, 4000000, 0, 0)},
            {"anchors": [], "lX": 20000, "lY": 25000, "split": 4000000, "rl": 1, "rr": 1, "out": ref.split_points([], 20000, 25000, 4000000, 1, 1)},
            {"anchors": [], "lX": 20000, "lY": 25000, "split": 4000000, "rl": 1, "rr": 0, "out": ref.split_points([], 20000, 25000, 4000000, 1, 0)},
            {"anchors": [], "lX": 20000, "lY": 25000, "split": 4000000, "rl": 0, "rr": 1, "out": ref.split_points([], 20000, 25000, 4000000, 0, 1)},
            {"anchors": [], "lX": 20000, "lY": 25000, "split": 4000000, "rl": 0, "rr": 0, "out": ref.split_points([], 20000, 25000, 4000000, 0, 0)},
        ],
        "logadd": [],
    }
    eight = [[2000, 2000, 0], [4002, 4001, 0], [5000, 5000, 0], [8000, 6000, 0], [9000, 9000, 0], [10000, 14000, 0], [15000, 15000, 0],
             [16000, 16000, 0]]
    extra["split_kat"].append({"anchors": eight, "lX": 20000, "lY": 25000, "split": 4000000, "rl": 0, "rr": 0,
                               "out": ref.split_points(np.array(eight), 20000, 25000, 4000000, 0, 0)})
    for _ in range(400):
        x = float(np.log(rng.random()) * rng.integers(1, 40))
        y = x + float(rng.normal() * [0.3, 2.0, 6.0, 20.0][int(rng.integers(0, 4))])
        extra["logadd"].append([float.hex(x), float.hex(y), float.hex(ref.logadd(x, y))])
    for x, y in [(0.0, 0.0), (-1.0, -2.0), (-2.0, -1.0), (-3.5, -1.0), (-1.0, -5.5), (-1.0, -8.5), (-9.0, -1.5), (float("-inf"), -3.0),
                 (-3.0, float("-inf")), (float("-inf"), float("-inf")), (-1.0, -3.5), (-1.0, -2.0 - 1e-15)]:
        extra["logadd"].append([float.hex(x), float.hex(y), float.hex(ref.logadd(x, y))])

    out = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "reference_cases.json"), "w") as f:
        json.dump({"generator": "tools/make_golden.py", "source": "oracle/_ref (cPecan impl/pairwiseAligner.c + impl/stateMachine.c, unmodified)",
                   "cases": cases, "extra": extra}, f)
    print("wrote %d cases" % len(cases), os.path.getsize(os.path.join(out, "reference_cases.json")), "bytes")


if __name__ == "__main__":
    main()
