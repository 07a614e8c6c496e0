"""Writes tests/golden/realign_cases.json: inputs and outputs of the reference's own list post-processing functions
(reweightAlignedPairs2, scoreBy*, getMaximalExpectedAccuracyPairwiseAlignment, leftShiftAlignment;
impl/pairwiseAligner.c:1519-1792), run through oracle/_ref (the reference compiled unmodified).  Needs /root/reference, so it
runs in the build container only; the fixture travels.  Usage: python tools/make_realign_golden.py"""
import json
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import helpers  # noqa: E402
import test_realign_host as t  # noqa: E402


def main():
    ref = helpers.ref_oracle()
    assert ref is not None, "oracle/_ref is not built (make -f oracle/Makefile ref)"
    ref = t.Ref(ref)
    rng = np.random.default_rng(2024)
    cases = []
    for i in range(16):
        sX, sY, tri, gapX, gapY = case = t.random_case(rng)
        gamma = [0.0, 0.2, 0.5, 0.9][i % 4]
        cases.append({"sX": sX, "sY": sY, "pairs": tri.tolist(), "gapX": gapX.tolist(), "gapY": gapY.tolist(), "gamma": gamma,
                      "want": t.all_results(ref, case, gamma)})
    out = os.path.join(ROOT, "tests", "golden", "realign_cases.json")
    json.dump({"source": "oracle/_ref (reference impl/pairwiseAligner.c, unmodified) via tools/make_realign_golden.py", "cases": cases}, open(out, "w"))
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
