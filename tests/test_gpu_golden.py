"""GPU: the CUDA path against the committed golden fixtures (outputs of the reference's own code, tools/make_golden.py).
These need no oracle at run time."""
import json
import os

import numpy as np
import pytest

import cpecan_b200 as cp
import helpers

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_cases.json")


def params_of(c):
    p = cp.pairwiseAlignmentBandingParameters_construct()
    for k, v in c["params"].items():
        setattr(p, k, v)
    return p


def test_golden_cases(ctx):
    with open(GOLDEN) as f:
        golden = json.load(f)
    npairs = 0
    for c in golden["cases"]:
        spec = helpers.ModelSpec(c["type"], c["transitions"], c["emissions"])
        p = params_of(c)
        a = np.asarray(c["anchors"], dtype=np.int64).reshape(-1, 3)
        m = spec.cpb()
        res = cp.getAlignedPairsWithIndelsUsingAnchors(m, c["sX"], c["sY"], a, p, c["raggedLeft"], c["raggedRight"], ctx=ctx)
        for got, key in zip(res, ("alignedPairs", "gapXPairs", "gapYPairs")):
            g, w = helpers.sort_triples(got), helpers.sort_triples(c[key])
            assert g.shape == w.shape and np.array_equal(g[:, 1:], w[:, 1:]), "%s %s" % (c["name"], key)
            assert np.array_equal(g[:, 0], w[:, 0])  # weights bit-exact (host libm recomputes the few next to an integer)
            npairs += g.shape[0]
        only = helpers.sort_triples(cp.getAlignedPairsUsingAnchors(m, c["sX"], c["sY"], a, p, c["raggedLeft"], c["raggedRight"], ctx=ctx))
        assert np.array_equal(only[:, 1:], helpers.sort_triples(c["alignedPairs"])[:, 1:])
        S = m.stateNumber
        hmm = cp.getExpectationsUsingAnchors(m, np.zeros(cp.hmm_len(S)), c["sX"], c["sY"], a, p, c["raggedLeft"], c["raggedRight"], ctx=ctx)
        want = np.array([float.fromhex(v) for v in c["expectations"]])
        np.testing.assert_allclose(hmm, want, rtol=1e-9, atol=1e-12, err_msg=c["name"])
        if "forwardLogProb" in c:
            got = cp.computeForwardProbability(c["sX"], c["sY"], a, p, m, c["raggedLeft"], c["raggedRight"], ctx=ctx)
            assert float.hex(got) == c["forwardLogProb"], c["name"]
        assert ctx.band(a, len(c["sX"]), len(c["sY"]), int(p.diagonalExpansion), bool(p.dynamicAnchorExpansion)).tolist() == c["band"]
    assert npairs > 3000
