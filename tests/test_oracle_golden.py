"""CPU tests (no GPU): pin the plain-C oracle (oracle/pairhmm_oracle.c) against

  * the reference's own known-answer tests (tests/pairwiseAlignerTest.c in the cPecan tree), written out literally, and
  * tests/golden/reference_cases.json, produced by tools/make_golden.py from the reference's unmodified sources
    (oracle/_ref).  Everything is compared bit for bit (floats travel as C99 hex literals).
"""
import json
import math
import os

import numpy as np
import pytest

import helpers

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_cases.json")


@pytest.fixture(scope="module")
def golden():
    with open(GOLDEN) as f:
        return json.load(f)


def spec_of(c):
    return helpers.ModelSpec(c["type"], c["transitions"], c["emissions"])


def params_of(o, c):
    p = o.default_params()
    for k, v in c["params"].items():
        setattr(p, k, v)
    return p


def test_band_kat(port):
    """test_bands, tests/pairwiseAlignerTest.c:69-93"""
    want = [(0, 0, 0), (1, -1, 1), (2, -2, 2), (3, -1, 3), (4, -2, 4), (5, -1, 3), (6, -2, 4), (7, -3, 3), (8, -2, 2), (9, -1, 3), (10, 0, 2),
            (11, 1, 1)]
    got = port.band([(1, 0, 2), (2, 1, 2), (3, 3, 2)], 6, 5, 2)
    assert [tuple(r) for r in got.tolist()] == want


def test_split_points_kat(port):
    """test_getSplitPoints, tests/pairwiseAlignerTest.c:578-647"""
    m = 2000 * 2000
    assert port.split_points([], 3000, 1000, m, 0, 0) == [(0, 0, 3000, 1000)]
    assert port.split_points([], 20000, 25000, m, 1, 1) == []
    assert port.split_points([], 20000, 25000, m, 1, 0) == [(18000, 23000, 20000, 25000)]
    assert port.split_points([], 20000, 25000, m, 0, 1) == [(0, 0, 2000, 2000)]
    assert port.split_points([], 20000, 25000, m, 0, 0) == [(0, 0, 2000, 2000), (18000, 23000, 20000, 25000)]
    anchors = [(2000, 2000, 0), (4002, 4001, 0), (5000, 5000, 0), (8000, 6000, 0), (9000, 9000, 0), (10000, 14000, 0), (15000, 15000, 0),
               (16000, 16000, 0)]
    assert port.split_points(anchors, 20000, 25000, m, 0, 0) == [(0, 0, 3001, 3001), (3002, 3001, 9500, 11001), (9501, 12000, 12001, 14500),
                                                                 (13000, 14501, 18000, 18001), (18001, 23000, 20000, 25000)]


def test_diagonal_dp_kat(port):
    """test_diagonalDPCalculations, tests/pairwiseAlignerTest.c:242-324: AGCG vs AGTTCG, threshold 0.2 -> exactly 4 pairs"""
    p = port.default_params()
    p.threshold = 0.2
    m = helpers.ModelSpec(0).orc()
    got = helpers.sort_triples(port.aligned_pairs(m, p, "AGCG", "AGTTCG", []))
    assert got.tolist() == [[9944673, 0, 0], [9259684, 1, 1], [8665179, 2, 4], [9893294, 3, 5]]
    assert port.forward_prob(m, p, "AGCG", "AGTTCG", []) == -17.51932116123855


def test_logadd_property(port):
    """test_logAdd, tests/pairwiseAlignerTest.c:134-144"""
    rng = np.random.default_rng(1)
    for _ in range(20000):
        i, j = rng.random(), rng.random()
        assert abs(math.exp(port.logadd(math.log(i), math.log(j))) - (i + j)) < 0.001


def test_logadd_golden(port, golden):
    for xh, yh, rh in golden["extra"]["logadd"]:
        x, y, r = float.fromhex(xh), float.fromhex(yh), float.fromhex(rh)
        got = port.logadd(x, y)
        assert got == r or (math.isnan(got) and math.isnan(r)), (x, y, got, r)


def test_default_params(port):
    """pairwiseAlignmentBandingParameters_construct, impl/pairwiseAligner.c:1334-1348"""
    p = port.default_params()
    assert (p.threshold, p.minDiagsBetweenTraceBack, p.traceBackDiagonals, p.diagonalExpansion, p.constraintDiagonalTrim,
            p.splitMatrixBiggerThanThis, p.dynamicAnchorExpansion) == (0.01, 1000, 40, 20, 14, 3000 * 3000, 0)


def test_golden_cases_bit_exact(port, golden):
    assert len(golden["cases"]) >= 16
    for c in golden["cases"]:
        spec, p = spec_of(c), params_of(port, c)
        a = np.asarray(c["anchors"], dtype=np.int64).reshape(-1, 3)
        om = spec.orc()
        got = port.aligned_pairs(om, p, c["sX"], c["sY"], a, c["raggedLeft"], c["raggedRight"])
        assert got.tolist() == c["alignedPairs"], c["name"]
        exp = port.expectations(om, p, c["sX"], c["sY"], a, c["raggedLeft"], c["raggedRight"])
        assert [float.hex(v) for v in exp] == c["expectations"], c["name"]
        assert port.band(a, len(c["sX"]), len(c["sY"]), p.diagonalExpansion, bool(p.dynamicAnchorExpansion)).tolist() == c["band"]
        assert [list(t) for t in port.split_points(a, len(c["sX"]), len(c["sY"]), p.splitMatrixBiggerThanThis, c["raggedLeft"], c["raggedRight"])] == \
            [list(t) for t in c["splitPoints"]]
        assert [float.hex(v) for v in port.model_dump(om)] == c["modelDump"], c["name"]
        if "forwardLogProb" in c:
            assert float.hex(port.forward_prob(om, p, c["sX"], c["sY"], a, c["raggedLeft"], c["raggedRight"])) == c["forwardLogProb"]
        ind = port.aligned_pairs_with_indels(om, p, c["sX"], c["sY"], a, c["raggedLeft"], c["raggedRight"])
        assert ind[0].tolist() == c["alignedPairs"] and ind[1].tolist() == c["gapXPairs"] and ind[2].tolist() == c["gapYPairs"]


def test_forward_equals_backward_property(port):
    """test_cell / test_diagonalDPCalculations: with one traceback block the posteriors of every row sum to ~1"""
    rng = np.random.default_rng(3)
    from cpecan_b200 import synth

    p = port.default_params()
    p.threshold = 0.0
    sX = synth.random_sequence(rng, 30, acgt_only=True)
    sY = synth.evolve_like_reference(rng, sX)
    t = port.aligned_pairs(helpers.ModelSpec(0).orc(), p, sX, sY, [])
    rows = np.bincount(t[:, 1], weights=t[:, 0] / 1e7, minlength=len(sX))
    assert rows.max() < 1.01


def test_expectation_block_boundary_quirk(port):
    """at every non-final traceback block boundary the ->match transitions of one diagonal are dropped
    (impl/pairwiseAligner.c:855 frees the forward diagonal the middle group needs), so with several blocks the match-transition
    mass is smaller than with one block."""
    rng = np.random.default_rng(4)
    from cpecan_b200 import synth

    pk = synth.evolved_pairs(1, 600, seed=9, trim=0, expansion=10)
    sx, sy, a = synth.unpack(pk, 0)
    m = helpers.ModelSpec(0).orc()
    one = port.default_params()
    one.diagonalExpansion = 10
    many = port.default_params()
    many.diagonalExpansion = 10
    many.minDiagsBetweenTraceBack = 100
    e1 = port.expectations(m, one, sx, sy, a)
    e2 = port.expectations(m, many, sx, sy, a)
    to_match_1 = e1[:25].reshape(5, 5)[:, 0].sum()
    to_match_2 = e2[:25].reshape(5, 5)[:, 0].sum()
    assert to_match_2 < to_match_1 - 1.0


@pytest.mark.skipif(helpers.ref_oracle() is None, reason="oracle/_ref not built (needs /root/reference)")
def test_port_matches_reference_build_live(port):
    """differential fuzz against the reference's own code, run wherever oracle/_ref exists"""
    from cpecan_b200 import synth

    ref = helpers.ref_oracle()
    rng = np.random.default_rng(77)
    for it in range(60):
        sX = synth.random_sequence(rng, int(rng.integers(0, 200)))
        sY = synth.evolve_like_reference(rng, sX)
        a = synth.random_anchor_pairs(rng, len(sX), len(sY))
        p = ref.default_params()
        p.traceBackDiagonals = int(rng.integers(1, 10))
        p.minDiagsBetweenTraceBack = p.traceBackDiagonals + int(rng.integers(2, 10))
        p.diagonalExpansion = 2 * int(rng.integers(0, 10))
        p.dynamicAnchorExpansion = int(rng.random() > 0.5)
        p.threshold = [0.01, 0.2, 0.0][it % 3]
        if it % 4 == 0:
            p.splitMatrixBiggerThanThis = int(rng.integers(5, 400))
        t = int(rng.integers(0, 4))
        spec = helpers.ModelSpec(t) if it % 2 else helpers.ModelSpec.random(rng, t)
        rl, rr = bool(rng.random() > 0.5), bool(rng.random() > 0.5)
        om = spec.orc()
        assert np.array_equal(ref.aligned_pairs(om, p, sX, sY, a, rl, rr), port.aligned_pairs(om, p, sX, sY, a, rl, rr))
        assert np.array_equal(ref.expectations(om, p, sX, sY, a, rl, rr), port.expectations(om, p, sX, sY, a, rl, rr))
        assert np.array_equal(ref.model_dump(om), port.model_dump(om))
        if not p.dynamicAnchorExpansion:
            f1, f2 = ref.forward_prob(om, p, sX, sY, a, rl, rr), port.forward_prob(om, p, sX, sY, a, rl, rr)
            assert f1 == f2 or (math.isnan(f1) and math.isnan(f2))
