"""CPU tests (no GPU) of the product's host side: the C-ABI library loads and exports every symbol include/cpecan_b200.h
declares, parameter / model construction matches the reference (through the golden model dumps and, where available, the
oracle), split points match, and compute entry points fail loudly without a device."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest

import cpecan_b200 as cp
import helpers

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "reference_cases.json")


def test_library_exports_every_declared_symbol():
    text = open(os.path.join(ROOT, "include", "cpecan_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = set(re.findall(r"\b(cpb_[a-z0-9_]+)\s*\(", text))
    assert len(names) >= 20
    lib = C.CDLL(cp.LIB_PATH)
    missing = [n for n in sorted(names) if not hasattr(lib, n)]
    assert not missing, "declared in include/cpecan_b200.h but not exported: %s" % missing


def test_struct_layouts_match_the_header():
    """sizeof() of the ctypes mirrors must equal what the C side was compiled with (checked through a round trip)"""
    p = cp.pairwiseAlignmentBandingParameters_construct()
    assert C.sizeof(cp.CpbParams) == 80
    assert (p.threshold, p.minDiagsBetweenTraceBack, p.traceBackDiagonals, p.diagonalExpansion, p.constraintDiagonalTrim) == (0.01, 1000, 40, 20, 14)
    assert (p.anchorMatrixBiggerThanThis, p.repeatMaskMatrixBiggerThanThis, p.splitMatrixBiggerThanThis) == (250000, 250000, 9000000)
    assert (p.alignAmbiguityCharacters, p.dynamicAnchorExpansion) == (0, 0) and abs(p.gapGamma - 0.5) < 1e-9
    assert C.sizeof(cp.CpbModel) == 8 + 8 * (20 + 13 + 35)


def model_table(m):
    """(group, from, to, eP, tP) records in the reference's issue order, for every (cX, cY): what orc_model_dump emits"""
    S = m.stateNumber
    if S == 5:
        lower = [(0, 1), (1, 1), (0, 3), (3, 3)]
        middle = [(0, 0), (1, 0), (2, 0), (3, 0), (4, 0)]
        upper = [(0, 2), (2, 2), (0, 4), (4, 4)]
    else:
        lower = [(0, 1), (1, 1), (2, 1)]
        middle = [(0, 0), (1, 0), (2, 0)]
        upper = [(0, 2), (2, 2), (1, 2)]
    out = []
    for v in (m.start, m.raggedStart, m.end, m.raggedEnd):
        out.extend(list(v)[:S])
    for cx in range(5):
        for cy in range(5):
            for k, (f, t) in enumerate(lower):
                out.extend([0, f, t, m.eGapX[cx], m.tLower[k]])
            for k, (f, t) in enumerate(middle):
                out.extend([1, f, t, m.eMatch[cx * 5 + cy], m.tMiddle[k]])
            for k, (f, t) in enumerate(upper):
                out.extend([2, f, t, m.eGapY[cy], m.tUpper[k]])
    return np.array(out, dtype=np.float64)


@pytest.mark.parametrize("type_", [0, 1, 2, 3])
def test_default_models_match_the_oracle(port, type_):
    """stateMachine5_construct / stateMachine3_construct (impl/stateMachine.c:482, :716)"""
    spec = helpers.ModelSpec(type_)
    assert np.array_equal(model_table(spec.cpb()), port.model_dump(spec.orc()))


def test_models_from_hmm_match_golden_dumps():
    """hmm_getStateMachine (impl/stateMachine.c:797) incl. the symmetric averaging and the short/long swap, bit for bit
    against dumps taken from the reference build"""
    with open(GOLDEN) as f:
        golden = json.load(f)
    seen = set()
    for c in golden["cases"]:
        spec = helpers.ModelSpec(c["type"], c["transitions"], c["emissions"])
        got = [float.hex(v) for v in model_table(spec.cpb())]
        assert got == c["modelDump"], c["name"]
        seen.add((c["type"], c["transitions"] is not None))
    assert len(seen) == 8


def test_models_from_random_hmms_match_the_oracle(port):
    rng = np.random.default_rng(12)
    for it in range(200):
        t = int(rng.integers(0, 4))
        spec = helpers.ModelSpec.random(rng, t)
        if it % 3 == 0 and t < 2:
            # force the short/long swap branch (stateMachine.c:544-550): make the "short" extend larger than the "long" one
            spec.transitions[1, 1], spec.transitions[3, 3] = 0.9, 0.1
            spec.transitions[2, 2], spec.transitions[4, 4] = 0.8, 0.05
        assert np.array_equal(model_table(spec.cpb()), port.model_dump(spec.orc()))


def test_wrong_model_type_is_rejected():
    with pytest.raises(cp.CpbError):
        cp.stateMachine5_construct(cp.threeState)
    with pytest.raises(cp.CpbError):
        cp.stateMachine3_construct(cp.fiveState)
    m = cp.CpbModel()
    assert cp.lib.cpb_model_default(7, C.byref(m)) != 0


def test_split_points_kat_and_random(port):
    """getSplitPoints (impl/pairwiseAligner.c:1230): the reference's golden tuples (tests/pairwiseAlignerTest.c:578-647) and random anchors"""
    m = 2000 * 2000
    assert cp.getSplitPoints([], 3000, 1000, m, 0, 0) == [(0, 0, 3000, 1000)]
    assert cp.getSplitPoints([], 20000, 25000, m, 1, 1) == []
    assert cp.getSplitPoints([], 20000, 25000, m, 1, 0) == [(18000, 23000, 20000, 25000)]
    assert cp.getSplitPoints([], 20000, 25000, m, 0, 1) == [(0, 0, 2000, 2000)]
    assert cp.getSplitPoints([], 20000, 25000, m, 0, 0) == [(0, 0, 2000, 2000), (18000, 23000, 20000, 25000)]
    anchors = [(2000, 2000, 0), (4002, 4001, 0), (5000, 5000, 0), (8000, 6000, 0), (9000, 9000, 0), (10000, 14000, 0), (15000, 15000, 0),
               (16000, 16000, 0)]
    assert cp.getSplitPoints(anchors, 20000, 25000, m, 0, 0) == [(0, 0, 3001, 3001), (3002, 3001, 9500, 11001), (9501, 12000, 12001, 14500),
                                                                  (13000, 14501, 18000, 18001), (18001, 23000, 20000, 25000)]
    from cpecan_b200 import synth

    rng = np.random.default_rng(2)
    for _ in range(300):
        lX, lY = int(rng.integers(0, 400)), int(rng.integers(0, 400))
        a = synth.random_anchor_pairs(rng, lX, lY)
        split = int(rng.integers(1, 500))
        rl, rr = int(rng.random() > 0.5), int(rng.random() > 0.5)
        assert cp.getSplitPoints(a, lX, lY, split, rl, rr) == port.split_points(a, lX, lY, split, rl, rr)


def test_compute_entry_points_fail_loudly_without_a_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(cp.CpbError, match="no CUDA device|CUDA"):
        cp.Context(0)
    with pytest.raises(cp.CpbError):
        cp.getAlignedPairsUsingAnchors(cp.stateMachine5_construct(), "ACGT", "ACGT", [], cp.pairwiseAlignmentBandingParameters_construct())


def test_product_never_touches_the_oracle():
    """the oracle is test infrastructure: nothing under cpecan_b200/ or include/ may mention it"""
    for base in ("cpecan_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for fn in files:
                if fn.endswith((".py", ".c", ".cu", ".cuh", ".h", "Makefile")):
                    text = open(os.path.join(dirpath, fn), errors="ignore").read()
                    assert "oracle" not in text.lower() or fn == "__init__.py", "%s mentions the oracle" % os.path.join(dirpath, fn)
