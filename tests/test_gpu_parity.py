"""GPU parity tests: the CUDA path through the C-ABI against the oracle on the same seeded inputs.

Bar (BASELINE.json north_star): aligned-pair triples (pInt, x, y) IDENTICAL -- the keep decision is taken in log space against
the host libm's own threshold crossing and the few weights floor(p * 1e7) that sit next to an integer are recomputed with the host's
exp (engine.cu), so nothing depends on the last place of CUDA's exp; forward log-probabilities bit-identical (pure logAdd
arithmetic); expectations within 1e-9 relative.
"""
import numpy as np
import pytest

import cpecan_b200 as cp
from cpecan_b200 import synth
import helpers

pytestmark = pytest.mark.gpu

EXPECT_RTOL = 1e-9


@pytest.fixture(autouse=True, params=["auto", "two_pass", "strips_only"])
def forward_schedule(request, monkeypatch):
    """Every test of this module runs three times: with the engine's own choices (narrow-band group kernels where every diagonal
    has at most 16 cells, else the strip kernels; one-sweep forward unless the run is chunked and made of long regions); with the
    checkpointed two-pass forward forced, which puts its block restarts through every edge case here (blocks a few diagonals
    long, ragged ends, splits, empty problems); and with the narrow-band kernels switched off, so that the strip kernels also see
    the narrow cases."""
    monkeypatch.delenv("CPB_TWO_PASS", raising=False)
    monkeypatch.delenv("CPB_NARROW", raising=False)
    if request.param == "two_pass":
        monkeypatch.setenv("CPB_TWO_PASS", "1")
    if request.param == "strips_only":
        monkeypatch.setenv("CPB_NARROW", "0")
    return request.param


def compare_triples(got, want, what=""):
    g = helpers.sort_triples(got)
    w = helpers.sort_triples(want)
    assert g.shape == w.shape, "%s: %d pairs from the GPU, %d from the oracle" % (what, g.shape[0], w.shape[0])
    if g.shape[0] == 0:
        return 0
    assert np.array_equal(g[:, 1:], w[:, 1:]), "%s: aligned-pair coordinate sets differ" % what
    diff = np.abs(g[:, 0] - w[:, 0])
    assert diff.max() == 0, "%s: pInt differs by %d" % (what, diff.max())  # tolerance 0: integer output is bit-exact
    return 0


def run_batch(ctx, spec, p, cases, mode):
    """cases: list of (sX, sY, anchors[k,3], rl, rr)"""
    b = cp.Batch(ctx, [c[0] for c in cases], [c[1] for c in cases], [c[2] for c in cases], [c[3] for c in cases], [c[4] for c in cases])
    b.run(spec.cpb(), p, mode)
    return b


def check_aligned_pairs(ctx, oracle, spec, p, cases, what):
    b = run_batch(ctx, spec, p, cases, cp.MODE_ALIGNED_PAIRS)
    off, tri = b.fetch_pairs(0)
    off2, tri2 = b.fetch_pairs(0, reference_order=True)
    assert np.array_equal(off, off2)
    op = helpers.orc_params_from(p)
    om = spec.orc()
    nd = 0
    total = 0
    for i, c in enumerate(cases):
        want = oracle.aligned_pairs(om, op, c[0], c[1], c[2], c[3], c[4])
        got = tri[off[i]:off[i + 1]]
        nd += compare_triples(got, want, "%s case %d (lX %d lY %d)" % (what, i, len(c[0]), len(c[1])))
        # the reference-order fetch reproduces the reference's list element for element (impl/pairwiseAligner.c:1411-1418)
        assert np.array_equal(tri2[off[i]:off[i + 1], 1:], want[:, 1:]), "%s case %d: list order differs from the reference's" % (what, i)
        total += want.shape[0]
    b.close()
    assert nd == 0
    return total


def small_cases(rng, n, max_len=100, ragged=False):
    cases = []
    for _ in range(n):
        sX = synth.random_sequence(rng, int(rng.integers(0, max_len)))
        sY = synth.evolve_like_reference(rng, sX)
        a = synth.random_anchor_pairs(rng, len(sX), len(sY))
        rl = bool(rng.random() > 0.5) if ragged else False
        rr = bool(rng.random() > 0.5) if ragged else False
        cases.append((sX, sY, a, rl, rr))
    return cases


def test_kat_agcg(ctx):
    """tests/pairwiseAlignerTest.c:242-324 (values from the reference build)"""
    p = cp.pairwiseAlignmentBandingParameters_construct()
    p.threshold = 0.2
    sM = cp.stateMachine5_construct(cp.fiveState)
    pairs = cp.getAlignedPairsUsingAnchors(sM, "AGCG", "AGTTCG", [], p, False, False, ctx=ctx)
    assert sorted(pairs, key=lambda t: t[1]) == [(9944673, 0, 0), (9259684, 1, 1), (8665179, 2, 4), (9893294, 3, 5)]
    assert cp.computeForwardProbability("AGCG", "AGTTCG", [], p, sM, False, False, ctx=ctx) == -17.51932116123855


def test_band_golden_and_random(ctx, oracle):
    """tests/pairwiseAlignerTest.c:69-93 on the device band builder, then random anchors against the oracle"""
    band = ctx.band([(1, 0, 2), (2, 1, 2), (3, 3, 2)], 6, 5, 2)
    want = [(0, 0, 0), (1, -1, 1), (2, -2, 2), (3, -1, 3), (4, -2, 4), (5, -1, 3), (6, -2, 4), (7, -3, 3), (8, -2, 2), (9, -1, 3), (10, 0, 2),
            (11, 1, 1)]
    assert [tuple(r) for r in band.tolist()] == want
    rng = np.random.default_rng(7)
    for it in range(40):
        lX, lY = int(rng.integers(0, 200)), int(rng.integers(0, 200))
        a = synth.random_anchor_pairs(rng, lX, lY)
        dyn = bool(it % 2)
        e = 2 * int(rng.integers(0, 12))
        assert np.array_equal(ctx.band(a, lX, lY, e, dyn), oracle.band(a, lX, lY, e, dyn))


@pytest.mark.parametrize("type_", [cp.fiveState, cp.fiveStateAsymmetric, cp.threeState, cp.threeStateAsymmetric])
def test_random_small_pairs(ctx, oracle, type_):
    """the shape of test_getAlignedPairsWithBanding (tests/pairwiseAlignerTest.c:403-438), but checked for values"""
    rng = np.random.default_rng(100 + type_)
    for rep in range(6):
        p = cp.pairwiseAlignmentBandingParameters_construct()
        p.traceBackDiagonals = int(rng.integers(1, 10))
        p.minDiagsBetweenTraceBack = p.traceBackDiagonals + int(rng.integers(2, 10))
        p.diagonalExpansion = int(rng.integers(0, 10)) * 2
        p.dynamicAnchorExpansion = int(rng.random() > 0.5)
        p.threshold = [0.01, 0.2, 0.0, 0.5, 0.01, 0.05][rep]
        if rep >= 4:
            p.splitMatrixBiggerThanThis = int(rng.integers(5, 400))
        spec = helpers.ModelSpec(type_) if rep % 2 == 0 else helpers.ModelSpec.random(rng, type_)
        cases = small_cases(rng, 25, 120, ragged=rep >= 2)
        check_aligned_pairs(ctx, oracle, spec, p, cases, "type %d rep %d" % (type_, rep))


def test_evolved_1kb_lib_defaults(ctx, oracle):
    """BASELINE config 1/2 shape: 1 kb evolved pairs, StateMachine5, library-default band and threshold"""
    p = cp.pairwiseAlignmentBandingParameters_construct()
    packed = synth.evolved_pairs(12, 1000, seed=11, trim=int(p.constraintDiagonalTrim), expansion=int(p.diagonalExpansion))
    cases = [synth.unpack(packed, i) + (False, False) for i in range(12)]
    n = check_aligned_pairs(ctx, oracle, helpers.ModelSpec(cp.fiveState), p, cases, "1kb lib")
    assert n > 12 * 900


def test_evolved_1kb_cli_defaults(ctx, oracle):
    """cPecanRealign's own defaults (cPecanRealign.c:355-357): trim 0, expansion 4, split 10 -> many tiny regions"""
    p = cp.pairwiseAlignmentBandingParameters_construct()
    p.constraintDiagonalTrim = 0
    p.diagonalExpansion = 4
    p.splitMatrixBiggerThanThis = 10
    packed = synth.evolved_pairs(12, 1000, seed=12, trim=0, expansion=4)
    cases = [synth.unpack(packed, i) + (False, False) for i in range(12)]
    check_aligned_pairs(ctx, oracle, helpers.ModelSpec(cp.fiveState), p, cases, "1kb cli")


def test_long_pair_many_blocks(ctx, oracle):
    """20 kb pair: ~36 traceback blocks, forward values around -3e4 (the FP64 contract matters here)"""
    p = cp.pairwiseAlignmentBandingParameters_construct()
    packed = synth.evolved_pairs(2, 20000, seed=13, trim=int(p.constraintDiagonalTrim), expansion=int(p.diagonalExpansion))
    cases = [synth.unpack(packed, i) + (bool(i), bool(i)) for i in range(2)]
    check_aligned_pairs(ctx, oracle, helpers.ModelSpec(cp.fiveState), p, cases, "20kb")
    p.diagonalExpansion = 10
    p.constraintDiagonalTrim = 0
    packed = synth.evolved_pairs(2, 20000, seed=14, trim=0, expansion=10)
    cases = [synth.unpack(packed, i) + (False, False) for i in range(2)]
    check_aligned_pairs(ctx, oracle, helpers.ModelSpec(cp.threeState), p, cases, "20kb sm3 e10")


def test_wide_bands_every_width_class(ctx, oracle):
    """no anchors => the band is the whole matrix: widths 40..1100 exercise the multi-warp CTA classes"""
    rng = np.random.default_rng(5)
    p = cp.pairwiseAlignmentBandingParameters_construct()
    cases = []
    for L in (40, 70, 130, 300, 600, 1100):
        sX = synth.random_sequence(rng, L, acgt_only=True)
        sY = synth.evolve_like_reference(rng, sX)[: L + 20]
        cases.append((sX, sY, np.zeros((0, 3), dtype=np.int64), False, False))
    check_aligned_pairs(ctx, oracle, helpers.ModelSpec(cp.fiveState), p, cases, "wide")
    check_aligned_pairs(ctx, oracle, helpers.ModelSpec(cp.threeStateAsymmetric), p, cases[:4], "wide sm3")


def test_very_wide_band_has_no_limit(ctx, oracle):
    """a 2.6k x 2.6k matrix without anchors (band 2601 cells wide): the warp-per-region kernels have no width classes"""
    rng = np.random.default_rng(6)
    p = cp.pairwiseAlignmentBandingParameters_construct()
    p.splitMatrixBiggerThanThis = 1 << 40
    sX = synth.random_sequence(rng, 2600, acgt_only=True)
    sY = sX[:1300] + synth.random_sequence(rng, 7, acgt_only=True) + sX[1300:]
    check_aligned_pairs(ctx, oracle, helpers.ModelSpec(cp.fiveState), p, [(sX, sY, np.zeros((0, 3), dtype=np.int64), False, False)], "very wide")


def test_edge_cases(ctx, oracle):
    p = cp.pairwiseAlignmentBandingParameters_construct()
    e = np.zeros((0, 3), dtype=np.int64)
    cases = [("", "", e, False, False), ("", "ACGT", e, False, False), ("ACGT", "", e, True, True), ("A", "A", e, False, False),
             ("NNNN", "acgt", e, False, True), ("ACGTACGT", "ACGTACGT", np.array([[3, 3, 0]], dtype=np.int64), True, False)]
    check_aligned_pairs(ctx, oracle, helpers.ModelSpec(cp.fiveState), p, cases, "edge")
    check_aligned_pairs(ctx, oracle, helpers.ModelSpec(cp.threeState), p, cases, "edge sm3")
    # an empty batch is fine too
    b = cp.Batch(ctx, [], [], [])
    b.run(cp.stateMachine5_construct(), p, cp.MODE_ALIGNED_PAIRS)
    assert b.result_count(0) == 0


def test_chunking_is_invisible(oracle):
    """a tiny scratch budget forces one chunk per few pairs; results must not change"""
    c2 = cp.Context(0)
    p = cp.pairwiseAlignmentBandingParameters_construct()
    packed = synth.evolved_pairs(40, 300, seed=21, trim=4, expansion=8)
    p.diagonalExpansion = 8
    p.minDiagsBetweenTraceBack = 100
    m = cp.stateMachine5_construct()
    b = cp.Batch(c2, None, None, packed=packed)
    b.run(m, p, cp.MODE_ALIGNED_PAIRS)
    off1, tri1 = b.fetch_pairs(0)
    assert b.stats().nChunks == 1
    c2.set_scratch_budget(600 * 1024)
    b.run(m, p, cp.MODE_ALIGNED_PAIRS)
    off2, tri2 = b.fetch_pairs(0)
    assert b.stats().nChunks > 3
    assert np.array_equal(off1, off2) and np.array_equal(tri1, tri2)
    b.run(m, p, cp.MODE_EXPECTATIONS)
    pp2, tot2 = b.fetch_expectations()
    c2.set_scratch_budget(0)
    b.run(m, p, cp.MODE_EXPECTATIONS)
    pp1, tot1 = b.fetch_expectations()
    assert np.array_equal(pp1, pp2)
    np.testing.assert_allclose(tot1, tot2, rtol=1e-13)
    b.close()
    c2.close()


@pytest.mark.parametrize("type_", [cp.threeState, cp.fiveState])
def test_indel_posteriors(ctx, oracle, type_):
    """getAlignedPairsWithIndelsUsingAnchors (tests/pairwiseAlignerTest.c:867-942 uses the three-state machine)"""
    rng = np.random.default_rng(300 + type_)
    p = cp.pairwiseAlignmentBandingParameters_construct()
    p.minDiagsBetweenTraceBack = 60
    spec = helpers.ModelSpec(type_)
    cases = small_cases(rng, 30, 150, ragged=True)
    b = run_batch(ctx, spec, p, cases, cp.MODE_ALIGNED_PAIRS_INDELS)
    res = [b.fetch_pairs(k) for k in range(3)]
    res2 = [b.fetch_pairs(k, reference_order=True) for k in range(3)]
    for i, c in enumerate(cases):
        want = oracle.aligned_pairs_with_indels(spec.orc(), helpers.orc_params_from(p), c[0], c[1], c[2], c[3], c[4])
        for k in range(3):
            off, tri = res[k]
            compare_triples(tri[off[i]:off[i + 1]], want[k], "indel list %d case %d" % (k, i))
            assert np.array_equal(res2[k][1][off[i]:off[i + 1], 1:], want[k][:, 1:]), "indel list %d case %d: list order" % (k, i)
    b.close()


@pytest.mark.parametrize("type_", [cp.fiveState, cp.fiveStateAsymmetric, cp.threeState, cp.threeStateAsymmetric])
def test_expectations(ctx, oracle, type_):
    """getExpectationsUsingAnchors incl. the block-boundary quirk: several blocks per pair"""
    rng = np.random.default_rng(400 + type_)
    p = cp.pairwiseAlignmentBandingParameters_construct()
    p.diagonalExpansion = 10
    p.minDiagsBetweenTraceBack = 50
    p.traceBackDiagonals = 7
    spec = helpers.ModelSpec.random(rng, type_) if type_ % 2 else helpers.ModelSpec(type_)
    cases = small_cases(rng, 16, 200, ragged=True)
    b = run_batch(ctx, spec, p, cases, cp.MODE_EXPECTATIONS)
    pp, tot = b.fetch_expectations()
    want_tot = np.zeros_like(tot)
    for i, c in enumerate(cases):
        want = oracle.expectations(spec.orc(), helpers.orc_params_from(p), c[0], c[1], c[2], c[3], c[4])
        np.testing.assert_allclose(pp[i], want, rtol=EXPECT_RTOL, atol=1e-12, err_msg="expectations of case %d" % i)
        want_tot += want
    np.testing.assert_allclose(tot, want_tot, rtol=EXPECT_RTOL, atol=1e-12)
    b.close()


def test_expectations_2kb_em_defaults(ctx, oracle):
    """BASELINE config 4 shape: 2 kb pairs, cPecanEm's realign options (--diagonalExpansion=10 --splitMatrixBiggerThanThis=3000, cPecanEm.py:371)"""
    p = cp.pairwiseAlignmentBandingParameters_construct()
    p.diagonalExpansion = 10
    p.constraintDiagonalTrim = 0
    p.splitMatrixBiggerThanThis = 3000 * 3000
    packed = synth.evolved_pairs(6, 2000, seed=31, trim=0, expansion=10)
    for type_ in (cp.fiveState, cp.threeState):
        spec = helpers.ModelSpec(type_)
        b = cp.Batch(ctx, None, None, packed=packed)
        b.run(spec.cpb(), p, cp.MODE_EXPECTATIONS)
        pp, tot = b.fetch_expectations()
        for i in range(6):
            sx, sy, a = synth.unpack(packed, i)
            want = oracle.expectations(spec.orc(), helpers.orc_params_from(p), sx, sy, a)
            np.testing.assert_allclose(pp[i], want, rtol=EXPECT_RTOL, atol=1e-12)
        b.close()


def test_forward_probability_bit_exact(ctx, oracle):
    """computeForwardProbability is pure logAdd arithmetic: the GPU must reproduce every bit (test_computeForwardProbability, :1157-1188)"""
    rng = np.random.default_rng(55)
    p = cp.pairwiseAlignmentBandingParameters_construct()
    for type_ in (cp.threeState, cp.fiveState):
        spec = helpers.ModelSpec(type_)
        cases = small_cases(rng, 40, 150, ragged=True)
        cases = [(c[0], c[1], np.zeros((0, 3), dtype=np.int64), c[3], c[4]) for c in cases]
        b = run_batch(ctx, spec, p, cases, cp.MODE_FORWARD)
        got = b.fetch_forward()
        for i, c in enumerate(cases):
            want = oracle.forward_prob(spec.orc(), helpers.orc_params_from(p), c[0], c[1], c[2], c[3], c[4])
            assert got[i] == want, "case %d: %r vs %r" % (i, got[i], want)
            if len(c[0]) >= 10:
                same = cp.computeForwardProbability(c[0], c[0], [], p, spec.cpb(), c[3], c[4], ctx=ctx)
                assert -np.inf < got[i] <= same <= 0.0
        b.close()


def test_ragged_ends_recover_the_shifted_diagonal(ctx):
    """test_getAlignedPairsWithRaggedEnds (tests/pairwiseAlignerTest.c:676-715) without the poset filter:
    the best pair of every core base must lie on x + 100 == y"""
    rng = np.random.default_rng(9)
    p = cp.pairwiseAlignmentBandingParameters_construct()
    sM = cp.stateMachine5_construct()
    for _ in range(5):
        core = synth.random_sequence(rng, 100, acgt_only=True)
        sY = synth.random_sequence(rng, 100, acgt_only=True) + core + synth.random_sequence(rng, 100, acgt_only=True)
        pairs = np.array(cp.getAlignedPairsUsingAnchors(sM, core, sY, [], p, True, True, ctx=ctx))
        best = {}
        for pi, x, y in pairs:
            if x not in best or pi > best[x][0]:
                best[x] = (pi, y)
        on_diag = sum(1 for x, (pi, y) in best.items() if y == x + 100)
        assert on_diag >= 95


def test_roundtrip_properties_at_bench_scale(ctx):
    """size-independent properties on a bench-shaped batch: unique (x,y) per pair, pInt range, every x row sums to <= 1
    (within the logAdd approximation), identical results on a re-run (determinism)"""
    p = cp.pairwiseAlignmentBandingParameters_construct()
    packed = synth.evolved_pairs(512, 1000, seed=77, trim=int(p.constraintDiagonalTrim), expansion=int(p.diagonalExpansion))
    b = cp.Batch(ctx, None, None, packed=packed)
    m = cp.stateMachine5_construct()
    b.run(m, p, cp.MODE_ALIGNED_PAIRS)
    off, tri = b.fetch_pairs(0)
    tri = tri.copy()
    st = b.stats()
    assert st.nPairs == 512 and st.cells > 512 * 50000
    assert tri[:, 0].min() >= int(0.01 * 1e7) and tri[:, 0].max() <= 10000000
    for i in range(0, 512, 37):
        t = tri[off[i]:off[i + 1]]
        lX = packed["xOff"][i + 1] - packed["xOff"][i]
        lY = packed["yOff"][i + 1] - packed["yOff"][i]
        assert t[:, 1].min() >= 0 and t[:, 1].max() < lX and t[:, 2].min() >= 0 and t[:, 2].max() < lY
        key = t[:, 1].astype(np.int64) * (lY + 1) + t[:, 2]
        assert np.unique(key).size == key.size
        rows = np.bincount(t[:, 1], weights=t[:, 0] / 1e7, minlength=lX)
        assert rows.max() < 1.02 and (rows > 0.5).sum() > 0.8 * lX
    b.run(m, p, cp.MODE_ALIGNED_PAIRS)
    off2, tri2 = b.fetch_pairs(0)
    assert np.array_equal(off, off2) and np.array_equal(tri, tri2)
    b.close()


def test_100kb_pair_with_anchor_banding_and_splits(ctx, oracle):
    """BASELINE config 3 shape: one 100 kb evolved pair, lastz-style anchors, library defaults (~180 traceback blocks,
    forward values down to -1.4e5), plus a copy with a 4 kb unalignable insert so that getSplitPoints cuts the matrix"""
    p = cp.pairwiseAlignmentBandingParameters_construct()
    packed = synth.evolved_pairs(1, 100000, seed=3, trim=int(p.constraintDiagonalTrim), expansion=int(p.diagonalExpansion))
    sx, sy, a = synth.unpack(packed, 0)
    cases = [(sx, sy, a, False, False)]
    # second case: an insert in Y in the middle; anchors right of it shift by its length => a > 3000 x 3000 anchor gap
    rng = np.random.default_rng(4)
    half = 20000
    ins = synth.random_sequence(rng, 4000, acgt_only=True)
    sx2, sy2 = bytes(sx[:40000]).decode(), bytes(sy[:40000]).decode()
    a2 = np.asarray(a, dtype=np.int64).reshape(-1, 3)
    a2 = a2[(a2[:, 0] < 39000) & (a2[:, 1] < 39000)]
    cut_y = int(a2[a2[:, 0] < half][-1, 1]) + 1
    left = a2[a2[:, 0] < half - 2000]
    right = a2[a2[:, 0] >= half + 2000].copy()
    right[:, 1] += len(ins)
    sy2 = sy2[:cut_y] + ins + sy2[cut_y:]
    cases.append((sx2, sy2, np.concatenate([left, right]), False, False))
    n = check_aligned_pairs(ctx, oracle, helpers.ModelSpec(cp.fiveState), p, cases, "100kb")
    assert n > 100000
    b = run_batch(ctx, helpers.ModelSpec(cp.fiveState), p, cases[1:], cp.MODE_ALIGNED_PAIRS)
    assert b.stats().nRegions >= 2  # the insert really split the matrix
    b.close()


def test_all_pairs_of_10kb_sequences(ctx, oracle):
    """BASELINE config 5 shape (makeAllPairwiseAlignments, impl/multipleAligner.c:668-681): every pair of a small family of
    10 kb sequences evolved from one ancestor, aligned in one batch"""
    p = cp.pairwiseAlignmentBandingParameters_construct()
    fam = synth.evolved_pairs(4, 10000, seed=8, trim=int(p.constraintDiagonalTrim), expansion=int(p.diagonalExpansion))
    # members: the Y sequences of four pairs that share... each pair has its own ancestor, so build the family from one X
    anc, _, _ = synth.unpack(fam, 0)
    rng = np.random.default_rng(9)
    members = [synth.evolve_with_alignment(rng, anc) for _ in range(4)]
    cases = []
    for i in range(4):
        for j in range(i + 1, 4):
            a = synth.anchors_between(members[i], members[j], trim=int(p.constraintDiagonalTrim), expansion=int(p.diagonalExpansion))
            cases.append((members[i][0], members[j][0], a, False, False))
    n = check_aligned_pairs(ctx, oracle, helpers.ModelSpec(cp.fiveState), p, cases, "all pairs 10kb")
    assert n > 6 * 8000


def test_two_pass_forward_is_bit_identical(ctx, oracle, monkeypatch):
    """Long regions: one plane-less sweep leaves a checkpoint in front of every traceback block, then every block recomputes
    its own forward cells (engine.cu, FWD_BLOCKS).  Same arithmetic in the same order, so every result is identical to the
    one-pass run bit for bit -- and the one-pass run is what the other tests pin to the oracle."""
    rng = np.random.default_rng(808)
    p = cp.pairwiseAlignmentBandingParameters_construct()
    p.minDiagsBetweenTraceBack = 150  # many blocks per pair
    cases = []
    for length in (3000, 4500, 700, 0, 37, 2500):
        sX = synth.random_sequence(rng, length, acgt_only=True)
        sY, pos = synth.evolve_with_alignment(rng, sX, sub_rate=0.08, indel_rate=0.02)
        a = synth.anchors_between((sX, np.arange(len(sX))), (sY, pos), trim=int(p.constraintDiagonalTrim), expansion=int(p.diagonalExpansion))
        cases.append((sX, sY, a, bool(length % 2), bool(length % 3)))
    for type_ in (cp.fiveState, cp.threeState):
        spec = helpers.ModelSpec(type_)
        for mode in (cp.MODE_ALIGNED_PAIRS, cp.MODE_ALIGNED_PAIRS_INDELS, cp.MODE_EXPECTATIONS):
            res = []
            for flag in ("0", "1"):
                monkeypatch.setenv("CPB_TWO_PASS", flag)
                b = run_batch(ctx, spec, p, cases, mode)
                if mode == cp.MODE_EXPECTATIONS:
                    res.append([b.fetch_expectations()[0]])
                else:
                    res.append([np.concatenate([b.fetch_pairs(k)[0], b.fetch_pairs(k)[1].ravel()]) for k in range(1 if mode == cp.MODE_ALIGNED_PAIRS else 3)])
                assert int(b.stats().nBlocks) > 60
                b.close()
            for x, y in zip(*res):
                assert np.array_equal(x, y), "two-pass forward differs (type %d mode %d)" % (type_, mode)
    # and against the oracle directly
    monkeypatch.setenv("CPB_TWO_PASS", "1")
    check_aligned_pairs(ctx, oracle, helpers.ModelSpec(cp.fiveState), p, cases[:3], "two-pass")


def test_device_reweighting_and_alignment_scores(ctx):
    """reweightAlignedPairs2 (impl/pairwiseAligner.c:1519-1560) and getAlignmentScore
    (impl/multipleAligner.c:604-619) on the compacted device output, against the host list functions that
    tests/test_realign_host.py pins to the reference build.  Integer arithmetic: exact."""
    from test_realign_host import Host

    host = Host()
    p = cp.pairwiseAlignmentBandingParameters_construct()
    packed = synth.evolved_pairs(40, 300, seed=91, trim=int(p.constraintDiagonalTrim), expansion=int(p.diagonalExpansion))
    b = cp.Batch(ctx, None, None, packed=packed)
    b.run(cp.stateMachine5_construct(), p, cp.MODE_ALIGNED_PAIRS)
    off, before = b.fetch_pairs(0)
    before = before.copy()
    scores = b.alignment_scores()
    for gamma in (0.0, 0.5, 0.9):
        b.run(cp.stateMachine5_construct(), p, cp.MODE_ALIGNED_PAIRS)
        assert b.stats().reweighted == 0
        b.reweight_pairs(gamma)
        assert b.stats().reweighted == (1 if gamma > 0 else 0)
        if gamma > 0:
            # the weights were rewritten in place: a second call would apply the gap weighting twice, and is refused
            with pytest.raises(cp.CpbError, match="already been reweighted"):
                b.reweight_pairs(gamma)
        off2, after = b.fetch_pairs(0)
        assert np.array_equal(off, off2) and np.array_equal(before[:, 1:], after[:, 1:])
        for i in range(40):
            sx, sy, _ = synth.unpack(packed, i)
            mine = before[off[i]:off[i + 1]]
            want = host.reweight(mine, len(sx), len(sy), gamma)
            assert np.array_equal(after[off[i]:off[i + 1], 0], want[:, 0]), "pair %d gamma %g" % (i, gamma)
    for i in range(40):
        sx, sy, _ = synth.unpack(packed, i)
        d = float(before[off[i]:off[i + 1], 0].astype(np.int64).sum()) / (max(1, min(len(sx), len(sy))) * 1e7)
        assert scores[i] == int(min(max(d, 0.0), 1.0) * 1e7)
    b.close()


@pytest.mark.parametrize("type_", [cp.fiveState, cp.fiveStateAsymmetric, cp.threeState, cp.threeStateAsymmetric])
def test_models_with_impossible_transitions_and_emissions(ctx, oracle, type_):
    """An Hmm with zero probabilities has log(0) = LOG_ZERO in its state machine (impl/stateMachine.c:529-620).  Inside the kernels
    LOG_ZERO is a finite stand-in (kernels.cuh, CPB_FINITE_LOG_ZERO); it has to behave exactly as the reference's -infinity:
    same pair sets, forward log-probabilities bit for bit, expectations."""
    rng = np.random.default_rng(900 + type_)
    S = 5 if type_ < 2 else 3
    p = cp.pairwiseAlignmentBandingParameters_construct()
    p.minDiagsBetweenTraceBack = 80
    for rep in range(3):
        t = rng.random((S, S)) + 0.01
        e = rng.random((S, 16)) + 0.01
        if S == 5:
            t[0, 3] = t[0, 4] = 0.0          # the long-gap states cannot be entered from match ...
            if rep > 0:
                t[3, 3] = t[4, 4] = 0.0      # ... nor extended
        else:
            t[1, 2] = t[2, 1] = 0.0          # no gap switches
        if rep == 2:
            e[0, 5] = e[0, 10] = 0.0         # c/c and g/g never emitted by the match state
        t /= t.sum(1, keepdims=True)
        e /= e.sum(1, keepdims=True)
        spec = helpers.ModelSpec(type_, t, e)
        cases = small_cases(rng, 12, 140, ragged=rep == 1)
        check_aligned_pairs(ctx, oracle, spec, p, cases, "zero-probability model type %d rep %d" % (type_, rep))
        b = run_batch(ctx, spec, p, cases, cp.MODE_EXPECTATIONS)
        pp, _ = b.fetch_expectations()
        b.close()
        unanchored = [(c[0], c[1], np.zeros((0, 3), dtype=np.int64), c[3], c[4]) for c in cases]
        b = run_batch(ctx, spec, p, unanchored, cp.MODE_FORWARD)
        fw = b.fetch_forward()
        b.close()
        for i, c in enumerate(cases):
            want = oracle.expectations(spec.orc(), helpers.orc_params_from(p), c[0], c[1], c[2], c[3], c[4])
            np.testing.assert_allclose(pp[i], want, rtol=EXPECT_RTOL, atol=1e-12, err_msg="case %d" % i)
            u = unanchored[i]
            wf = oracle.forward_prob(spec.orc(), helpers.orc_params_from(p), u[0], u[1], u[2], u[3], u[4])
            assert float(fw[i]).hex() == float(wf).hex(), "forward log-probability case %d: %r vs %r" % (i, fw[i], wf)


def test_every_weight_through_the_host_recomputation_path(ctx, oracle, monkeypatch):
    """CPB_PINT_TOLERANCE=0.5 declares every kept cell's weight unsafe: all of them are recomputed with the host's libm and patched
    into the device lists (normally a few per 1e8).  The lists must still be the reference's, and the statistics must say so."""
    monkeypatch.setenv("CPB_PINT_TOLERANCE", "0.5")
    monkeypatch.setenv("CPB_PINT_FIXUP_CAP", "400000")
    rng = np.random.default_rng(77)
    p = cp.pairwiseAlignmentBandingParameters_construct()
    cases = small_cases(rng, 12, max_len=300)
    spec = helpers.ModelSpec(cp.fiveState)
    b = run_batch(ctx, spec, p, cases, cp.MODE_ALIGNED_PAIRS_INDELS)
    om, op = spec.orc(), helpers.orc_params_from(p)
    n = 0
    for which in range(3):
        off, tri = b.fetch_pairs(which)
        for i, c in enumerate(cases):
            want = oracle.aligned_pairs_with_indels(om, op, c[0], c[1], c[2], c[3], c[4])[which]
            compare_triples(tri[off[i]:off[i + 1]], want, "list %d case %d" % (which, i))
        n += int(off[-1])
    assert 0 < b.stats().pintFixups <= n  # all but the cells with p = 1 exactly (lp >= 0), which need no exp
    b.close()
    monkeypatch.setenv("CPB_PINT_FIXUP_CAP", "8")
    with pytest.raises(cp.CpbError, match="CPB_PINT_FIXUP_CAP"):
        run_batch(ctx, spec, p, cases, cp.MODE_ALIGNED_PAIRS)


def test_threshold_is_decided_in_log_space(ctx, oracle):
    """thresholds placed exactly on a computed posterior: p >= threshold must come out as the reference's libm says, for the cell
    itself and its neighbours in value"""
    rng = np.random.default_rng(78)
    p = cp.pairwiseAlignmentBandingParameters_construct()
    cases = small_cases(rng, 6, max_len=200)
    spec = helpers.ModelSpec(cp.fiveState)
    om = spec.orc()
    p.threshold = 0.0  # every band cell is kept (p >= 0), also the ones with p = 0
    total0 = check_aligned_pairs(ctx, oracle, spec, p, cases, "threshold 0")
    assert total0 > 0
    p.threshold = 0.01
    want = oracle.aligned_pairs(om, helpers.orc_params_from(p), cases[0][0], cases[0][1], cases[0][2], cases[0][3], cases[0][4])
    for w in sorted(set(int(v) for v in want[:, 0]))[:40]:
        for t in (w / 1e7, np.nextafter(w / 1e7, 1.0), np.nextafter(w / 1e7, 0.0), min((w + 1) / 1e7, 1.0)):
            p.threshold = float(t)
            check_aligned_pairs(ctx, oracle, spec, p, cases[:2], "threshold %r" % t)


def test_result_sink_delivers_the_same_lists(ctx, oracle):
    """cpb_batch_set_result_sink: the run copies every chunk's triples to a host buffer on a second stream while the next chunk
    computes; the lists must be what an ordinary fetch returns -- over several chunks, with weights the host had to recompute, and
    when the sink is too small (then the fetch copies as usual)."""
    import torch

    p = cp.pairwiseAlignmentBandingParameters_construct()
    packed = synth.evolved_pairs(48, 600, seed=21, trim=int(p.constraintDiagonalTrim), expansion=int(p.diagonalExpansion))
    spec = helpers.ModelSpec(cp.fiveState)
    b = cp.Batch(ctx, None, None, packed=packed)
    b.run(spec.cpb(), p, cp.MODE_ALIGNED_PAIRS)
    off0, tri0 = b.fetch_pairs(0)
    tri0 = tri0.copy()
    n = int(off0[-1])
    ctx.set_scratch_budget(48 << 20)  # a few chunks
    try:
        sink = torch.zeros((n + 100, 3), dtype=torch.int32).pin_memory().numpy()
        b.set_result_sink(0, sink)
        b.run(spec.cpb(), p, cp.MODE_ALIGNED_PAIRS)
        assert b.stats().nChunks > 1
        off1, tri1 = b.fetch_pairs(0, out=sink)
        assert np.array_equal(off0, off1) and np.array_equal(tri0, tri1)
        off2, tri2 = b.fetch_pairs(0)  # an ordinary fetch into another buffer still works
        assert np.array_equal(tri0, tri2)
        off3, tri3 = b.fetch_pairs(0, out=sink, reference_order=True)
        b.set_result_sink(0, None)
        b.run(spec.cpb(), p, cp.MODE_ALIGNED_PAIRS)
        assert np.array_equal(b.fetch_pairs(0, reference_order=True)[1], tri3)
        small = torch.zeros((n // 2, 3), dtype=torch.int32).pin_memory().numpy()
        b.set_result_sink(0, small)
        b.run(spec.cpb(), p, cp.MODE_ALIGNED_PAIRS)
        assert np.array_equal(b.fetch_pairs(0)[1], tri0)
    finally:
        ctx.set_scratch_budget(0)
        b.close()


def test_plan_is_kept_between_runs_with_the_same_parameters(ctx, oracle, monkeypatch):
    """A batch that is run again with the same mode, state count and banding parameters keeps its regions, bands, schedule,
    chunks and work lists (CpbRunStats.planReused) -- what every EM iteration of a resident batch does, with a new model each
    time.  The results must be those of a fresh batch; any change of the key plans anew."""
    rng = np.random.default_rng(77)
    cases = small_cases(rng, 24, max_len=260, ragged=True)
    p = cp.pairwiseAlignmentBandingParameters_construct()
    p.minDiagsBetweenTraceBack, p.traceBackDiagonals, p.splitMatrixBiggerThanThis = 60, 20, 40 * 40

    def fresh(spec, q, mode):
        b = run_batch(ctx, spec, q, cases, mode)
        assert b.stats().planReused == 0
        out = (b.fetch_pairs(0)[1].copy(), b.fetch_pairs(0)[0].copy()) if mode == cp.MODE_ALIGNED_PAIRS else b.fetch_expectations()
        b.close()
        return out

    five, three = helpers.ModelSpec(cp.fiveState), helpers.ModelSpec(cp.threeState)
    b = run_batch(ctx, five, p, cases, cp.MODE_ALIGNED_PAIRS)
    first = b.fetch_pairs(0)[1].copy()
    b.run(five.cpb(), p, cp.MODE_ALIGNED_PAIRS)
    assert b.stats().planReused == 1 and b.stats().cells > 0 and b.stats().nBlocks > 0
    assert np.array_equal(b.fetch_pairs(0)[1], first) and np.array_equal(first, fresh(five, p, cp.MODE_ALIGNED_PAIRS)[0])
    # another state count, another mode, other banding parameters: each plans anew, and gives what a fresh batch gives
    b.run(three.cpb(), p, cp.MODE_ALIGNED_PAIRS)
    assert b.stats().planReused == 0
    assert np.array_equal(b.fetch_pairs(0)[1], fresh(three, p, cp.MODE_ALIGNED_PAIRS)[0])
    b.run(three.cpb(), p, cp.MODE_EXPECTATIONS)
    assert b.stats().planReused == 0
    per0, tot0 = b.fetch_expectations()
    # EM: the same plan, a different model
    other = helpers.ModelSpec.random(rng, cp.threeState)
    b.run(other.cpb(), p, cp.MODE_EXPECTATIONS)
    assert b.stats().planReused == 1
    per1, tot1 = b.fetch_expectations()
    perF, totF = fresh(other, p, cp.MODE_EXPECTATIONS)
    assert np.array_equal(per1, perF, equal_nan=True) and np.array_equal(tot1, totF, equal_nan=True)
    assert not np.array_equal(per0, per1, equal_nan=True)
    q = cp.pairwiseAlignmentBandingParameters_construct()
    q.minDiagsBetweenTraceBack, q.traceBackDiagonals, q.splitMatrixBiggerThanThis, q.diagonalExpansion = 60, 20, 40 * 40, 12
    b.run(five.cpb(), q, cp.MODE_ALIGNED_PAIRS)
    assert b.stats().planReused == 0
    assert np.array_equal(b.fetch_pairs(0)[1], fresh(five, q, cp.MODE_ALIGNED_PAIRS)[0])
    b.run(five.cpb(), q, cp.MODE_ALIGNED_PAIRS)
    assert b.stats().planReused == 1
    monkeypatch.setenv("CPB_NO_PLAN_CACHE", "1")
    b.run(five.cpb(), q, cp.MODE_ALIGNED_PAIRS)
    assert b.stats().planReused == 0
    b.close()


def test_regions_from_device_split_flags(ctx, oracle):
    """The pairs are cut into regions where k_split_flags marks a gap between anchors (getSplitPoints,
    impl/pairwiseAligner.c:1230-1257): the number of regions must be what the host restatement counts, for every ragged-end
    combination, cut threshold and anchor density -- including pairs without anchors, empty sequences and cuts at the first and
    last gap -- and the aligned pairs those of the oracle's splitting run."""
    rng = np.random.default_rng(4242)
    cases = []
    for k in range(120):
        lX, lY = int(rng.integers(0, 90)), int(rng.integers(0, 90))
        sX, sY = synth.random_sequence(rng, lX), synth.random_sequence(rng, lY)
        a = synth.random_anchor_pairs(rng, lX, lY)
        if k % 3 == 0 and len(a) > 2:
            a = a[rng.random(len(a)) > 0.7]  # sparse anchors: wide gaps
        if k % 7 == 0:
            a = a[:0]
        cases.append((sX, sY, a, bool(k & 1), bool(k & 2)))
    spec = helpers.ModelSpec(cp.fiveState)
    for split in (0, 1, 9, 64, 400, 10 ** 9):
        p = cp.pairwiseAlignmentBandingParameters_construct()
        p.splitMatrixBiggerThanThis = split
        p.minDiagsBetweenTraceBack, p.traceBackDiagonals = 30, 10
        want = sum(len(cp.getSplitPoints(c[2], len(c[0]), len(c[1]), split, c[3], c[4])) for c in cases)
        b = run_batch(ctx, spec, p, cases, cp.MODE_ALIGNED_PAIRS)
        assert b.stats().nRegions == want, "split %d: %d regions, the host restatement has %d" % (split, b.stats().nRegions, want)
        b.close()
        check_aligned_pairs(ctx, oracle, spec, p, cases[:40], "split %d" % split)


@pytest.mark.parametrize("team", [2, 4, 16])
def test_forward_teams_are_bit_identical(ctx, oracle, monkeypatch, team):
    """A region's strips dealt to a team of warps and pipelined through the rings (FWD_TEAMS; the engine's choice for few, long
    regions) against one warp per region: forced here on short, ragged, split and empty problems -- teams larger than a region has
    strips, strips that are empty, several regions per team one after the other -- for aligned pairs, expectations and the forward
    probability."""
    monkeypatch.setenv("CPB_TEAM", str(team))
    rng = np.random.default_rng(600 + team)
    cases = small_cases(rng, 20, max_len=400, ragged=True)
    p = cp.pairwiseAlignmentBandingParameters_construct()
    p.minDiagsBetweenTraceBack, p.traceBackDiagonals, p.splitMatrixBiggerThanThis = 90, 30, 60 * 60
    for type_ in (cp.fiveState, cp.threeState):
        spec = helpers.ModelSpec(type_)
        check_aligned_pairs(ctx, oracle, spec, p, cases, "team %d" % team)
        b = run_batch(ctx, spec, p, cases, cp.MODE_FORWARD)
        got = b.fetch_forward()
        b.close()
        monkeypatch.delenv("CPB_TEAM")
        b = run_batch(ctx, spec, p, cases, cp.MODE_FORWARD)
        want = b.fetch_forward()
        b.close()
        monkeypatch.setenv("CPB_TEAM", str(team))
        assert np.array_equal(got, want, equal_nan=True), "forward probabilities differ with teams of %d" % team
