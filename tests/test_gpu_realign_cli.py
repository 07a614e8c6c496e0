"""cPecanRealign (cpecan_b200/csrc/host/cPecanRealign.c) end to end on the GPU: FASTA + cigars in, cigars / expectations out,
against the same pipeline assembled from the oracle's aligned pairs (the reference's DP) and the host list functions that
tests/test_realign_host.py pins to the reference (cPecanRealign.c:509-605 of the reference is the loop being replaced)."""
import os
import subprocess

import numpy as np
import pytest

import cpecan_b200 as cp
from cpecan_b200 import synth
import helpers
from test_realign_host import Host

pytestmark = pytest.mark.gpu

ROOT = helpers.ROOT
EXE = os.path.join(ROOT, "cpecan_b200", "lib", "cPecanRealign")
COMP = str.maketrans("ACGTacgt", "TGCAtgca")


def revcomp(s):
    return s.translate(COMP)[::-1]


def ops_from_columns(cols, l1, l2):
    """cigar operations [(type, length)] (0 match, 1 indel X, 2 indel Y) of an alignment given as increasing (x, y) columns"""
    ops = []

    def add(t, n):
        if n > 0:
            if ops and ops[-1][0] == t:
                ops[-1][1] += n
            else:
                ops.append([t, n])

    px, py = -1, -1
    for x, y in list(cols) + [(l1, l2)]:
        add(1, x - px - 1)
        add(2, y - py - 1)
        if x < l1:
            add(0, 1)
        px, py = x, y
    return [tuple(o) for o in ops]


def cigar_line(name1, s1, e1, strand1, name2, s2, e2, strand2, score, ops):
    return "cigar: %s %d %d %s %s %d %d %s %f%s" % (name2, s2, e2, "+" if strand2 else "-", name1, s1, e1, "+" if strand1 else "-", score,
                                                  "".join(" %s %d" % ("MDI"[t], n) for t, n in ops))


def make_inputs(rng, n, length):
    """n local alignments between sub-ranges of sequences X_i and Y_i; odd ones have sequence 2 on the minus strand"""
    fasta, jobs = [], []
    for i in range(n):
        core = synth.random_sequence(rng, length, acgt_only=True)
        coreY, pos = synth.evolve_with_alignment(rng, core, sub_rate=0.08, indel_rate=0.02)
        cols = [(x, int(pos[x])) for x in range(len(core)) if pos[x] >= 0]
        # the alignment must start and end with a column for the cigar to be a local alignment of exactly [0, len) of both
        x0, y0 = cols[0]
        x1, y1 = cols[-1]
        subX, subY = core[x0:x1 + 1], coreY[y0:y1 + 1]
        cols = [(x - x0, y - y0) for x, y in cols]
        padX = [synth.random_sequence(rng, int(rng.integers(0, 30)), acgt_only=True) for _ in range(2)]
        padY = [synth.random_sequence(rng, int(rng.integers(0, 30)), acgt_only=True) for _ in range(2)]
        minus = i % 2 == 1
        seqX = padX[0] + subX + padX[1]
        seqY = padY[0] + (revcomp(subY) if minus else subY) + padY[1]
        fasta.append((">x%d some description" % i, seqX))
        fasta.append((">y%d" % i, seqY))
        s1, e1 = len(padX[0]), len(padX[0]) + len(subX)
        lo2, hi2 = len(padY[0]), len(padY[0]) + len(subY)
        s2, e2 = (hi2, lo2) if minus else (lo2, hi2)
        jobs.append(dict(name1="x%d" % i, name2="y%d" % i, s1=s1, e1=e1, s2=s2, e2=e2, strand2=0 if minus else 1, subX=subX, subY=subY, cols=cols,
                         ops=ops_from_columns(cols, len(subX), len(subY))))
    return fasta, jobs


def run_cli(tmp_path, fasta, jobs, args):
    fa = tmp_path / "seqs.fa"
    fa.write_text("".join("%s\n%s\n" % (h, "\n".join(s[k:k + 70] for k in range(0, len(s), 70))) for h, s in fasta))
    cig = "".join(cigar_line(j["name1"], j["s1"], j["e1"], 1, j["name2"], j["s2"], j["e2"], j["strand2"], 1.0, j["ops"]) + "\n" for j in jobs)
    out = subprocess.run([EXE] + args + [str(fa)], input=cig, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    return out.stdout.splitlines()


def cli_params():
    p = cp.pairwiseAlignmentBandingParameters_construct()  # cPecanRealign.c:359-361
    p.constraintDiagonalTrim, p.splitMatrixBiggerThanThis, p.diagonalExpansion = 0, 10, 4
    return p


def oracle_pairs(oracle, job, p):
    """what getAlignedPairsUsingAnchors gives the CLI for this job: anchors = matching columns of the input alignment"""
    sX, sY = job["subX"], job["subY"]
    anchors = np.array([(x, y, int(p.diagonalExpansion)) for x, y in job["cols"] if sX[x].upper() == sY[y].upper() and sX[x].upper() != "N"],
                       dtype=np.int64).reshape(-1, 3)
    return oracle.aligned_pairs(helpers.ModelSpec(cp.fiveState).orc(), helpers.orc_params_from(p), sX, sY, anchors, True, True)


def expected_line(job, score, cols):
    lX, lY = len(job["subX"]), len(job["subY"])
    ops = ops_from_columns(cols, lX, lY)
    if job["strand2"]:
        return cigar_line(job["name1"], job["s1"], job["e1"], 1, job["name2"], job["s2"], job["e2"], 1, score, ops)
    return cigar_line(job["name1"], job["s1"], job["e1"], 1, job["name2"], job["s2"], job["e2"], 0, score, ops)


def test_realign_defaults(tmp_path):
    oracle = helpers.best_oracle()
    host = Host()
    rng = np.random.default_rng(42)
    fasta, jobs = make_inputs(rng, 12, 300)
    got = run_cli(tmp_path, fasta, jobs, ["--batchBases", "2000"])  # several device passes
    assert len(got) == len(jobs)
    p = cli_params()
    for j, line in zip(jobs, got):
        pairs = oracle_pairs(oracle, j, p)
        rw = host.reweight(pairs, len(j["subX"]), len(j["subY"]), 0.5)
        chain = host.chain(rw, j["subX"], j["subY"], 0.85)
        assert line == expected_line(j, 1.0, [(int(t[1]), int(t[2])) for t in chain]), j["name1"]


def test_rescore_original_alignment_by_posterior_prob(tmp_path):
    oracle = helpers.best_oracle()
    host = Host()
    rng = np.random.default_rng(43)
    fasta, jobs = make_inputs(rng, 6, 200)
    got = run_cli(tmp_path, fasta, jobs, ["--rescoreOriginalAlignment", "--rescoreByPosteriorProb"])
    p = cli_params()
    for j, line in zip(jobs, got):
        pairs = oracle_pairs(oracle, j, p)
        weight = {(int(t[1]), int(t[2])): int(t[0]) for t in pairs}
        scored = np.array([(weight.get(c, 0), c[0], c[1]) for c in j["cols"]], dtype=np.int64)
        score = host.scores(j["subX"], j["subY"], scored)[2]
        want = expected_line(j, score, j["cols"])  # the output alignment is the input alignment
        head = lambda s: s.split()[:9] + s.split()[10:]  # noqa: E731  (everything but the score)
        assert head(line) == head(want)
        assert abs(float(line.split()[9]) - score) < 2e-5  # pInt may differ by one unit of 1e-7 per pair


def test_expectations_file(tmp_path):
    oracle = helpers.best_oracle()
    rng = np.random.default_rng(44)
    fasta, jobs = make_inputs(rng, 8, 250)
    hmm_file = tmp_path / "exp.hmm"
    got = run_cli(tmp_path, fasta, jobs, ["--outputExpectations", str(hmm_file), "--diagonalExpansion", "10"])
    assert got == []
    p = cli_params()
    p.diagonalExpansion = 10
    want = np.zeros(cp.hmm_len(5))
    spec = helpers.ModelSpec(cp.fiveState)
    for j in jobs:
        sX, sY = j["subX"], j["subY"]
        anchors = np.array([(x, y, 10) for x, y in j["cols"] if sX[x] == sY[y]], dtype=np.int64).reshape(-1, 3)
        want += oracle.expectations(spec.orc(), helpers.orc_params_from(p), sX, sY, anchors, True, True)
    lines = hmm_file.read_text().splitlines()
    head = lines[0].split()
    assert int(head[0]) == cp.fiveState
    trans = np.array([float(v) for v in head[1:26]])
    emis = np.array([float(v) for v in lines[1].split()])
    np.testing.assert_allclose(trans, want[:25] + 1e-12, rtol=1e-6, atol=2e-6)  # hmm_write prints %f
    np.testing.assert_allclose(emis, want[25:105] + 1e-12, rtol=1e-6, atol=2e-6)
    assert abs(float(head[26]) - want[105]) < 1e-4 * max(1.0, abs(want[105]))


def test_split_long_indels_and_posterior_file(tmp_path):
    rng = np.random.default_rng(45)
    fasta, jobs = make_inputs(rng, 3, 200)
    probs = tmp_path / "probs.tsv"
    got = run_cli(tmp_path, fasta, jobs, ["--splitIndelsLongerThanThis", "2", "--outputPosteriorProbs", str(probs)])
    assert len(got) >= len(jobs)
    for line in got:
        f = line.split()
        ops = list(zip(f[10::2], [int(v) for v in f[11::2]]))
        assert ops[0][0] == "M" and ops[-1][0] == "M"  # pieces never start or end with an indel
        run = 0
        for t, n in ops:
            run = 0 if t == "M" else run + n
            assert run <= 2
        assert abs(int(f[3]) - int(f[2])) == sum(n for t, n in ops if t != "D")  # sequence 2 comes first on the line
        assert abs(int(f[7]) - int(f[6])) == sum(n for t, n in ops if t != "I")
    # the file holds the last alignment's pairs in original coordinates (the reference reopens it per alignment, cPecanRealign.c:305)
    j = jobs[-1]
    rows = [r.split("\t") for r in probs.read_text().splitlines()]
    assert rows and all(j["s1"] <= int(r[0]) < j["e1"] and 0.0 < float(r[2]) <= 1.0 for r in rows)
    lo2, hi2 = min(j["s2"], j["e2"]), max(j["s2"], j["e2"])
    assert all(lo2 <= int(r[1]) < hi2 for r in rows)


def test_missing_sequence_is_an_error(tmp_path):
    fa = tmp_path / "s.fa"
    fa.write_text(">a\nACGT\n")
    out = subprocess.run([EXE, str(fa)], input="cigar: b 0 4 + a 0 4 + 0 M 4\n", capture_output=True, text=True)
    assert out.returncode != 0 and "no sequence named b" in out.stderr


# ---- cPecanEm: the EM trainer over the same inputs ----
EM_EXE = os.path.join(ROOT, "cpecan_b200", "lib", "cPecanEm")


def read_model(path):
    lines = open(path).read().splitlines()
    head = lines[0].split()
    S = 5 if int(head[0]) < 2 else 3
    return dict(type=int(head[0]), transitions=np.array([float(v) for v in head[1:1 + S * S]]), likelihood=float(head[1 + S * S]),
                emissions=np.array([float(v) for v in lines[1].split()]), running=[float(v) for v in lines[2].split()] if len(lines) > 2 else [])


def em_params():
    p = cp.pairwiseAlignmentBandingParameters_construct()  # cPecanRealign's defaults, then cPecanEm.py:371
    p.constraintDiagonalTrim, p.diagonalExpansion, p.splitMatrixBiggerThanThis = 0, 10, 3000 * 3000
    return p


def run_em(tmp_path, fasta, jobs, args):
    fa = tmp_path / "seqs.fa"
    fa.write_text("".join("%s\n%s\n" % (h, s) for h, s in fasta))
    cig = "".join(cigar_line(j["name1"], j["s1"], j["e1"], 1, j["name2"], j["s2"], j["e2"], j["strand2"], 1.0, j["ops"]) + "\n" for j in jobs)
    out = subprocess.run([EM_EXE] + args + [str(fa)], input=cig, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    return out


@pytest.mark.parametrize("type_,name", [(cp.fiveState, "fiveState"), (cp.threeState, "threeState")])
def test_em_iterations_match_the_oracle(tmp_path, type_, name):
    """three iterations from the all-equal model, emissions trained: expectation pass (oracle: the reference's
    getExpectationsUsingAnchors) + normalise, as cPecanEm.py:176-209 does through expectation files"""
    from cpecan_b200 import sharding

    oracle = helpers.best_oracle()
    rng = np.random.default_rng(46)
    fasta, jobs = make_inputs(rng, 10, 250)
    model = tmp_path / "out.hmm"
    run_em(tmp_path, fasta, jobs, ["--iterations", "3", "--trainEmissions", "--modelType", name, "--outputModel", str(model)])
    got = read_model(model)
    S = 5 if type_ == cp.fiveState else 3
    p = em_params()
    trans, emis, running = np.full(S * S, 1.0 / S), np.full(S * 16, 1.0 / 16), []
    for _ in range(3):
        spec = helpers.ModelSpec(type_, trans.reshape(S, S), emis.reshape(S, 16))
        total = np.zeros(cp.hmm_len(S))
        for j in jobs:
            sX, sY = j["subX"], j["subY"]
            anchors = np.array([(x, y, 10) for x, y in j["cols"] if sX[x] == sY[y]], dtype=np.int64).reshape(-1, 3)
            total += oracle.expectations(spec.orc(), helpers.orc_params_from(p), sX, sY, anchors, True, True)
        total[:-1] += 1e-12
        running.append(total[-1])
        hmm = sharding.normalise_hmm(total, S)
        trans, emis = hmm[:S * S], hmm[S * S:S * S + 16 * S]
    assert got["type"] == type_
    np.testing.assert_allclose(got["transitions"], trans, rtol=1e-7, atol=1e-12)
    np.testing.assert_allclose(got["emissions"], emis, rtol=1e-7, atol=1e-12)
    np.testing.assert_allclose(got["running"], running, rtol=1e-9)
    assert running[2] > running[1] > running[0]  # EM increases the likelihood (tests/pairwiseAlignerTest.c:1091-1155)


def test_em_random_trials_keep_the_best(tmp_path):
    rng = np.random.default_rng(47)
    fasta, jobs = make_inputs(rng, 6, 200)
    model = tmp_path / "best.hmm"
    run_em(tmp_path, fasta, jobs, ["--iterations", "2", "--randomStart", "--trials", "3", "--seed", "7", "--outputTrialHmms", "--outputModel", str(model)])
    trials = [read_model(str(model) + "_%d" % i) for i in range(3)]
    best = max(trials, key=lambda m: m["likelihood"])
    got = read_model(model)
    assert got["likelihood"] == best["likelihood"] and np.array_equal(got["transitions"], best["transitions"])
    assert len({m["likelihood"] for m in trials}) == 3  # different random starts
    # emissions are not trained by default: they stay the random start's (cPecanEm.py:200-202)
    assert all(len(m["running"]) == 2 for m in trials)
