"""getAlignedPairs / getExpectations from raw sequences (no anchors passed): libcpecan.so anchors matrices bigger than
anchorMatrixBiggerThanThis in process (host/anchors.c) where the reference runs LASTZ in a subprocess.

LASTZ's own output is not a parity target; what must hold is that the banded DP placed by these anchors
finds what the DP placed by the TRUE alignment's anchors finds: the confident aligned pairs are the same, with the same weights.
"""
import ctypes as C
import os

import numpy as np
import pytest

import helpers
from cpecan_b200 import synth

pytestmark = pytest.mark.gpu
LIB = os.path.join(helpers.ROOT, "cpecan_b200", "lib", "libcpecan.so")


@pytest.fixture(scope="module")
def lib():
    L = C.CDLL(LIB)
    for f in ("stateMachine5_construct", "pairwiseAlignmentBandingParameters_construct", "getAlignedPairs", "stList_get", "hmm_constructEmpty"):
        getattr(L, f).restype = C.c_void_p
    for f in ("stList_length", "stIntTuple_get"):
        getattr(L, f).restype = C.c_int64
    L.getAlignedPairs.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_void_p, C.c_bool, C.c_bool]
    L.stList_get.argtypes = [C.c_void_p, C.c_int64]
    L.stList_length.argtypes = [C.c_void_p]
    L.stIntTuple_get.argtypes = [C.c_void_p, C.c_int64]
    L.stList_destruct.argtypes = [C.c_void_p]
    yield L
    L.cpecan_shutdown()


def _triples(L, l):
    n = L.stList_length(l)
    out = np.array([[L.stIntTuple_get(L.stList_get(l, i), k) for k in range(3)] for i in range(n)], dtype=np.int64).reshape(-1, 3)
    L.stList_destruct(l)
    return out


@pytest.mark.parametrize("length", [2000, 12000])
def test_aligned_pairs_from_raw_sequences(lib, oracle, length):
    import cpecan_b200 as cp

    p_ref = cp.pairwiseAlignmentBandingParameters_construct()
    packed = synth.evolved_pairs(2, length, seed=40 + length, trim=int(p_ref.constraintDiagonalTrim), expansion=int(p_ref.diagonalExpansion))
    sM = lib.stateMachine5_construct(0)
    p = lib.pairwiseAlignmentBandingParameters_construct()
    om, op = helpers.ModelSpec(0).orc(), helpers.orc_params_from(p_ref)
    for i in range(2):
        sx, sy, a = synth.unpack(packed, i)
        assert len(sx) * len(sy) > p_ref.anchorMatrixBiggerThanThis
        got = _triples(lib, lib.getAlignedPairs(sM, sx, sy, p, False, False))
        want = oracle.aligned_pairs(om, op, sx, sy, a)  # the reference's DP, banded by the true alignment's anchors
        g = {(int(x), int(y)): int(w) for w, x, y in got}
        w = {(int(x), int(y)): int(v) for v, x, y in want}
        confident = [k for k, v in w.items() if v >= 5000000]
        assert len(confident) > 0.7 * length
        missing = [k for k in confident if k not in g]
        assert not missing, "%d of %d confident pairs are missing" % (len(missing), len(confident))
        # inside both bands the posteriors are the same DP: weights agree to a few 1e-7 except where one band clips probability mass
        diff = np.array([abs(g[k] - w[k]) for k in confident])
        print("weight differences between the two bands: median %d, 99th percentile %d, max %d" % (np.median(diff), np.percentile(diff, 99), diff.max()))
        assert np.median(diff) <= 100 and np.percentile(diff, 99) <= 100000, (np.median(diff), np.percentile(diff, 99), diff.max())
        # and nothing confident appears that the true-anchor run does not know at all
        extra = [k for k, v in g.items() if v >= 5000000 and k not in w]
        assert not extra
