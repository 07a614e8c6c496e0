"""cpb_batch_create / cpb_batch_run reject malformed input instead of building a wrong band.

The reference states these preconditions as asserts (impl/pairwiseAligner.c:156-166, :214-219: anchors inside the matrix, strictly
increasing in x and in y, expansions non-negative and -- for dynamic bands -- even).  The flat C ABI checks them once, on the host,
while it converts the caller's int64 triples, and names the offending pair.
"""
import numpy as np
import pytest

import cpecan_b200 as cp

pytestmark = pytest.mark.gpu


def _packed(anchors, aOff=None, xOff=(0, 8), yOff=(0, 8)):
    a = np.asarray(anchors, dtype=np.int64).reshape(-1)
    return dict(seqX=np.frombuffer(b"ACGTACGT" * 4, dtype=np.uint8).copy(), xOff=np.asarray(xOff, dtype=np.int64),
                seqY=np.frombuffer(b"ACGTACGT" * 4, dtype=np.uint8).copy(), yOff=np.asarray(yOff, dtype=np.int64),
                anchors=a if a.size else np.zeros(3, dtype=np.int64),
                aOff=np.asarray(aOff if aOff is not None else (0, a.size // 3), dtype=np.int64))


@pytest.mark.parametrize("anchors, what", [
    ([(2, 2, 0), (2, 3, 0)], "x not increasing"),
    ([(2, 3, 0), (3, 3, 0)], "y not increasing"),
    ([(3, 3, 0), (2, 2, 0)], "descending"),
    ([(-1, 0, 0)], "negative x"),
    ([(0, 8, 0)], "y == lY"),
    ([(8, 0, 0)], "x == lX"),
    ([(1, 1, -2)], "negative expansion"),
    ([(1 << 33, 1, 0)], "does not fit int32"),
])
def test_malformed_anchors_are_rejected(ctx, anchors, what):
    with pytest.raises(cp.CpbError, match="pair 0"):
        cp.Batch(ctx, None, None, packed=_packed(anchors))


def test_the_offending_pair_is_named(ctx):
    pk = _packed([(1, 1, 0), (2, 2, 0), (5, 5, 0), (4, 6, 0)], aOff=(0, 2, 4), xOff=(0, 8, 16), yOff=(0, 8, 16))
    with pytest.raises(cp.CpbError, match="pair 1"):
        cp.Batch(ctx, None, None, packed=pk)


def test_offsets_must_be_monotone_and_start_at_zero(ctx):
    with pytest.raises(cp.CpbError, match="pair 0"):
        cp.Batch(ctx, None, None, packed=_packed([], xOff=(0, -1)))
    with pytest.raises(cp.CpbError, match="start at 0"):
        cp.Batch(ctx, None, None, packed=_packed([], xOff=(1, 8)))
    with pytest.raises(cp.CpbError, match="pair 1"):
        cp.Batch(ctx, None, None, packed=_packed([(1, 1, 0)], aOff=(0, 1, 0), xOff=(0, 8, 16), yOff=(0, 8, 16)))


def test_odd_expansions_only_matter_for_dynamic_bands(ctx):
    b = cp.Batch(ctx, None, None, packed=_packed([(1, 1, 3), (4, 4, 3)]))
    p = cp.pairwiseAlignmentBandingParameters_construct()
    b.run(cp.stateMachine5_construct(), p, cp.MODE_ALIGNED_PAIRS)  # static band: the per-anchor expansion is not used
    p.dynamicAnchorExpansion = 1
    with pytest.raises(cp.CpbError, match="expansion must be even"):
        b.run(cp.stateMachine5_construct(), p, cp.MODE_ALIGNED_PAIRS)
    b.close()


def test_valid_input_still_passes_after_a_rejection(ctx):
    with pytest.raises(cp.CpbError):
        cp.Batch(ctx, None, None, packed=_packed([(3, 3, 0), (2, 2, 0)]))
    b = cp.Batch(ctx, None, None, packed=_packed([(2, 2, 0), (3, 3, 0)]))
    b.run(cp.stateMachine5_construct(), cp.pairwiseAlignmentBandingParameters_construct(), cp.MODE_ALIGNED_PAIRS)
    off, tri = b.fetch_pairs(0)
    assert off[-1] == tri.shape[0] > 0
    b.close()
