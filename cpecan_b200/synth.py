"""Synthetic benchmark / test inputs in the spirit of cPecan's impl/randomSequences.c.

``evolved_pairs`` makes the "randomSequences-style evolved pairs" of BASELINE.json: X is uniform ACGT
(getRandomACGTSequence, impl/randomSequences.c:38-45); Y is X with per-base substitutions and short
indel events (about 10 % divergence at the defaults); anchors are what a
lastz-style chain would hand over (convertPairwiseForwardStrandAlignmentToAnchorPairs,
impl/pairwiseAligner.c:979-1003): every column of each maximal run of identical aligned columns,
trimmed by ``trim`` on both sides, each with ``expansion``.  Fully vectorised numpy, seeded.

``evolve_like_reference`` follows evolveSequence (impl/randomSequences.c:50-73) for small test cases.
"""
import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def evolved_pairs(n_pairs, length, seed=0xC0FFEE, sub_rate=0.08, indel_rate=0.02, max_indel=4, trim=14, expansion=20,
                  batch=4096, n_rate=0.0):
    """Returns a packed dict: seqX, xOff, seqY, yOff (uint8 ASCII + int64 offsets), anchors (flat int64 triples), aOff."""
    rng = np.random.default_rng(seed)
    xs, ys, xoffs, yoffs, ancs, aoffs = [], [], [0], [0], [], [0]
    for start in range(0, n_pairs, batch):
        n = min(batch, n_pairs - start)
        L = length
        x = rng.integers(0, 4, size=(n, L), dtype=np.int8)
        sub = rng.random((n, L)) < sub_rate
        newbase = ((x + 1 + rng.integers(0, 3, size=(n, L), dtype=np.int8)) % 4).astype(np.int8)
        ev = rng.random((n, L))
        is_del = ev < indel_rate / 2
        is_ins = (ev >= indel_rate / 2) & (ev < indel_rate)
        ev_len = rng.integers(1, max_indel + 1, size=(n, L), dtype=np.int8)
        # deleted positions: union of [i, i+len) for deletion events (clipped to the row)
        deleted = np.zeros((n, L), dtype=bool)
        for k in range(max_indel):
            m = is_del & (ev_len > k)
            if k == 0:
                deleted |= m
            else:
                deleted[:, k:] |= m[:, :-k]
        ins_len = np.where(is_ins, ev_len, 0).astype(np.int64)  # inserted after position i
        emit = (~deleted).astype(np.int64) + ins_len
        ybase = np.where(sub, newbase, x)
        # build Y rows (ragged) in one flat pass
        flat_emit = emit.reshape(-1)
        total = int(flat_emit.sum())
        src = np.repeat(np.arange(n * L, dtype=np.int64), flat_emit)  # source X position for every Y base
        first = np.ones(total, dtype=bool)
        first[1:] = src[1:] != src[:-1]
        kept_flat = (~deleted).reshape(-1)
        from_x = first & kept_flat[src]  # first emitted base of a kept position is the aligned base; the rest are insertions
        ychars = np.where(from_x, ybase.reshape(-1)[src], rng.integers(0, 4, size=total, dtype=np.int8))
        ylen = emit.sum(axis=1)
        ystart = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(ylen, out=ystart[1:])
        # y coordinate (within the pair) of every kept X position
        ypos_flat = np.cumsum(flat_emit) - flat_emit  # index in flat Y of the first base emitted at this position
        ycoord = (ypos_flat.reshape(n, L) - ystart[:-1, None])
        # identical aligned columns and maximal diagonal runs
        col = (~deleted) & (~sub)
        brk = np.ones((n, L), dtype=bool)
        # column i continues the run of i-1 iff both are identical aligned columns and nothing was inserted after i-1
        brk[:, 1:] = ~(col[:, 1:] & col[:, :-1] & (ins_len[:, :-1] == 0))
        run_id = np.cumsum(brk.reshape(-1)) - 1
        run_len = np.bincount(run_id, weights=col.reshape(-1).astype(np.float64)).astype(np.int64)
        run_start = np.zeros(run_len.size, dtype=np.int64)
        run_start[run_id[brk.reshape(-1)]] = np.nonzero(brk.reshape(-1))[0]
        pos_in_run = np.arange(n * L, dtype=np.int64) - run_start[run_id]
        keep = col.reshape(-1) & (pos_in_run >= trim) & (pos_in_run < run_len[run_id] - trim)
        idx = np.nonzero(keep)[0]
        ax = idx % L
        ay = ycoord.reshape(-1)[idx]
        pair_of = idx // L
        a = np.stack([ax, ay, np.full(idx.size, expansion, dtype=np.int64)], axis=1).astype(np.int64)
        cnt = np.bincount(pair_of, minlength=n)
        if n_rate > 0.0:
            x = np.where(rng.random((n, L)) < n_rate, 4, x)
        xs.append(np.concatenate([_ACGT, np.frombuffer(b"N", dtype=np.uint8)])[x.reshape(-1)])
        ys.append(_ACGT[ychars])
        xoffs.extend((xoffs[-1] + L * (np.arange(n) + 1)).tolist())
        yoffs.extend((yoffs[-1] + ystart[1:]).tolist())
        ancs.append(a.reshape(-1))
        aoffs.extend((aoffs[-1] + np.cumsum(cnt)).tolist())
    anchors = np.concatenate(ancs) if ancs else np.zeros(0, dtype=np.int64)
    if anchors.size == 0:
        anchors = np.zeros(3, dtype=np.int64)
    return dict(
        seqX=np.ascontiguousarray(np.concatenate(xs + [np.zeros(1, dtype=np.uint8)])),
        xOff=np.asarray(xoffs, dtype=np.int64),
        seqY=np.ascontiguousarray(np.concatenate(ys + [np.zeros(1, dtype=np.uint8)])),
        yOff=np.asarray(yoffs, dtype=np.int64),
        anchors=np.ascontiguousarray(anchors, dtype=np.int64),
        aOff=np.asarray(aoffs, dtype=np.int64),
    )


def unpack(packed, i):
    """-> (sX bytes, sY bytes, anchors int64[k,3]) of pair i"""
    sx = packed["seqX"][packed["xOff"][i]:packed["xOff"][i + 1]].tobytes()
    sy = packed["seqY"][packed["yOff"][i]:packed["yOff"][i + 1]].tobytes()
    a = packed["anchors"][3 * packed["aOff"][i]:3 * packed["aOff"][i + 1]].reshape(-1, 3)
    return sx, sy, a


def subset(packed, idx):
    """Packed dict holding only the pairs in idx (in that order)."""
    sx, sy, an = [], [], []
    xo, yo, ao = [0], [0], [0]
    for i in idx:
        a, b, c = unpack(packed, i)
        sx.append(a)
        sy.append(b)
        an.append(c.reshape(-1))
        xo.append(xo[-1] + len(a))
        yo.append(yo[-1] + len(b))
        ao.append(ao[-1] + c.shape[0])
    anchors = np.concatenate(an) if an and ao[-1] > 0 else np.zeros(3, dtype=np.int64)
    return dict(seqX=np.frombuffer(b"".join(sx) + b"\0", dtype=np.uint8).copy(), xOff=np.asarray(xo, dtype=np.int64),
                seqY=np.frombuffer(b"".join(sy) + b"\0", dtype=np.uint8).copy(), yOff=np.asarray(yo, dtype=np.int64),
                anchors=np.ascontiguousarray(anchors, dtype=np.int64), aOff=np.asarray(ao, dtype=np.int64))


# ---- small, reference-shaped random inputs for the property tests ----
_RANDOM_CHARS = "AaCcGgTt" * 11 + "N"  # getRandomChar, impl/randomSequences.c:13-16


def random_sequence(rng, length, acgt_only=False):
    alphabet = "ACGT" if acgt_only else _RANDOM_CHARS
    return "".join(alphabet[i] for i in rng.integers(0, len(alphabet), size=length))


def evolve_like_reference(rng, seq):
    """evolveSequence, impl/randomSequences.c:50-73: 20 % substitutions, then global replace-style indels."""
    s = list(seq)
    for i in range(len(s)):
        if rng.random() > 0.8:
            s[i] = _RANDOM_CHARS[rng.integers(0, len(_RANDOM_CHARS))]
    s = "".join(s)
    while rng.random() > 0.2:
        a = random_sequence(rng, int(rng.integers(2, 4)))
        b = random_sequence(rng, int(rng.integers(0, 10)))
        s = s.replace(a, b)
    return s


def random_anchor_pairs(rng, lX, lY):
    """getRandomAnchorPairs, tests/pairwiseAlignerTest.c:326-342"""
    out = []
    x = y = -1
    while True:
        x += int(rng.integers(1, 20))
        y += int(rng.integers(1, 20))
        e = 2 * int(rng.integers(0, 5))
        if x >= lX or y >= lY:
            break
        out.append((x, y, e))
    return np.asarray(out, dtype=np.int64).reshape(-1, 3)


# ---- families of sequences evolved from one ancestor (the all-pairs shape of makeAllPairwiseAlignments) ----
def evolve_with_alignment(rng, ancestor, sub_rate=0.04, indel_rate=0.01, max_indel=4):
    """-> (sequence str, pos int64[len(ancestor)]): pos[i] = index of ancestor base i in the sequence, or -1 if deleted."""
    anc = ancestor.decode() if isinstance(ancestor, (bytes, bytearray)) else ancestor
    out, pos = [], np.full(len(anc), -1, dtype=np.int64)
    i = 0
    while i < len(anc):
        r = rng.random()
        if r < indel_rate / 2:
            i += int(rng.integers(1, max_indel + 1))  # deletion
            continue
        pos[i] = len(out)
        out.append("ACGT"[rng.integers(0, 4)] if rng.random() < sub_rate else anc[i])
        if r > 1.0 - indel_rate / 2:
            out.extend("ACGT"[k] for k in rng.integers(0, 4, size=int(rng.integers(1, max_indel + 1))))  # insertion
        i += 1
    return "".join(out), pos


def anchors_between(m1, m2, trim=14, expansion=20):
    """Anchor triples a lastz-style chain would give for two family members: maximal runs of ancestor columns present in
    both, consecutive in both and identical, trimmed by `trim` on each side (impl/pairwiseAligner.c:979-1003)."""
    (s1, p1), (s2, p2) = m1, m2
    a1 = np.frombuffer(s1.encode(), dtype=np.uint8)
    a2 = np.frombuffer(s2.encode(), dtype=np.uint8)
    both = (p1 >= 0) & (p2 >= 0)
    same = np.zeros(p1.size, dtype=bool)
    same[both] = a1[p1[both]] == a2[p2[both]]
    out, run = [], []
    for i in range(p1.size + 1):
        ok = i < p1.size and same[i]
        if ok and run and p1[i] == p1[run[-1]] + 1 and p2[i] == p2[run[-1]] + 1:
            run.append(i)
            continue
        if len(run) > 2 * trim:
            out.extend((int(p1[k]), int(p2[k]), expansion) for k in run[trim:len(run) - trim])
        run = [i] if ok else []
    return np.asarray(out, dtype=np.int64).reshape(-1, 3)
