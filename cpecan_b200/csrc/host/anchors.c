/*
 * host/anchors.c -- anchors without a subprocess.
 *
 * The reference finds the anchors of getAlignedPairs / getExpectations by writing both sequences to temporary files and running
 * LASTZ through popen, once per pair and once more per large gap (getBlastPairs, impl/pairwiseAligner.c:1005-1080).  Here the same
 * three functions of its header -- getBlastPairs, filterToRemoveOverlap, getBlastPairsForPairwiseAlignmentParameters
 * (inc/pairwiseAligner.h:251-257) -- run in process:
 *
 *   getBlastPairs       a seed-and-chain aligner written for the job the anchors have: exact k-mer seeds that are unique in Y, the
 *                       heaviest colinear chain of them (both coordinates strictly increasing), neighbouring seeds of one diagonal
 *                       joined into gapless runs when the stretch between them is mostly matches, and every run handed over column
 *                       by column with `trim` columns dropped at both ends -- exactly what the reference makes of a LASTZ cigar
 *                       (convertPairwiseForwardStrandAlignmentToAnchorPairs, :979-1003).  It is NOT LASTZ: the anchor sets differ, and
 *                       they may, because anchors only place the band; the posteriors inside the band come from the same DP.
 *   filterToRemoveOverlap                        the reference's filter (:1095-1135), restated
 *   getBlastPairsForPairwiseAlignmentParameters  the reference's two-level scheme (:1137-1196): anchors of the whole matrix from
 *                       soft-masked sequence, then every gap between them that is still bigger than anchorMatrixBiggerThanThis
 *                       anchored again on its own (without the soft mask unless it is bigger than repeatMaskMatrixBiggerThanThis)
 *
 * This is host code on purpose: it is a CALLER of the hot path (the reference's own is a CPU subprocess), it works on one pair at
 * a time in O(l) memory, and the batch entry points run it on a host thread per problem.
 */
#include <stdlib.h>
#include <string.h>

#include "cpecan/pairwiseAligner.h"
#include "host_internal.h"

typedef struct {
    int32_t x, y, len; /* an exact match of `len` bases starting at (x, y) */
} Seed;

static inline int base_code(char ch, int maskLower) {
    switch (ch) {
    case 'A': return 0;
    case 'C': return 1;
    case 'G': return 2;
    case 'T': return 3;
    case 'a': return maskLower ? -1 : 0;
    case 'c': return maskLower ? -1 : 1;
    case 'g': return maskLower ? -1 : 2;
    case 't': return maskLower ? -1 : 3;
    default: return -1;
    }
}

static inline int same_base(char p, char q) {
    const int u = base_code(p, 0);
    return u >= 0 && u == base_code(q, 0);
}

static inline uint64_t mix(uint64_t k) {
    k ^= k >> 33;
    k *= 0xFF51AFD7ED558CCDull;
    k ^= k >> 33;
    return k;
}

/* word length: long enough that a random k-mer of X is unlikely to have a partner anywhere in Y */
static int word_length(int64_t lX, int64_t lY) {
    int k = 11;
    double space = 4194304.0; /* 4^11 */
    while (k < 24 && space < 64.0 * (double) (lX < lY ? lY : lX)) {
        k++;
        space *= 4.0;
    }
    return k;
}

/* exact k-mer matches (x, y) with the k-mer unique in Y, in ascending x; consecutive matches of one diagonal are merged */
static Seed *find_seeds(const char *sX, const char *sY, int64_t lX, int64_t lY, int k, int maskLower, int64_t *nOut) {
    *nOut = 0;
    if (lX < k || lY < k) return NULL;
    int64_t cap = 16;
    while (cap < 2 * lY) cap <<= 1;
    /* open addressing; slot = position + 1 of the k-mer's first occurrence in Y, negated once it occurs again */
    int32_t *slot = calloc((size_t) cap, sizeof(int32_t));
    uint64_t *keys = cpecan_malloc((size_t) cap * sizeof(uint64_t));
    if (slot == NULL) st_errAbort("cpecan: out of memory indexing %lld bases", (long long) lY);
    const uint64_t wordMask = k < 32 ? (((uint64_t) 1 << (2 * k)) - 1) : ~(uint64_t) 0;
    uint64_t w = 0;
    int valid = 0;
    for (int64_t i = 0; i < lY; i++) {
        const int c = base_code(sY[i], maskLower);
        if (c < 0) {
            valid = 0;
            continue;
        }
        w = ((w << 2) | (uint64_t) c) & wordMask;
        if (++valid < k) continue;
        uint64_t h = mix(w) & (uint64_t) (cap - 1);
        for (;;) {
            if (slot[h] == 0) {
                slot[h] = (int32_t) (i - k + 2);
                keys[h] = w;
                break;
            }
            if (keys[h] == w) {
                if (slot[h] > 0) slot[h] = -slot[h];
                break;
            }
            h = (h + 1) & (uint64_t) (cap - 1);
        }
    }
    int64_t n = 0, room = 1024;
    Seed *seeds = cpecan_malloc((size_t) room * sizeof(Seed));
    w = 0;
    valid = 0;
    for (int64_t i = 0; i < lX; i++) {
        const int c = base_code(sX[i], maskLower);
        if (c < 0) {
            valid = 0;
            continue;
        }
        w = ((w << 2) | (uint64_t) c) & wordMask;
        if (++valid < k) continue;
        uint64_t h = mix(w) & (uint64_t) (cap - 1);
        while (slot[h] != 0 && keys[h] != w) h = (h + 1) & (uint64_t) (cap - 1);
        if (slot[h] <= 0) continue;
        const int32_t x = (int32_t) (i - k + 1), y = slot[h] - 1;
        if (n > 0 && seeds[n - 1].x + seeds[n - 1].len - k + 1 == x && seeds[n - 1].y + seeds[n - 1].len - k + 1 == y) {
            seeds[n - 1].len++; /* the previous match shifted by one: one longer exact run */
            continue;
        }
        if (n == room) {
            room *= 2;
            seeds = realloc(seeds, (size_t) room * sizeof(Seed));
            if (seeds == NULL) st_errAbort("cpecan: out of memory");
        }
        seeds[n].x = x;
        seeds[n].y = y;
        seeds[n].len = k;
        n++;
    }
    free(slot);
    free(keys);
    *nOut = n;
    return seeds;
}

/* The heaviest chain of seeds with x and y strictly increasing from one seed's end to the next one's start: weight = matched
 * bases.  Seeds arrive in ascending x; best[] over y is kept in a Fenwick tree of prefix maxima.  Returns the chain in order. */
static int64_t chain_seeds(Seed *seeds, int64_t n, int64_t lY, Seed **chainOut) {
    *chainOut = NULL;
    if (n == 0) return 0;
    int64_t *score = cpecan_malloc((size_t) n * sizeof(int64_t)), *prev = cpecan_malloc((size_t) n * sizeof(int64_t));
    int64_t *treeScore = calloc((size_t) lY + 2, sizeof(int64_t)), *treeWho = cpecan_malloc((size_t) (lY + 2) * sizeof(int64_t));
    if (treeScore == NULL) st_errAbort("cpecan: out of memory");
    for (int64_t i = 0; i <= lY + 1; i++) treeWho[i] = -1;
    /* a seed may only follow seeds that END before it starts in x: seeds are released into the tree in order of their end */
    int64_t *byEnd = cpecan_malloc((size_t) n * sizeof(int64_t));
    for (int64_t i = 0; i < n; i++) byEnd[i] = i;
    /* ends are nearly sorted already (starts ascend, lengths are short): insertion sort */
    for (int64_t i = 1; i < n; i++) {
        const int64_t v = byEnd[i];
        const int64_t e = (int64_t) seeds[v].x + seeds[v].len;
        int64_t j = i - 1;
        while (j >= 0 && (int64_t) seeds[byEnd[j]].x + seeds[byEnd[j]].len > e) {
            byEnd[j + 1] = byEnd[j];
            j--;
        }
        byEnd[j + 1] = v;
    }
    int64_t released = 0, bestEnd = -1, bestScore = 0;
    for (int64_t i = 0; i < n; i++) {
        while (released < n && (int64_t) seeds[byEnd[released]].x + seeds[byEnd[released]].len <= seeds[i].x) {
            const int64_t v = byEnd[released++];
            for (int64_t t = (int64_t) seeds[v].y + seeds[v].len; t <= lY + 1; t += t & -t) { /* key: the seed's end in y (1-based exclusive) */
                if (score[v] > treeScore[t]) {
                    treeScore[t] = score[v];
                    treeWho[t] = v;
                }
            }
        }
        int64_t s = 0, who = -1;
        for (int64_t t = seeds[i].y; t > 0; t -= t & -t) { /* predecessors whose end in y is <= this seed's start */
            if (treeScore[t] > s) {
                s = treeScore[t];
                who = treeWho[t];
            }
        }
        score[i] = s + seeds[i].len;
        prev[i] = who;
        if (score[i] > bestScore) {
            bestScore = score[i];
            bestEnd = i;
        }
    }
    int64_t m = 0;
    for (int64_t v = bestEnd; v >= 0; v = prev[v]) m++;
    Seed *chain = cpecan_malloc((size_t) (m > 0 ? m : 1) * sizeof(Seed));
    int64_t at = m;
    for (int64_t v = bestEnd; v >= 0; v = prev[v]) chain[--at] = seeds[v];
    free(score);
    free(prev);
    free(treeScore);
    free(treeWho);
    free(byEnd);
    *chainOut = chain;
    return m;
}

static int sort_by_x_plus_y(const void *a, const void *b) {
    const int64_t k = stIntTuple_get((stIntTuple *) a, 0) + stIntTuple_get((stIntTuple *) a, 1);
    const int64_t l = stIntTuple_get((stIntTuple *) b, 0) + stIntTuple_get((stIntTuple *) b, 1);
    return k > l ? 1 : (k < l ? -1 : 0);
}

stList *getBlastPairs(const char *sX, const char *sY, int64_t lX, int64_t lY, int64_t trim, int64_t diagonalExpansion, bool repeatMask) {
    stList *pairs = stList_construct3(0, (void (*)(void *)) stIntTuple_destruct);
    if (lX == 0 || lY == 0) return pairs;
    if (lX > 0x7FFFFFF0 || lY > 0x7FFFFFF0) st_errAbort("getBlastPairs: sequences of %lld and %lld bases are too long", (long long) lX, (long long) lY);
    const int k = word_length(lX, lY);
    int64_t n = 0;
    Seed *seeds = find_seeds(sX, sY, lX, lY, k, repeatMask, &n);
    Seed *chain = NULL;
    const int64_t m = chain_seeds(seeds, n, lY, &chain);
    free(seeds);
    /* runs: chained seeds of one diagonal are joined when at least 6 in 10 of the columns between them match (a gapless stretch
     * with substitutions, what a cigar calls one match operation); a change of diagonal ends the run */
    int64_t i = 0;
    while (i < m) {
        int64_t x0 = chain[i].x, y0 = chain[i].y, x1 = x0 + chain[i].len;
        int64_t j = i + 1;
        while (j < m && (int64_t) chain[j].x - chain[j].y == x0 - y0) {
            const int64_t gap = chain[j].x - x1;
            int64_t same = 0;
            for (int64_t t = 0; t < gap; t++) same += same_base(sX[x1 + t], sY[x1 + t - (x0 - y0)]);
            if (10 * same < 6 * gap) break;
            x1 = (int64_t) chain[j].x + chain[j].len;
            j++;
        }
        /* a lone word that joined nothing is as likely a chance match inside a gap as a piece of the alignment: it anchors nothing */
        if (j - i > 1 || x1 - x0 >= 2 * k) {
            for (int64_t l = trim; l < x1 - x0 - trim; l++) stList_append(pairs, stIntTuple_construct3(x0 + l, y0 + l, diagonalExpansion));
        }
        i = j;
    }
    free(chain);
    stList_sort(pairs, sort_by_x_plus_y); /* as the reference does (:1063); the chain is already in this order */
    return pairs;
}

/* impl/pairwiseAligner.c:1095-1135: of pairs sorted by (x, y), keep those that no later pair undercuts in x or y (backward sweep) and
 * that exceed every earlier pair in both (forward sweep) */
stList *filterToRemoveOverlap(stList *sortedOverlappingPairs) {
    const int64_t n = stList_length(sortedOverlappingPairs);
    stList *out = stList_construct3(0, (void (*)(void *)) stIntTuple_destruct);
    char *alive = cpecan_malloc((size_t) n + 1);
    int64_t pX = INT64_MAX, pY = INT64_MAX;
    for (int64_t i = n - 1; i >= 0; i--) {
        stIntTuple *pair = stList_get(sortedOverlappingPairs, i);
        const int64_t x = stIntTuple_get(pair, 0), y = stIntTuple_get(pair, 1);
        alive[i] = x < pX && y < pY;
        pX = x < pX ? x : pX;
        pY = y < pY ? y : pY;
    }
    /* the reference keeps the survivors of the backward sweep in a sorted SET of tuples: a pair equal (in all elements) to a
     * surviving one counts as surviving too */
    pX = INT64_MIN;
    pY = INT64_MIN;
    for (int64_t i = 0; i < n; i++) {
        stIntTuple *pair = stList_get(sortedOverlappingPairs, i);
        const int64_t x = stIntTuple_get(pair, 0), y = stIntTuple_get(pair, 1);
        int survives = alive[i];
        for (int64_t j = i + 1; !survives && j < n && stIntTuple_equalsFn(pair, stList_get(sortedOverlappingPairs, j)); j++) survives = alive[j];
        for (int64_t j = i - 1; !survives && j >= 0 && stIntTuple_equalsFn(pair, stList_get(sortedOverlappingPairs, j)); j--) survives = alive[j];
        if (x > pX && y > pY && survives) stList_append(out, stIntTuple_construct3(x, y, stIntTuple_get(pair, 2)));
        pX = x > pX ? x : pX;
        pY = y > pY ? y : pY;
    }
    free(alive);
    return out;
}

/* anchors of the sub-matrix [pX, x) x [pY, y) if it is still bigger than anchorMatrixBiggerThanThis (:1137-1160) */
static void anchor_gap(const char *sX, const char *sY, int64_t pX, int64_t pY, int64_t x, int64_t y, PairwiseAlignmentParameters *p, stList *combined) {
    const int64_t lX2 = x - pX, lY2 = y - pY;
    if (lX2 <= 0 || lY2 <= 0) return;
    const int64_t matrixSize = lX2 * lY2;
    if (matrixSize <= p->anchorMatrixBiggerThanThis) return;
    stList *unfiltered = getBlastPairs(sX + pX, sY + pY, lX2, lY2, p->constraintDiagonalTrim, p->diagonalExpansion, matrixSize > p->repeatMaskMatrixBiggerThanThis);
    stList_sort(unfiltered, stIntTuple_cmpFn);
    stList *bottom = filterToRemoveOverlap(unfiltered);
    stList_destruct(unfiltered);
    for (int64_t i = 0; i < stList_length(bottom); i++) {
        stIntTuple *t = stList_get(bottom, i);
        stList_append(combined, stIntTuple_construct3(stIntTuple_get(t, 0) + pX, stIntTuple_get(t, 1) + pY, stIntTuple_get(t, 2)));
    }
    stList_destruct(bottom);
}

stList *getBlastPairsForPairwiseAlignmentParameters(const char *sX, const char *sY, const int64_t lX, const int64_t lY, PairwiseAlignmentParameters *p) {
    if (lX * lY <= p->anchorMatrixBiggerThanThis) return stList_construct3(0, (void (*)(void *)) stIntTuple_destruct);
    stList *unfiltered = getBlastPairs(sX, sY, lX, lY, p->constraintDiagonalTrim, p->diagonalExpansion, 1);
    stList_sort(unfiltered, stIntTuple_cmpFn);
    stList *top = filterToRemoveOverlap(unfiltered);
    stList_destruct(unfiltered);
    stList *combined = stList_construct3(0, (void (*)(void *)) stIntTuple_destruct);
    int64_t pX = 0, pY = 0;
    for (int64_t i = 0; i < stList_length(top); i++) {
        stIntTuple *anchor = stList_get(top, i);
        const int64_t x = stIntTuple_get(anchor, 0), y = stIntTuple_get(anchor, 1);
        anchor_gap(sX, sY, pX, pY, x, y, p, combined);
        stList_append(combined, stIntTuple_construct3(x, y, stIntTuple_get(anchor, 2)));
        pX = x + 1;
        pY = y + 1;
    }
    anchor_gap(sX, sY, pX, pY, lX, lY, p, combined);
    stList_destruct(top);
    return combined;
}
