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
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "cpecan/pairwiseAligner.h"
#include "host_internal.h"

typedef struct {
    int32_t x, y, len; /* an exact match of `len` bases starting at (x, y) */
} Seed;

/* a, c, g, t -> 0..3, anything else -1; lower case counts as "anything else" where the soft mask is honoured.  A table, not a switch:
 * on sequence data a switch is four unpredictable branches per base. */
static const signed char kCode[2][256] = {
    { ['A'] = 1, ['C'] = 2, ['G'] = 3, ['T'] = 4, ['a'] = 1, ['c'] = 2, ['g'] = 3, ['t'] = 4 },
    { ['A'] = 1, ['C'] = 2, ['G'] = 3, ['T'] = 4 },
};
static inline int base_code(char ch, int maskLower) { return kCode[maskLower != 0][(unsigned char) ch] - 1; }

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

/* exact k-mer matches (x, y) with the k-mer unique in Y, in ascending x; consecutive matches of one diagonal are merged.
 * The index of Y is an open-addressing table of 16-byte entries (the k-mer and the position + 1 of its first occurrence, negated once
 * it occurs again) -- one cache line per probe -- and both passes work through the sequence 32 positions at a time: roll the
 * window over the block, ask for the 32 table lines, then probe them (the table of a 100 kb sequence is 4 MB: every probe is a
 * cache miss, and the next one cannot be issued before the window has moved on unless the addresses are computed ahead). */
typedef struct {
    uint64_t key;
    int32_t slot;
    int32_t pad_;
} IndexEntry;
enum { SEED_BLOCK = 32 };

/* rolls the k-mer window over positions [i0, i1) of s; for every position that ends a whole word: its word, its home slot (also
 * prefetched) and its index in the block.  Returns how many there are. */
static inline int roll_block(const char *s, int64_t i0, int64_t i1, int k, int maskLower, uint64_t wordMask, uint64_t capMask, const IndexEntry *table,
                             uint64_t *w, int *valid, uint64_t *words, uint64_t *homes, int *at) {
    int m = 0;
    for (int64_t i = i0; i < i1; i++) {
        const int c = base_code(s[i], maskLower);
        if (c < 0) {
            *valid = 0;
            continue;
        }
        *w = ((*w << 2) | (uint64_t) c) & wordMask;
        if (++*valid < k) continue;
        words[m] = *w;
        homes[m] = mix(*w) & capMask;
        at[m] = (int) (i - i0);
        __builtin_prefetch(table + homes[m]);
        m++;
    }
    return m;
}

static Seed *find_seeds(const char *sX, const char *sY, int64_t lX, int64_t lY, int k, int maskLower, int64_t *nOut) {
    *nOut = 0;
    if (lX < k || lY < k) return NULL;
    int64_t cap = 16;
    while (cap < 2 * lY) cap <<= 1;
    IndexEntry *table = calloc((size_t) cap, sizeof(IndexEntry));
    if (table == NULL) st_errAbort("cpecan: out of memory indexing %lld bases", (long long) lY);
    const uint64_t wordMask = k < 32 ? (((uint64_t) 1 << (2 * k)) - 1) : ~(uint64_t) 0, capMask = (uint64_t) (cap - 1);
    uint64_t words[SEED_BLOCK], homes[SEED_BLOCK];
    int at[SEED_BLOCK];
    uint64_t w = 0;
    int valid = 0;
    for (int64_t i0 = 0; i0 < lY; i0 += SEED_BLOCK) {
        const int64_t i1 = i0 + SEED_BLOCK < lY ? i0 + SEED_BLOCK : lY;
        const int m = roll_block(sY, i0, i1, k, maskLower, wordMask, capMask, table, &w, &valid, words, homes, at);
        for (int j = 0; j < m; j++) {
            uint64_t h = homes[j];
            for (;;) {
                if (table[h].slot == 0) {
                    table[h].slot = (int32_t) (i0 + at[j] - k + 2);
                    table[h].key = words[j];
                    break;
                }
                if (table[h].key == words[j]) {
                    if (table[h].slot > 0) table[h].slot = -table[h].slot;
                    break;
                }
                h = (h + 1) & capMask;
            }
        }
    }
    int64_t n = 0, room = 1024;
    Seed *seeds = cpecan_malloc((size_t) room * sizeof(Seed));
    w = 0;
    valid = 0;
    for (int64_t i0 = 0; i0 < lX; i0 += SEED_BLOCK) {
        const int64_t i1 = i0 + SEED_BLOCK < lX ? i0 + SEED_BLOCK : lX;
        const int m = roll_block(sX, i0, i1, k, maskLower, wordMask, capMask, table, &w, &valid, words, homes, at);
        for (int j = 0; j < m; j++) {
            uint64_t h = homes[j];
            while (table[h].slot != 0 && table[h].key != words[j]) h = (h + 1) & capMask;
            if (table[h].slot <= 0) continue;
            const int32_t x = (int32_t) (i0 + at[j] - k + 1), y = table[h].slot - 1;
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
    }
    free(table);
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

/* ---- anchors as flat (x, y, expansion) int32 triples: the three functions below hand stLists to their callers, but between themselves
 * they work on arrays (a tuple per anchor, three times over, was half the time of anchoring a 100 kb pair) ---- */
typedef struct {
    int32_t *v; /* 3 per anchor */
    int64_t n, cap;
} Triples;

static void triples_push(Triples *t, int64_t x, int64_t y, int64_t e) {
    if (t->n == t->cap) {
        t->cap = t->cap ? 2 * t->cap : 1024;
        t->v = realloc(t->v, (size_t) t->cap * 3 * sizeof(int32_t));
        if (t->v == NULL) st_errAbort("cpecan: out of memory");
    }
    t->v[3 * t->n] = (int32_t) x;
    t->v[3 * t->n + 1] = (int32_t) y;
    t->v[3 * t->n + 2] = (int32_t) e;
    t->n++;
}

static int by_x_plus_y(const void *a, const void *b) {
    const int32_t *p = a, *q = b;
    const int64_t k = (int64_t) p[0] + p[1], l = (int64_t) q[0] + q[1];
    return k > l ? 1 : (k < l ? -1 : 0);
}
static int by_elements(const void *a, const void *b) { /* stIntTuple_cmpFn on tuples of three */
    const int32_t *p = a, *q = b;
    for (int i = 0; i < 3; i++) {
        if (p[i] != q[i]) return p[i] < q[i] ? -1 : 1;
    }
    return 0;
}
/* sorts unless the triples are in order already (a chain's anchors are: x and y both ascend) */
static void triples_sort(Triples *t, int (*cmp)(const void *, const void *)) {
    int64_t i = 1;
    while (i < t->n && cmp(t->v + 3 * (i - 1), t->v + 3 * i) <= 0) i++;
    if (i < t->n) qsort(t->v, (size_t) t->n, 3 * sizeof(int32_t), cmp);
}

static void blast_pairs(const char *sX, const char *sY, int64_t lX, int64_t lY, int64_t trim, int64_t diagonalExpansion, bool repeatMask, Triples *pairs) {
    if (lX == 0 || lY == 0) return;
    if (lX > 0x7FFFFFF0 || lY > 0x7FFFFFF0) st_errAbort("getBlastPairs: sequences of %lld and %lld bases are too long", (long long) lX, (long long) lY);
    if (diagonalExpansion < 0 || diagonalExpansion > 0x7FFFFFF0) st_errAbort("getBlastPairs: diagonal expansion %lld out of range", (long long) diagonalExpansion);
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
            for (int64_t l = trim; l < x1 - x0 - trim; l++) triples_push(pairs, x0 + l, y0 + l, diagonalExpansion);
        }
        i = j;
    }
    free(chain);
    triples_sort(pairs, by_x_plus_y); /* as the reference does (:1063); the chain is in this order already */
}

stList *getBlastPairs(const char *sX, const char *sY, int64_t lX, int64_t lY, int64_t trim, int64_t diagonalExpansion, bool repeatMask) {
    Triples pairs = { NULL, 0, 0 };
    blast_pairs(sX, sY, lX, lY, trim, diagonalExpansion, repeatMask, &pairs);
    stList *list = cpecan_tripleList_construct(pairs.v, pairs.n);
    free(pairs.v);
    return list;
}

/* impl/pairwiseAligner.c:1095-1135 on triples sorted by (x, y, expansion): keep those that no later one undercuts in x or y (backward
 * sweep) and that exceed every earlier one in both (forward sweep).  The reference keeps the survivors of the backward sweep in a
 * sorted SET of tuples: a triple equal (in all elements) to a surviving one counts as surviving too. */
static void filter_overlap(const int32_t *in, int64_t n, int64_t dx, int64_t dy, Triples *out) {
    char *alive = cpecan_malloc((size_t) n + 1);
    int64_t pX = INT64_MAX, pY = INT64_MAX;
    for (int64_t i = n - 1; i >= 0; i--) {
        const int64_t x = in[3 * i], y = in[3 * i + 1];
        alive[i] = x < pX && y < pY;
        pX = x < pX ? x : pX;
        pY = y < pY ? y : pY;
    }
    pX = INT64_MIN;
    pY = INT64_MIN;
    for (int64_t i = 0; i < n; i++) {
        const int64_t x = in[3 * i], y = in[3 * i + 1];
        int survives = alive[i];
        for (int64_t j = i + 1; !survives && j < n && by_elements(in + 3 * i, in + 3 * j) == 0; j++) survives = alive[j];
        for (int64_t j = i - 1; !survives && j >= 0 && by_elements(in + 3 * i, in + 3 * j) == 0; j--) survives = alive[j];
        if (x > pX && y > pY && survives) triples_push(out, x + dx, y + dy, in[3 * i + 2]);
        pX = x > pX ? x : pX;
        pY = y > pY ? y : pY;
    }
    free(alive);
}

stList *filterToRemoveOverlap(stList *sortedOverlappingPairs) {
    const int64_t n = stList_length(sortedOverlappingPairs);
    Triples in = { NULL, 0, 0 }, out = { NULL, 0, 0 };
    for (int64_t i = 0; i < n; i++) {
        stIntTuple *pair = stList_get(sortedOverlappingPairs, i);
        const int64_t x = stIntTuple_get(pair, 0), y = stIntTuple_get(pair, 1), e = stIntTuple_get(pair, 2);
        if (x < INT32_MIN || x > INT32_MAX || y < INT32_MIN || y > INT32_MAX || e < INT32_MIN || e > INT32_MAX) {
            st_errAbort("filterToRemoveOverlap: (%lld, %lld, %lld) is outside the 32-bit range", (long long) x, (long long) y, (long long) e);
        }
        triples_push(&in, x, y, e);
    }
    filter_overlap(in.v, in.n, 0, 0, &out);
    stList *list = cpecan_tripleList_construct(out.v, out.n);
    free(in.v);
    free(out.v);
    return list;
}

/* anchors of the sub-matrix [pX, x) x [pY, y) if it is still bigger than anchorMatrixBiggerThanThis (:1137-1160) */
static void anchor_gap(const char *sX, const char *sY, int64_t pX, int64_t pY, int64_t x, int64_t y, PairwiseAlignmentParameters *p, Triples *combined) {
    const int64_t lX2 = x - pX, lY2 = y - pY;
    if (lX2 <= 0 || lY2 <= 0) return;
    const int64_t matrixSize = lX2 * lY2;
    if (matrixSize <= p->anchorMatrixBiggerThanThis) return;
    Triples unfiltered = { NULL, 0, 0 };
    blast_pairs(sX + pX, sY + pY, lX2, lY2, p->constraintDiagonalTrim, p->diagonalExpansion, matrixSize > p->repeatMaskMatrixBiggerThanThis, &unfiltered);
    triples_sort(&unfiltered, by_elements);
    filter_overlap(unfiltered.v, unfiltered.n, pX, pY, combined);
    free(unfiltered.v);
}

stList *getBlastPairsForPairwiseAlignmentParameters(const char *sX, const char *sY, const int64_t lX, const int64_t lY, PairwiseAlignmentParameters *p) {
    if (lX * lY <= p->anchorMatrixBiggerThanThis) return stList_construct3(0, (void (*)(void *)) stIntTuple_destruct);
    Triples unfiltered = { NULL, 0, 0 }, top = { NULL, 0, 0 }, combined = { NULL, 0, 0 };
    blast_pairs(sX, sY, lX, lY, p->constraintDiagonalTrim, p->diagonalExpansion, 1, &unfiltered);
    triples_sort(&unfiltered, by_elements);
    filter_overlap(unfiltered.v, unfiltered.n, 0, 0, &top);
    free(unfiltered.v);
    int64_t pX = 0, pY = 0;
    for (int64_t i = 0; i < top.n; i++) {
        const int64_t x = top.v[3 * i], y = top.v[3 * i + 1];
        anchor_gap(sX, sY, pX, pY, x, y, p, &combined);
        triples_push(&combined, x, y, top.v[3 * i + 2]);
        pX = x + 1;
        pY = y + 1;
    }
    anchor_gap(sX, sY, pX, pY, lX, lY, p, &combined);
    free(top.v);
    stList *list = cpecan_tripleList_construct(combined.v, combined.n);
    free(combined.v);
    return list;
}
