/*
 * host/banding.c -- getPosteriorProbsWithBanding and getPosteriorProbsWithBandingSplittingAlignmentsByLargeGaps
 * (inc/pairwiseAligner.h:245, :264; impl/pairwiseAligner.c:756-877, :1273-1326) over the batched engine.
 *
 * The reference passes a per-diagonal callback into its banded forward-backward; here the three callbacks of its own wrappers
 * are the three device modes, recognised by address (include/cpecan/pairwiseAligner.h).  All split regions of a call go through
 * ONE device pass as independent problems; afterwards the results are handed over region by region exactly as the reference's
 * loop does: the region's tuples are appended to the caller's lists in the reference's emission order, then
 * coordinateCorrectionFn(x1, y1, extraArgs) runs.
 */
#include <stdlib.h>
#include <string.h>

#include "cpecan/pairwiseAligner.h"
#include "cpecan_b200.h"
#include "host_internal.h"

static void marker_called(const char *name) {
    st_errAbort("cpecan: %s is a marker for getPosteriorProbsWithBanding (the per-diagonal work is compiled into the CUDA kernels and no "
                "DpMatrix exists on the host); it cannot be called directly",
                name);
}
#define MARKER(name)                                                                                                                      \
    void name(StateMachine *sM, int64_t xay, DpMatrix *forwardDpMatrix, DpMatrix *backwardDpMatrix, const SymbolString sX,               \
              const SymbolString sY, double totalProbability, PairwiseAlignmentParameters *p, void *extraArgs) {                         \
        (void) sM, (void) xay, (void) forwardDpMatrix, (void) backwardDpMatrix, (void) sX, (void) sY, (void) totalProbability, (void) p, \
            (void) extraArgs;                                                                                                             \
        marker_called(#name);                                                                                                             \
    }
MARKER(diagonalCalculationPosteriorMatchProbs)
MARKER(diagonalCalculationPosteriorProbs)
MARKER(diagonalCalculationExpectations)

SymbolString symbolString_construct(const char *sequence, int64_t length) {
    SymbolString s;
    s.sequence = symbol_convertStringToSymbols(sequence, length);
    s.length = length;
    return s;
}

void symbolString_destruct(SymbolString s) { free(s.sequence); }

static int mode_of(DiagonalPosteriorProbFn fn) {
    if (fn == diagonalCalculationPosteriorMatchProbs) return CPB_MODE_ALIGNED_PAIRS;
    if (fn == diagonalCalculationPosteriorProbs) return CPB_MODE_ALIGNED_PAIRS_INDELS;
    if (fn == diagonalCalculationExpectations) return CPB_MODE_EXPECTATIONS;
    st_errAbort("cpecan: getPosteriorProbsWithBanding runs on the device, where only the reference's own per-diagonal callbacks exist: pass "
                "diagonalCalculationPosteriorMatchProbs, diagonalCalculationPosteriorProbs or diagonalCalculationExpectations");
    return -1;
}

/* regions [x1, x2) x [y1, y2) of one pair of strings as independent problems, anchors rebased (impl/pairwiseAligner.c:1282-1309) */
typedef struct {
    int64_t n;
    int64_t *x1, *y1;
    char **sX, **sY;
    stList **anchors;
    bool *rl, *rr;
} Regions;

static void regions_free(Regions *g) {
    for (int64_t i = 0; i < g->n; i++) {
        free(g->sX[i]);
        free(g->sY[i]);
        stList_destruct(g->anchors[i]);
    }
    free(g->x1);
    free(g->y1);
    free(g->sX);
    free(g->sY);
    free(g->anchors);
    free(g->rl);
    free(g->rr);
}

static char *substring(const char *s, int64_t from, int64_t length) {
    char *out = cpecan_malloc((size_t) length + 1);
    memcpy(out, s + from, (size_t) length);
    out[length] = '\0';
    return out;
}

/* one device pass over the regions, then the reference's per-region hand-over */
static void run_regions(StateMachine *sM, Regions *g, PairwiseAlignmentParameters *p, DiagonalPosteriorProbFn fn, void (*coordinateCorrectionFn)(),
                        void *extraArgs) {
    const int mode = mode_of(fn);
    if (g->n == 0) return;
    PairwiseAlignmentParameters q = *p;
    q.splitMatrixBiggerThanThis = INT64_MAX; /* the regions are already split */
    if (mode == CPB_MODE_EXPECTATIONS) {
        getExpectationsUsingAnchorsBatch(sM, (Hmm *) extraArgs, g->n, (const char *const *) g->sX, (const char *const *) g->sY, g->anchors, &q, g->rl,
                                         g->rr);
        for (int64_t i = 0; i < g->n && coordinateCorrectionFn != NULL; i++) coordinateCorrectionFn(g->x1[i], g->y1[i], extraArgs);
        return;
    }
    stList **lists[3] = { NULL, NULL, NULL };
    const int nLists = mode == CPB_MODE_ALIGNED_PAIRS ? 1 : 3;
    if (nLists == 1)
        lists[0] = getAlignedPairsUsingAnchorsBatch(sM, g->n, (const char *const *) g->sX, (const char *const *) g->sY, g->anchors, &q, g->rl, g->rr);
    else
        getAlignedPairsWithIndelsUsingAnchorsBatch(sM, g->n, (const char *const *) g->sX, (const char *const *) g->sY, g->anchors, &q, &lists[0],
                                                   &lists[1], &lists[2], g->rl, g->rr);
    for (int64_t i = 0; i < g->n; i++) {
        for (int l = 0; l < nLists; l++) {
            /* the batch entry points return a one-region list as the reference's wrappers do, i.e. popped off the callback's list
             * (impl/pairwiseAligner.c:1411-1418): the callback's own order is the reverse */
            stList *out = ((void **) extraArgs)[2 * l];
            stList *got = lists[l][i];
            while (stList_length(got) > 0) stList_append(out, stList_pop(got));
            stList_destruct(got);
        }
        if (coordinateCorrectionFn != NULL) coordinateCorrectionFn(g->x1[i], g->y1[i], extraArgs);
    }
    for (int l = 0; l < nLists; l++) free(lists[l]);
}

void getPosteriorProbsWithBanding(StateMachine *sM, stList *anchorPairs, const SymbolString sX, const SymbolString sY,
                                  PairwiseAlignmentParameters *p, bool alignmentHasRaggedLeftEnd, bool alignmentHasRaggedRightEnd,
                                  DiagonalPosteriorProbFn diagonalPosteriorProbFn, void *extraArgs) {
    Regions g;
    memset(&g, 0, sizeof(g));
    g.n = 1;
    g.x1 = cpecan_malloc(sizeof(int64_t));
    g.y1 = cpecan_malloc(sizeof(int64_t));
    g.sX = cpecan_malloc(sizeof(char *));
    g.sY = cpecan_malloc(sizeof(char *));
    g.anchors = cpecan_malloc(sizeof(stList *));
    g.rl = cpecan_malloc(sizeof(bool));
    g.rr = cpecan_malloc(sizeof(bool));
    g.x1[0] = g.y1[0] = 0;
    /* the engine takes characters: symbols a, c, g, t, n go back to their letters (the engine maps them to the same symbols) */
    g.sX[0] = cpecan_malloc((size_t) sX.length + 1);
    g.sY[0] = cpecan_malloc((size_t) sY.length + 1);
    for (int64_t i = 0; i < sX.length; i++) g.sX[0][i] = symbol_convertSymbolToChar(sX.sequence[i]);
    for (int64_t i = 0; i < sY.length; i++) g.sY[0][i] = symbol_convertSymbolToChar(sY.sequence[i]);
    g.sX[0][sX.length] = g.sY[0][sY.length] = '\0';
    g.anchors[0] = stList_construct();
    if (anchorPairs != NULL) stList_appendAll(g.anchors[0], anchorPairs); /* borrowed: the new list has no destructor */
    g.rl[0] = alignmentHasRaggedLeftEnd;
    g.rr[0] = alignmentHasRaggedRightEnd;
    run_regions(sM, &g, p, diagonalPosteriorProbFn, NULL, extraArgs);
    regions_free(&g);
}

void getPosteriorProbsWithBandingSplittingAlignmentsByLargeGaps(StateMachine *sM, stList *anchorPairs, const char *sX, const char *sY, int64_t lX,
                                                                int64_t lY, PairwiseAlignmentParameters *p, bool alignmentHasRaggedLeftEnd,
                                                                bool alignmentHasRaggedRightEnd, DiagonalPosteriorProbFn diagonalPosteriorProbFn,
                                                                void (*coordinateCorrectionFn)(), void *extraArgs) {
    stList *splitPoints = getSplitPoints(anchorPairs, lX, lY, p->splitMatrixBiggerThanThis, alignmentHasRaggedLeftEnd, alignmentHasRaggedRightEnd);
    Regions g;
    memset(&g, 0, sizeof(g));
    g.n = stList_length(splitPoints);
    const size_t cap = (size_t) (g.n > 0 ? g.n : 1);
    g.x1 = cpecan_malloc(cap * sizeof(int64_t));
    g.y1 = cpecan_malloc(cap * sizeof(int64_t));
    g.sX = cpecan_malloc(cap * sizeof(char *));
    g.sY = cpecan_malloc(cap * sizeof(char *));
    g.anchors = cpecan_malloc(cap * sizeof(stList *));
    g.rl = cpecan_malloc(cap * sizeof(bool));
    g.rr = cpecan_malloc(cap * sizeof(bool));
    const int64_t nA = anchorPairs != NULL ? stList_length(anchorPairs) : 0;
    int64_t j = 0;
    for (int64_t i = 0; i < g.n; i++) {
        stIntTuple *region = stList_get(splitPoints, i);
        const int64_t x1 = stIntTuple_get(region, 0), y1 = stIntTuple_get(region, 1), x2 = stIntTuple_get(region, 2), y2 = stIntTuple_get(region, 3);
        g.x1[i] = x1;
        g.y1[i] = y1;
        g.sX[i] = substring(sX, x1, x2 - x1);
        g.sY[i] = substring(sY, y1, y2 - y1);
        g.anchors[i] = stList_construct3(0, (void (*)(void *)) stIntTuple_destruct);
        /* the anchors in front of the region's far corner, rebased; the third element (expansion) is kept when there is one */
        for (; j < nA; j++) {
            stIntTuple *anchor = stList_get(anchorPairs, j);
            const int64_t x = stIntTuple_get(anchor, 0), y = stIntTuple_get(anchor, 1);
            if (x + y >= x2 + y2) break;
            if (x < x1 || x >= x2 || y < y1 || y >= y2)
                st_errAbort("%s: anchor (%lld, %lld) lies outside its split region", PAIRWISE_ALIGNMENT_EXCEPTION_ID, (long long) x, (long long) y);
            stList_append(g.anchors[i], stIntTuple_length(anchor) > 2 ? stIntTuple_construct3(x - x1, y - y1, stIntTuple_get(anchor, 2))
                                                                       : stIntTuple_construct2(x - x1, y - y1));
        }
        g.rl[i] = alignmentHasRaggedLeftEnd || i > 0;
        g.rr[i] = alignmentHasRaggedRightEnd || i < g.n - 1;
    }
    run_regions(sM, &g, p, diagonalPosteriorProbFn, coordinateCorrectionFn, extraArgs);
    regions_free(&g);
    stList_destruct(splitPoints);
}
