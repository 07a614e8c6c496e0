/*
 * host/pairwiseAligner.c -- cPecan's pairwise-alignment entry points (inc/pairwiseAligner.h of the reference) as
 * plain C over the batched CUDA engine (include/cpecan_b200.h).
 *
 * What the reference does per call on the host -- split at large anchor gaps, build the band, run the banded
 * forward / backward sweeps block by block, threshold the posteriors (impl/pairwiseAligner.c:756-877, :1273-1326,
 * :1431-1513) -- happens on the device inside cpb_batch_run.  This file only translates between the reference's
 * argument conventions (NUL-terminated strings, stList of stIntTuple) and the engine's flat batch, for one pair
 * (the reference's own signatures) or many (the *Batch forms).  There is no host implementation of the DP here.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "cpecan/pairwiseAligner.h"
#include "cpecan_b200.h"
#include "host_internal.h"
#include "minijson.h"

const char *PAIRWISE_ALIGNMENT_EXCEPTION_ID = "PAIRWISE_ALIGNMENT_EXCEPTION";

void *cpecan_malloc(size_t bytes) {
    void *p = malloc(bytes ? bytes : 1);
    if (p == NULL) st_errAbort("cpecan: out of memory allocating %zu bytes", bytes);
    return p;
}

/* ------------------------------------------------------------------------------------ parameters */

PairwiseAlignmentParameters *pairwiseAlignmentBandingParameters_construct(void) {
    /* one source of truth for the defaults of impl/pairwiseAligner.c:1334-1348: the engine's cpb_params_default */
    CpbParams d;
    cpb_params_default(&d);
    PairwiseAlignmentParameters *p = cpecan_malloc(sizeof(*p));
    p->threshold = d.threshold;
    p->minDiagsBetweenTraceBack = d.minDiagsBetweenTraceBack;
    p->traceBackDiagonals = d.traceBackDiagonals;
    p->diagonalExpansion = d.diagonalExpansion;
    p->constraintDiagonalTrim = d.constraintDiagonalTrim;
    p->anchorMatrixBiggerThanThis = d.anchorMatrixBiggerThanThis;
    p->repeatMaskMatrixBiggerThanThis = d.repeatMaskMatrixBiggerThanThis;
    p->splitMatrixBiggerThanThis = d.splitMatrixBiggerThanThis;
    p->alignAmbiguityCharacters = d.alignAmbiguityCharacters != 0;
    p->gapGamma = d.gapGamma;
    p->dynamicAnchorExpansion = d.dynamicAnchorExpansion != 0;
    return p;
}

void pairwiseAlignmentBandingParameters_destruct(PairwiseAlignmentParameters *p) { free(p); }

static int int_member(MiniJson *j, int64_t *out) {
    double v;
    if (minijson_number(j, &v) != 0) return -1;
    *out = (int64_t) v;
    return 0;
}

static int params_member(MiniJson *j, const char *key, void *extra) {
    PairwiseAlignmentParameters *p = extra;
    int b;
    double v;
    if (strcmp(key, "threshold") == 0) return minijson_number(j, &p->threshold);
    if (strcmp(key, "minDiagsBetweenTraceBack") == 0) return int_member(j, &p->minDiagsBetweenTraceBack);
    if (strcmp(key, "traceBackDiagonals") == 0) return int_member(j, &p->traceBackDiagonals);
    if (strcmp(key, "diagonalExpansion") == 0) return int_member(j, &p->diagonalExpansion);
    if (strcmp(key, "constraintDiagonalTrim") == 0) return int_member(j, &p->constraintDiagonalTrim);
    if (strcmp(key, "anchorMatrixBiggerThanThis") == 0) return int_member(j, &p->anchorMatrixBiggerThanThis);
    if (strcmp(key, "repeatMaskMatrixBiggerThanThis") == 0) return int_member(j, &p->repeatMaskMatrixBiggerThanThis);
    if (strcmp(key, "splitMatrixBiggerThanThis") == 0) return int_member(j, &p->splitMatrixBiggerThanThis);
    if (strcmp(key, "alignAmbiguityCharacters") == 0) {
        if (minijson_bool(j, &b) != 0) return -1;
        p->alignAmbiguityCharacters = b != 0;
        return 0;
    }
    if (strcmp(key, "gapGamma") == 0) {
        if (minijson_number(j, &v) != 0) return -1;
        p->gapGamma = (float) v;
        return 0;
    }
    if (strcmp(key, "dynamicAnchorExpansion") == 0) {
        if (minijson_bool(j, &b) != 0) return -1;
        p->dynamicAnchorExpansion = b != 0;
        return 0;
    }
    st_errAbort("ERROR: Unrecognised key in pairwise alignment parameters json: %s\n", key);
    return -1;
}

/* impl/pairwiseAligner.c:1354-1409: defaults for every key that is absent, abort on an unknown key */
PairwiseAlignmentParameters *pairwiseAlignmentParameters_jsonParse(char *buf, size_t r) {
    PairwiseAlignmentParameters *p = pairwiseAlignmentBandingParameters_construct();
    MiniJson j;
    minijson_init(&j, buf, r);
    if (minijson_object(&j, params_member, p) != 0) st_errAbort("ERROR: could not parse pairwise alignment parameters json: %s\n", j.error);
    return p;
}

static void to_engine_params(const PairwiseAlignmentParameters *p, CpbParams *q) {
    memset(q, 0, sizeof(*q));
    q->threshold = p->threshold;
    q->minDiagsBetweenTraceBack = p->minDiagsBetweenTraceBack;
    q->traceBackDiagonals = p->traceBackDiagonals;
    q->diagonalExpansion = p->diagonalExpansion;
    q->constraintDiagonalTrim = p->constraintDiagonalTrim;
    q->anchorMatrixBiggerThanThis = p->anchorMatrixBiggerThanThis;
    q->repeatMaskMatrixBiggerThanThis = p->repeatMaskMatrixBiggerThanThis;
    q->splitMatrixBiggerThanThis = p->splitMatrixBiggerThanThis;
    q->alignAmbiguityCharacters = p->alignAmbiguityCharacters;
    q->gapGamma = p->gapGamma;
    q->dynamicAnchorExpansion = p->dynamicAnchorExpansion;
}

/* --------------------------------------------------------------------------------- device context */

static cpb_context *g_ctx = NULL;
static int g_device = -1;
static CpecanAnchorProvider g_anchorProvider = NULL;
static void *g_anchorExtra = NULL;

void cpecan_setDevice(int device) {
    if (g_ctx != NULL && device != g_device) st_errAbort("cpecan_setDevice: the device context already exists on device %d", g_device);
    g_device = device;
}

void cpecan_shutdown(void) {
    if (g_ctx != NULL) cpb_context_destroy(g_ctx);
    g_ctx = NULL;
}

void cpecan_setAnchorProvider(CpecanAnchorProvider provider, void *extra) {
    g_anchorProvider = provider;
    g_anchorExtra = extra;
}

static cpb_context *context(void) {
    if (g_ctx == NULL) {
        if (g_device < 0) {
            const char *e = getenv("CPECAN_DEVICE");
            g_device = e != NULL ? atoi(e) : 0;
        }
        if (cpb_context_create(g_device, NULL, &g_ctx) != CPB_OK)
            st_errAbort("cpecan: cannot use CUDA device %d: %s (this library has no CPU implementation of the pair-HMM)", g_device, cpb_last_error());
    }
    return g_ctx;
}

/* ----------------------------------------------------------------------------- batch marshalling */

typedef struct {
    int64_t n;
    char *seqX, *seqY;
    int64_t *xOff, *yOff, *aOff, *anchors;
    uint8_t *rl, *rr;
} Packed;

static void packed_free(Packed *k) {
    free(k->seqX);
    free(k->seqY);
    free(k->xOff);
    free(k->yOff);
    free(k->aOff);
    free(k->anchors);
    free(k->rl);
    free(k->rr);
}

static void pack(Packed *k, int64_t n, const char *const *sX, const char *const *sY, stList *const *anchorPairs, const bool *raggedLeft,
                 const bool *raggedRight, int64_t defaultExpansion) {
    memset(k, 0, sizeof(*k));
    k->n = n;
    k->xOff = cpecan_malloc((size_t) (n + 1) * sizeof(int64_t));
    k->yOff = cpecan_malloc((size_t) (n + 1) * sizeof(int64_t));
    k->aOff = cpecan_malloc((size_t) (n + 1) * sizeof(int64_t));
    k->rl = cpecan_malloc((size_t) n);
    k->rr = cpecan_malloc((size_t) n);
    k->xOff[0] = k->yOff[0] = k->aOff[0] = 0;
    for (int64_t i = 0; i < n; i++) {
        k->xOff[i + 1] = k->xOff[i] + (int64_t) strlen(sX[i]);
        k->yOff[i + 1] = k->yOff[i] + (int64_t) strlen(sY[i]);
        k->aOff[i + 1] = k->aOff[i] + (anchorPairs != NULL && anchorPairs[i] != NULL ? stList_length(anchorPairs[i]) : 0);
        k->rl[i] = raggedLeft != NULL && raggedLeft[i];
        k->rr[i] = raggedRight != NULL && raggedRight[i];
    }
    k->seqX = cpecan_malloc((size_t) k->xOff[n] + 1);
    k->seqY = cpecan_malloc((size_t) k->yOff[n] + 1);
    k->anchors = cpecan_malloc((size_t) (3 * k->aOff[n] + 3) * sizeof(int64_t));
    for (int64_t i = 0; i < n; i++) {
        memcpy(k->seqX + k->xOff[i], sX[i], (size_t) (k->xOff[i + 1] - k->xOff[i]));
        memcpy(k->seqY + k->yOff[i], sY[i], (size_t) (k->yOff[i + 1] - k->yOff[i]));
        const int64_t nA = k->aOff[i + 1] - k->aOff[i];
        for (int64_t a = 0; a < nA; a++) {
            stIntTuple *t = stList_get(anchorPairs[i], a);
            int64_t *out = k->anchors + 3 * (k->aOff[i] + a);
            out[0] = stIntTuple_get(t, 0);
            out[1] = stIntTuple_get(t, 1);
            /* anchors are (x, y, expansion); two-element tuples (as the reference's tests build them) use p->diagonalExpansion */
            out[2] = stIntTuple_length(t) > 2 ? stIntTuple_get(t, 2) : defaultExpansion;
        }
    }
}

/* runs one engine pass over the packed problems; aborts with the engine's message on failure (the reference has no error codes either) */
static cpb_batch *run(Packed *k, StateMachine *sM, PairwiseAlignmentParameters *p, int mode) {
    cpb_context *ctx = context();
    cpb_batch *b = NULL;
    if (cpb_batch_create(ctx, k->n, k->seqX, k->xOff, k->seqY, k->yOff, k->anchors, k->aOff, k->rl, k->rr, &b) != CPB_OK)
        st_errAbort("cpecan: %s", cpb_last_error());
    CpbParams q;
    to_engine_params(p, &q);
    const int rc = cpb_batch_run(b, cpecan_model_of(sM), &q, mode);
    if (rc == CPB_ERR_BAND) st_errAbort("%s: %s", PAIRWISE_ALIGNMENT_EXCEPTION_ID, cpb_last_error());
    if (rc != CPB_OK) st_errAbort("cpecan: %s", cpb_last_error());
    return b;
}

/* list `which` of the last run as n stLists of (pInt, x, y) tuples */
static stList **fetch_lists(cpb_batch *b, int64_t n, int which) {
    const int64_t total = cpb_batch_result_count(b, which);
    int64_t *off = cpecan_malloc((size_t) (n + 1) * sizeof(int64_t));
    int32_t *tri = cpecan_malloc((size_t) (3 * total + 3) * sizeof(int32_t));
    /* same list order as the reference's own lists (its callers may depend on it, e.g. the MEA walk-back) */
    if (cpb_batch_fetch_pairs_reference_order(b, which, off, tri) != CPB_OK) st_errAbort("cpecan: %s", cpb_last_error());
    stList **lists = cpecan_malloc((size_t) (n > 0 ? n : 1) * sizeof(stList *));
    for (int64_t i = 0; i < n; i++) {
        lists[i] = stList_construct3(0, (void (*)(void *)) stIntTuple_destruct);
        for (int64_t t = off[i]; t < off[i + 1]; t++) stList_append(lists[i], stIntTuple_construct3(tri[3 * t], tri[3 * t + 1], tri[3 * t + 2]));
    }
    free(off);
    free(tri);
    return lists;
}

/* ------------------------------------------------------------------------------- batched entries */

stList **getAlignedPairsUsingAnchorsBatch(StateMachine *sM, int64_t n, const char *const *sX, const char *const *sY,
                                          stList *const *anchorPairs, PairwiseAlignmentParameters *p, const bool *raggedLeft,
                                          const bool *raggedRight) {
    Packed k;
    pack(&k, n, sX, sY, anchorPairs, raggedLeft, raggedRight, p->diagonalExpansion);
    cpb_batch *b = run(&k, sM, p, CPB_MODE_ALIGNED_PAIRS);
    stList **lists = fetch_lists(b, n, 0);
    cpb_batch_destroy(b);
    packed_free(&k);
    return lists;
}

stList **getReweightedAlignedPairsUsingAnchorsBatch(StateMachine *sM, int64_t n, const char *const *sX, const char *const *sY,
                                                    stList *const *anchorPairs, PairwiseAlignmentParameters *p, const bool *raggedLeft,
                                                    const bool *raggedRight, double gapGamma) {
    Packed k;
    pack(&k, n, sX, sY, anchorPairs, raggedLeft, raggedRight, p->diagonalExpansion);
    cpb_batch *b = run(&k, sM, p, CPB_MODE_ALIGNED_PAIRS);
    if (cpb_batch_reweight_pairs(b, gapGamma) != CPB_OK) st_errAbort("cpecan: %s", cpb_last_error());
    stList **lists = fetch_lists(b, n, 0);
    cpb_batch_destroy(b);
    packed_free(&k);
    return lists;
}

void getAlignedPairsWithIndelsUsingAnchorsBatch(StateMachine *sM, int64_t n, const char *const *sX, const char *const *sY,
                                                stList *const *anchorPairs, PairwiseAlignmentParameters *p, stList ***alignedPairs,
                                                stList ***gapXPairs, stList ***gapYPairs, const bool *raggedLeft, const bool *raggedRight) {
    Packed k;
    pack(&k, n, sX, sY, anchorPairs, raggedLeft, raggedRight, p->diagonalExpansion);
    cpb_batch *b = run(&k, sM, p, CPB_MODE_ALIGNED_PAIRS_INDELS);
    *alignedPairs = fetch_lists(b, n, 0);
    *gapXPairs = fetch_lists(b, n, 1);
    *gapYPairs = fetch_lists(b, n, 2);
    cpb_batch_destroy(b);
    packed_free(&k);
}

void getExpectationsUsingAnchorsBatch(StateMachine *sM, Hmm *hmmExpectations, int64_t n, const char *const *sX, const char *const *sY,
                                      stList *const *anchorPairs, PairwiseAlignmentParameters *p, const bool *raggedLeft,
                                      const bool *raggedRight) {
    if (hmmExpectations->stateNumber != sM->stateNumber)
        st_errAbort("getExpectations: the Hmm has %lld states, the state machine %lld", (long long) hmmExpectations->stateNumber,
                    (long long) sM->stateNumber);
    Packed k;
    pack(&k, n, sX, sY, anchorPairs, raggedLeft, raggedRight, p->diagonalExpansion);
    cpb_batch *b = run(&k, sM, p, CPB_MODE_EXPECTATIONS);
    const int64_t S = sM->stateNumber;
    double total[CPB_HMM_LEN(5)];
    if (cpb_batch_fetch_expectations(b, NULL, total) != CPB_OK) st_errAbort("cpecan: %s", cpb_last_error());
    for (int64_t i = 0; i < S * S; i++) hmmExpectations->transitions[i] += total[i];
    for (int64_t i = 0; i < S * 16; i++) hmmExpectations->emissions[i] += total[S * S + i];
    hmmExpectations->likelihood += total[S * S + S * 16];
    cpb_batch_destroy(b);
    packed_free(&k);
}

/* ---- resident batches: the inputs go to the device once, every EM iteration is one more pass with a new model ---- */

struct _cpecanResidentBatch {
    cpb_batch *b;
    int64_t n;
};

CpecanResidentBatch *cpecanResidentBatch_construct(int64_t n, const char *const *sX, const char *const *sY, stList *const *anchorPairs,
                                                   PairwiseAlignmentParameters *p, const bool *raggedLeft, const bool *raggedRight) {
    Packed k;
    pack(&k, n, sX, sY, anchorPairs, raggedLeft, raggedRight, p->diagonalExpansion);
    CpecanResidentBatch *r = cpecan_malloc(sizeof(*r));
    r->n = n;
    r->b = NULL;
    if (cpb_batch_create(context(), k.n, k.seqX, k.xOff, k.seqY, k.yOff, k.anchors, k.aOff, k.rl, k.rr, &r->b) != CPB_OK)
        st_errAbort("cpecan: %s", cpb_last_error());
    packed_free(&k);
    return r;
}

void cpecanResidentBatch_destruct(CpecanResidentBatch *r) {
    if (r == NULL) return;
    cpb_batch_destroy(r->b);
    free(r);
}

void cpecanResidentBatch_getExpectations(CpecanResidentBatch *r, StateMachine *sM, Hmm *hmmExpectations, PairwiseAlignmentParameters *p) {
    if (hmmExpectations->stateNumber != sM->stateNumber)
        st_errAbort("getExpectations: the Hmm has %lld states, the state machine %lld", (long long) hmmExpectations->stateNumber,
                    (long long) sM->stateNumber);
    CpbParams q;
    to_engine_params(p, &q);
    const int rc = cpb_batch_run(r->b, cpecan_model_of(sM), &q, CPB_MODE_EXPECTATIONS);
    if (rc == CPB_ERR_BAND) st_errAbort("%s: %s", PAIRWISE_ALIGNMENT_EXCEPTION_ID, cpb_last_error());
    if (rc != CPB_OK) st_errAbort("cpecan: %s", cpb_last_error());
    const int64_t S = sM->stateNumber;
    double total[CPB_HMM_LEN(5)];
    if (cpb_batch_fetch_expectations(r->b, NULL, total) != CPB_OK) st_errAbort("cpecan: %s", cpb_last_error());
    for (int64_t i = 0; i < S * S; i++) hmmExpectations->transitions[i] += total[i];
    for (int64_t i = 0; i < S * 16; i++) hmmExpectations->emissions[i] += total[S * S + i];
    hmmExpectations->likelihood += total[S * S + S * 16];
}

void computeForwardProbabilityBatch(StateMachine *sM, int64_t n, const char *const *sX, const char *const *sY, stList *const *anchorPairs,
                                    PairwiseAlignmentParameters *p, const bool *raggedLeft, const bool *raggedRight, double *logProbs) {
    Packed k;
    pack(&k, n, sX, sY, anchorPairs, raggedLeft, raggedRight, p->diagonalExpansion);
    cpb_batch *b = run(&k, sM, p, CPB_MODE_FORWARD);
    if (cpb_batch_fetch_forward(b, logProbs) != CPB_OK) st_errAbort("cpecan: %s", cpb_last_error());
    cpb_batch_destroy(b);
    packed_free(&k);
}

/* ---------------------------------------------------------------- the reference's one-pair forms */

stList *getAlignedPairsUsingAnchors(StateMachine *sM, const char *sX, const char *sY, stList *anchorPairs, PairwiseAlignmentParameters *p,
                                    bool alignmentHasRaggedLeftEnd, bool alignmentHasRaggedRightEnd) {
    stList **lists = getAlignedPairsUsingAnchorsBatch(sM, 1, &sX, &sY, &anchorPairs, p, &alignmentHasRaggedLeftEnd, &alignmentHasRaggedRightEnd);
    stList *out = lists[0];
    free(lists);
    return out;
}

void getAlignedPairsWithIndelsUsingAnchors(StateMachine *sM, const char *sX, const char *sY, stList *anchorPairs,
                                           PairwiseAlignmentParameters *p, stList **alignedPairs, stList **gapXPairs, stList **gapYPairs,
                                           bool alignmentHasRaggedLeftEnd, bool alignmentHasRaggedRightEnd) {
    stList **m, **gx, **gy;
    getAlignedPairsWithIndelsUsingAnchorsBatch(sM, 1, &sX, &sY, &anchorPairs, p, &m, &gx, &gy, &alignmentHasRaggedLeftEnd,
                                               &alignmentHasRaggedRightEnd);
    *alignedPairs = m[0];
    *gapXPairs = gx[0];
    *gapYPairs = gy[0];
    free(m);
    free(gx);
    free(gy);
}

void getExpectationsUsingAnchors(StateMachine *sM, Hmm *hmmExpectations, const char *sX, const char *sY, stList *anchorPairs,
                                 PairwiseAlignmentParameters *p, bool alignmentHasRaggedLeftEnd, bool alignmentHasRaggedRightEnd) {
    getExpectationsUsingAnchorsBatch(sM, hmmExpectations, 1, &sX, &sY, &anchorPairs, p, &alignmentHasRaggedLeftEnd, &alignmentHasRaggedRightEnd);
}

double computeForwardProbability(char *seqX, char *seqY, stList *anchorPairs, PairwiseAlignmentParameters *p, StateMachine *sM,
                                 bool alignmentHasRaggedLeftEnd, bool alignmentHasRaggedRightEnd) {
    const char *sX = seqX, *sY = seqY;
    double v = 0.0;
    computeForwardProbabilityBatch(sM, 1, &sX, &sY, &anchorPairs, p, &alignmentHasRaggedLeftEnd, &alignmentHasRaggedRightEnd, &v);
    return v;
}

/* anchors for the forms that take none (getBlastPairsForPairwiseAlignmentParameters, impl/pairwiseAligner.c:1162-1196) */
static stList *anchors_for(const char *sX, const char *sY, PairwiseAlignmentParameters *p) {
    const int64_t lX = (int64_t) strlen(sX), lY = (int64_t) strlen(sY);
    if (lX * lY <= p->anchorMatrixBiggerThanThis) return stList_construct3(0, (void (*)(void *)) stIntTuple_destruct);
    if (g_anchorProvider == NULL)
        st_errAbort("cpecan: a %lld x %lld matrix is bigger than anchorMatrixBiggerThanThis (%lld) and needs anchors; the reference gets them "
                    "from a LASTZ subprocess, which is outside this library: pass anchors to the ...UsingAnchors form or register a provider "
                    "with cpecan_setAnchorProvider",
                    (long long) lX, (long long) lY, (long long) p->anchorMatrixBiggerThanThis);
    return g_anchorProvider(sX, sY, lX, lY, p, g_anchorExtra);
}

stList *getAlignedPairs(StateMachine *sM, const char *string1, const char *string2, PairwiseAlignmentParameters *p,
                        bool alignmentHasRaggedLeftEnd, bool alignmentHasRaggedRightEnd) {
    stList *anchorPairs = anchors_for(string1, string2, p);
    stList *out = getAlignedPairsUsingAnchors(sM, string1, string2, anchorPairs, p, alignmentHasRaggedLeftEnd, alignmentHasRaggedRightEnd);
    stList_destruct(anchorPairs);
    return out;
}

void getAlignedPairsWithIndels(StateMachine *sM, const char *string1, const char *string2, PairwiseAlignmentParameters *p,
                               stList **alignedPairs, stList **gapXPairs, stList **gapYPairs, bool alignmentHasRaggedLeftEnd,
                               bool alignmentHasRaggedRightEnd) {
    stList *anchorPairs = anchors_for(string1, string2, p);
    getAlignedPairsWithIndelsUsingAnchors(sM, string1, string2, anchorPairs, p, alignedPairs, gapXPairs, gapYPairs, alignmentHasRaggedLeftEnd,
                                          alignmentHasRaggedRightEnd);
    stList_destruct(anchorPairs);
}

void getExpectations(StateMachine *sM, Hmm *hmmExpectations, const char *sX, const char *sY, PairwiseAlignmentParameters *p,
                     bool alignmentHasRaggedLeftEnd, bool alignmentHasRaggedRightEnd) {
    stList *anchorPairs = anchors_for(sX, sY, p);
    getExpectationsUsingAnchors(sM, hmmExpectations, sX, sY, anchorPairs, p, alignmentHasRaggedLeftEnd, alignmentHasRaggedRightEnd);
    stList_destruct(anchorPairs);
}

/* ------------------------------------------------------------------------ test-visible helpers */

Diagonal diagonal_construct(int64_t xay, int64_t xmyL, int64_t xmyR) {
    /* impl/pairwiseAligner.c:30-43: both ends share the parity of xay and xmyL <= xmyR */
    if ((xay + xmyL) % 2 != 0 || (xay + xmyR) % 2 != 0 || xmyL > xmyR)
        st_errAbort("%s: Attempt to create diagonal with invalid coordinates: xay %lld xmyL %lld xmyR %lld", PAIRWISE_ALIGNMENT_EXCEPTION_ID,
                    (long long) xay, (long long) xmyL, (long long) xmyR);
    Diagonal d = { xay, xmyL, xmyR };
    return d;
}
int64_t diagonal_getXay(Diagonal d) { return d.xay; }
int64_t diagonal_getMinXmy(Diagonal d) { return d.xmyL; }
int64_t diagonal_getMaxXmy(Diagonal d) { return d.xmyR; }
int64_t diagonal_getWidth(Diagonal d) { return (d.xmyR - d.xmyL) / 2 + 1; }
int64_t diagonal_getXCoordinate(int64_t xay, int64_t xmy) { return (xay + xmy) / 2; }
int64_t diagonal_getYCoordinate(int64_t xay, int64_t xmy) { return (xay - xmy) / 2; }
int64_t diagonal_equals(Diagonal d1, Diagonal d2) { return d1.xay == d2.xay && d1.xmyL == d2.xmyL && d1.xmyR == d2.xmyR; }

struct _band {
    Diagonal *diagonals;
    int64_t lXalY; /* lX + lY: diagonals 0 .. lXalY */
};
struct _bandIterator {
    Band *band;
    int64_t index;
};

static Band *band_from_device(stList *anchorPairs, int64_t lX, int64_t lY, int64_t expansion, int dynamic) {
    const int64_t nA = anchorPairs != NULL ? stList_length(anchorPairs) : 0;
    int64_t *an = cpecan_malloc((size_t) (3 * nA + 3) * sizeof(int64_t));
    for (int64_t i = 0; i < nA; i++) {
        stIntTuple *t = stList_get(anchorPairs, i);
        an[3 * i] = stIntTuple_get(t, 0);
        an[3 * i + 1] = stIntTuple_get(t, 1);
        an[3 * i + 2] = stIntTuple_length(t) > 2 ? stIntTuple_get(t, 2) : expansion;
    }
    int64_t *out3 = cpecan_malloc((size_t) (3 * (lX + lY + 1)) * sizeof(int64_t));
    const int rc = cpb_band(context(), an, nA, lX, lY, expansion, dynamic, out3);
    if (rc == CPB_ERR_BAND) st_errAbort("%s: %s", PAIRWISE_ALIGNMENT_EXCEPTION_ID, cpb_last_error());
    if (rc != CPB_OK) st_errAbort("cpecan: %s", cpb_last_error());
    Band *band = cpecan_malloc(sizeof(*band));
    band->lXalY = lX + lY;
    band->diagonals = cpecan_malloc((size_t) (lX + lY + 1) * sizeof(Diagonal));
    for (int64_t d = 0; d <= lX + lY; d++) {
        band->diagonals[d].xay = out3[3 * d];
        band->diagonals[d].xmyL = out3[3 * d + 1];
        band->diagonals[d].xmyR = out3[3 * d + 2];
    }
    free(an);
    free(out3);
    return band;
}

Band *band_construct(stList *anchorPairs, int64_t lX, int64_t lY, int64_t expansion) { return band_from_device(anchorPairs, lX, lY, expansion, 0); }
Band *band_constructDynamic(stList *anchorPairs, int64_t lX, int64_t lY) { return band_from_device(anchorPairs, lX, lY, 0, 1); }

void band_destruct(Band *band) {
    if (band == NULL) return;
    free(band->diagonals);
    free(band);
}

BandIterator *bandIterator_construct(Band *band) {
    BandIterator *it = cpecan_malloc(sizeof(*it));
    it->band = band;
    it->index = 0;
    return it;
}
BandIterator *bandIterator_clone(BandIterator *bandIterator) {
    BandIterator *it = cpecan_malloc(sizeof(*it));
    *it = *bandIterator;
    return it;
}
void bandIterator_destruct(BandIterator *bandIterator) { free(bandIterator); }

/* impl/pairwiseAligner.c:263-277: the index runs over [0, lX+lY+1]; reads clamp to the first / last diagonal */
Diagonal bandIterator_getNext(BandIterator *it) {
    const int64_t last = it->band->lXalY;
    Diagonal d = it->band->diagonals[it->index > last ? last : it->index];
    if (it->index <= last) it->index++;
    return d;
}
Diagonal bandIterator_getPrevious(BandIterator *it) {
    if (it->index > 0) it->index--;
    return it->band->diagonals[it->index];
}

Symbol symbol_convertCharToSymbol(char i) {
    switch (i) {
    case 'A':
    case 'a':
        return a;
    case 'C':
    case 'c':
        return c;
    case 'G':
    case 'g':
        return g;
    case 'T':
    case 't':
        return t;
    default:
        return n;
    }
}

char symbol_convertSymbolToChar(Symbol i) {
    static const char letters[] = "ACGTN";
    return letters[(int) i >= 0 && (int) i < 4 ? (int) i : 4];
}

Symbol *symbol_convertStringToSymbols(const char *s, int64_t sL) {
    Symbol *out = cpecan_malloc((size_t) (sL > 0 ? sL : 1) * sizeof(Symbol));
    for (int64_t i = 0; i < sL; i++) out[i] = symbol_convertCharToSymbol(s[i]);
    return out;
}

stList *getSplitPoints(stList *anchorPairs, int64_t lX, int64_t lY, int64_t maxMatrixSize, bool alignmentHasRaggedLeftEnd,
                       bool alignmentHasRaggedRightEnd) {
    const int64_t nA = anchorPairs != NULL ? stList_length(anchorPairs) : 0;
    int64_t *an = cpecan_malloc((size_t) (3 * nA + 3) * sizeof(int64_t));
    for (int64_t i = 0; i < nA; i++) {
        stIntTuple *t = stList_get(anchorPairs, i);
        an[3 * i] = stIntTuple_get(t, 0);
        an[3 * i + 1] = stIntTuple_get(t, 1);
        an[3 * i + 2] = 0;
    }
    int64_t cap = nA + 2;
    int64_t *out4 = cpecan_malloc((size_t) (4 * cap) * sizeof(int64_t));
    const int64_t nReg = cpb_split_points(an, nA, lX, lY, maxMatrixSize, alignmentHasRaggedLeftEnd, alignmentHasRaggedRightEnd, out4, cap);
    if (nReg < 0 || nReg > cap) st_errAbort("cpecan: getSplitPoints failed: %s", cpb_last_error());
    stList *out = stList_construct3(0, (void (*)(void *)) stIntTuple_destruct);
    for (int64_t r = 0; r < nReg; r++) stList_append(out, stIntTuple_construct4(out4[4 * r], out4[4 * r + 1], out4[4 * r + 2], out4[4 * r + 3]));
    free(an);
    free(out4);
    return out;
}
