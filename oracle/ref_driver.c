/*
 * oracle/ref_driver.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Flat-C adaptor (oracle_api.h) over the *reference's own* public functions, linked with the
 * reference's unmodified impl/pairwiseAligner.c + impl/stateMachine.c compiled where they lie
 * under /root/reference (oracle/Makefile, target ref).  Output: oracle/_ref/libcpecan_ref.so.
 * Used to (i) pin the restatement in pairhmm_oracle.c, (ii) generate tests/golden fixtures,
 * (iii) serve as the "reference" CPU baseline in bench.py.
 */
#include <pthread.h>

#include "sonLib.h"
#include "pairwiseAligner.h"
#include "oracle_api.h"

const char *orc_identity(void) { return "reference"; }

static PairwiseAlignmentParameters *toParams(const OrcParams *o) {
    PairwiseAlignmentParameters *p = pairwiseAlignmentBandingParameters_construct();
    p->threshold = o->threshold;
    p->minDiagsBetweenTraceBack = o->minDiagsBetweenTraceBack;
    p->traceBackDiagonals = o->traceBackDiagonals;
    p->diagonalExpansion = o->diagonalExpansion;
    p->constraintDiagonalTrim = o->constraintDiagonalTrim;
    p->splitMatrixBiggerThanThis = o->splitMatrixBiggerThanThis;
    p->dynamicAnchorExpansion = o->dynamicAnchorExpansion != 0;
    return p;
}

void orc_default_params(OrcParams *o) {
    PairwiseAlignmentParameters *p = pairwiseAlignmentBandingParameters_construct();
    o->threshold = p->threshold;
    o->minDiagsBetweenTraceBack = p->minDiagsBetweenTraceBack;
    o->traceBackDiagonals = p->traceBackDiagonals;
    o->diagonalExpansion = p->diagonalExpansion;
    o->constraintDiagonalTrim = p->constraintDiagonalTrim;
    o->splitMatrixBiggerThanThis = p->splitMatrixBiggerThanThis;
    o->dynamicAnchorExpansion = p->dynamicAnchorExpansion;
    pairwiseAlignmentBandingParameters_destruct(p);
}

static StateMachine *toStateMachine(const OrcModel *m) {
    StateMachineType type = (StateMachineType) m->type;
    if (!m->fromHmm) {
        return (type == fiveState || type == fiveStateAsymmetric) ? stateMachine5_construct(type) : stateMachine3_construct(type);
    }
    Hmm *hmm = hmm_constructEmpty(0.0, type);
    int64_t S = hmm->stateNumber;
    memcpy(hmm->transitions, m->transitions, sizeof(double) * S * S);
    memcpy(hmm->emissions, m->emissions, sizeof(double) * S * 16);
    StateMachine *sM = hmm_getStateMachine(hmm);
    hmm_destruct(hmm);
    return sM;
}

static stList *toAnchors(const int64_t *a, int64_t n) {
    stList *l = stList_construct3(0, (void (*)(void *)) stIntTuple_destruct);
    for (int64_t i = 0; i < n; i++) {
        stList_append(l, stIntTuple_construct3(a[3 * i], a[3 * i + 1], a[3 * i + 2]));
    }
    return l;
}

static int64_t drain(stList *l, int64_t *out, int64_t cap) {
    int64_t n = stList_length(l);
    for (int64_t i = 0; i < n && i < cap; i++) {
        stIntTuple *t = stList_get(l, i);
        out[3 * i] = stIntTuple_get(t, 0);
        out[3 * i + 1] = stIntTuple_get(t, 1);
        out[3 * i + 2] = stIntTuple_get(t, 2);
    }
    stList_destruct(l);
    return n;
}

double orc_logadd(double x, double y) { return logAdd(x, y); }

int64_t orc_band(const int64_t *anchors, int64_t nAnchors, int64_t lX, int64_t lY, int64_t expansion, int dynamic,
                 int64_t *out) {
    extern Band *band_constructDynamic(stList *anchorPairs, int64_t lX, int64_t lY); /* not in the header */
    stList *a = toAnchors(anchors, nAnchors);
    Band *band = dynamic ? band_constructDynamic(a, lX, lY) : band_construct(a, lX, lY, expansion);
    BandIterator *it = bandIterator_construct(band);
    for (int64_t i = 0; i <= lX + lY; i++) {
        Diagonal d = bandIterator_getNext(it);
        out[3 * i] = diagonal_getXay(d);
        out[3 * i + 1] = diagonal_getMinXmy(d);
        out[3 * i + 2] = diagonal_getMaxXmy(d);
    }
    bandIterator_destruct(it);
    band_destruct(band);
    stList_destruct(a);
    return lX + lY + 1;
}

int64_t orc_split_points(const int64_t *anchors, int64_t nAnchors, int64_t lX, int64_t lY, int64_t split, int raggedLeft,
                         int raggedRight, int64_t *out, int64_t cap) {
    stList *a = toAnchors(anchors, nAnchors);
    stList *s = getSplitPoints(a, lX, lY, split, raggedLeft, raggedRight);
    int64_t n = stList_length(s);
    for (int64_t i = 0; i < n && i < cap; i++) {
        for (int64_t j = 0; j < 4; j++) out[4 * i + j] = stIntTuple_get(stList_get(s, i), j);
    }
    stList_destruct(s);
    stList_destruct(a);
    return n;
}

/* --- model dump through the reference's own cellCalculate vtable --- */
typedef struct {
    double *out;
    int64_t n, cap, group;
} DumpCtx;
static void dumpTransition(double *fromCells, double *toCells, int64_t from, int64_t to, double eP, double tP, void *extra) {
    (void) toCells;
    DumpCtx *c = extra;
    /* group is identified by which neighbour array was passed: we tag arrays with their group id in [0] */
    double rec[5] = { fromCells[0], (double) from, (double) to, eP, tP };
    for (int i = 0; i < 5; i++) {
        if (c->n < c->cap) c->out[c->n] = rec[i];
        c->n++;
    }
}
int64_t orc_model_dump(const OrcModel *m, double *out, int64_t cap) {
    StateMachine *sM = toStateMachine(m);
    int64_t S = sM->stateNumber;
    DumpCtx c = { out, 0, cap, 0 };
    double (*vecs[4])(StateMachine *, int64_t) = { sM->startStateProb, sM->raggedStartStateProb, sM->endStateProb,
            sM->raggedEndStateProb };
    for (int v = 0; v < 4; v++) {
        for (int64_t s = 0; s < S; s++) {
            if (c.n < c.cap) out[c.n] = vecs[v](sM, s);
            c.n++;
        }
    }
    double cur[5], lower[5] = { 0 }, middle[5] = { 1 }, upper[5] = { 2 };
    lower[0] = 0;
    middle[0] = 1;
    upper[0] = 2;
    for (int cX = 0; cX < 5; cX++) {
        for (int cY = 0; cY < 5; cY++) {
            sM->cellCalculate(sM, cur, lower, middle, upper, (Symbol) cX, (Symbol) cY, dumpTransition, &c);
        }
    }
    stateMachine_destruct(sM);
    return c.n;
}

int64_t orc_aligned_pairs(const OrcModel *m, const OrcParams *o, const char *sX, const char *sY, const int64_t *anchors,
                          int64_t nAnchors, int raggedLeft, int raggedRight, int64_t *out, int64_t cap) {
    StateMachine *sM = toStateMachine(m);
    PairwiseAlignmentParameters *p = toParams(o);
    stList *a = toAnchors(anchors, nAnchors);
    stList *pairs = getAlignedPairsUsingAnchors(sM, sX, sY, a, p, raggedLeft, raggedRight);
    int64_t n = drain(pairs, out, cap);
    stList_destruct(a);
    pairwiseAlignmentBandingParameters_destruct(p);
    stateMachine_destruct(sM);
    return n;
}

void orc_aligned_pairs_with_indels(const OrcModel *m, const OrcParams *o, const char *sX, const char *sY,
                                   const int64_t *anchors, int64_t nAnchors, int raggedLeft, int raggedRight,
                                   int64_t *outMatch, int64_t *outGapX, int64_t *outGapY, int64_t cap, int64_t *counts) {
    StateMachine *sM = toStateMachine(m);
    PairwiseAlignmentParameters *p = toParams(o);
    stList *a = toAnchors(anchors, nAnchors);
    stList *match = NULL, *gapX = NULL, *gapY = NULL;
    getAlignedPairsWithIndelsUsingAnchors(sM, sX, sY, a, p, &match, &gapX, &gapY, raggedLeft, raggedRight);
    counts[0] = drain(match, outMatch, cap);
    counts[1] = drain(gapX, outGapX, cap);
    counts[2] = drain(gapY, outGapY, cap);
    stList_destruct(a);
    pairwiseAlignmentBandingParameters_destruct(p);
    stateMachine_destruct(sM);
}

static void expectationsInto(StateMachine *sM, PairwiseAlignmentParameters *p, const char *sX, const char *sY, stList *a,
                             int raggedLeft, int raggedRight, double *hmmOut) {
    Hmm *hmm = hmm_constructEmpty(0.0, sM->type);
    int64_t S = hmm->stateNumber;
    getExpectationsUsingAnchors(sM, hmm, sX, sY, a, p, raggedLeft, raggedRight);
    for (int64_t i = 0; i < S * S; i++) hmmOut[i] += hmm->transitions[i];
    for (int64_t i = 0; i < S * 16; i++) hmmOut[S * S + i] += hmm->emissions[i];
    hmmOut[S * S + S * 16] += hmm->likelihood;
    hmm_destruct(hmm);
}

void orc_expectations(const OrcModel *m, const OrcParams *o, const char *sX, const char *sY, const int64_t *anchors,
                      int64_t nAnchors, int raggedLeft, int raggedRight, double *hmmOut) {
    StateMachine *sM = toStateMachine(m);
    PairwiseAlignmentParameters *p = toParams(o);
    stList *a = toAnchors(anchors, nAnchors);
    expectationsInto(sM, p, sX, sY, a, raggedLeft, raggedRight, hmmOut);
    stList_destruct(a);
    pairwiseAlignmentBandingParameters_destruct(p);
    stateMachine_destruct(sM);
}

double orc_forward_prob(const OrcModel *m, const OrcParams *o, const char *sX, const char *sY, const int64_t *anchors,
                        int64_t nAnchors, int raggedLeft, int raggedRight) {
    StateMachine *sM = toStateMachine(m);
    PairwiseAlignmentParameters *p = toParams(o);
    stList *a = toAnchors(anchors, nAnchors);
    double v = computeForwardProbability((char *) sX, (char *) sY, a, p, sM, raggedLeft, raggedRight);
    stList_destruct(a);
    pairwiseAlignmentBandingParameters_destruct(p);
    stateMachine_destruct(sM);
    return v;
}

/* --- batch over pthreads (CPU baseline) --- */
typedef struct {
    const OrcModel *m;
    const OrcParams *o;
    int64_t nPairs;
    const char *seqX, *seqY;
    const int64_t *xOff, *yOff, *anchors, *aOff;
    const uint8_t *rl, *rr;
    int mode, tid, nThreads;
    int64_t *counts, *checksum;
    double hmm[ORC_HMM_LEN(5)];
} BatchArg;

static void *batchWorker(void *v) {
    BatchArg *b = v;
    StateMachine *sM = toStateMachine(b->m);
    PairwiseAlignmentParameters *p = toParams(b->o);
    for (int64_t i = b->tid; i < b->nPairs; i += b->nThreads) {
        char *sX = stString_getSubString(b->seqX, b->xOff[i], b->xOff[i + 1] - b->xOff[i]);
        char *sY = stString_getSubString(b->seqY, b->yOff[i], b->yOff[i + 1] - b->yOff[i]);
        stList *a = toAnchors(b->anchors + 3 * b->aOff[i], b->aOff[i + 1] - b->aOff[i]);
        int rl = b->rl ? b->rl[i] : 0, rr = b->rr ? b->rr[i] : 0;
        if (b->mode == 0) {
            stList *pairs = getAlignedPairsUsingAnchors(sM, sX, sY, a, p, rl, rr);
            int64_t cs = 0;
            for (int64_t j = 0; j < stList_length(pairs); j++) {
                stIntTuple *t = stList_get(pairs, j);
                cs += stIntTuple_get(t, 0) * (stIntTuple_get(t, 1) + 1) + stIntTuple_get(t, 2);
            }
            if (b->counts) b->counts[i] = stList_length(pairs);
            if (b->checksum) b->checksum[i] = cs;
            stList_destruct(pairs);
        } else {
            expectationsInto(sM, p, sX, sY, a, rl, rr, b->hmm);
        }
        stList_destruct(a);
        free(sX);
        free(sY);
    }
    pairwiseAlignmentBandingParameters_destruct(p);
    stateMachine_destruct(sM);
    return NULL;
}

void orc_batch(const OrcModel *m, const OrcParams *o, int64_t nPairs, const char *seqX, const int64_t *xOff, const char *seqY,
               const int64_t *yOff, const int64_t *anchors, const int64_t *aOff, const uint8_t *raggedLeft,
               const uint8_t *raggedRight, int mode, int nThreads, int64_t *counts, int64_t *checksum, double *hmmOut) {
    if (nThreads < 1) nThreads = 1;
    BatchArg *args = st_calloc(nThreads, sizeof(BatchArg));
    pthread_t *th = st_calloc(nThreads, sizeof(pthread_t));
    for (int t = 0; t < nThreads; t++) {
        BatchArg b = { m, o, nPairs, seqX, seqY, xOff, yOff, anchors, aOff, raggedLeft, raggedRight, mode, t, nThreads, counts,
                checksum, { 0 } };
        args[t] = b;
        pthread_create(&th[t], NULL, batchWorker, &args[t]);
    }
    int64_t S = (m->type == 0 || m->type == 1) ? 5 : 3;
    for (int t = 0; t < nThreads; t++) {
        pthread_join(th[t], NULL);
        if (mode == 1 && hmmOut) {
            for (int64_t i = 0; i < ORC_HMM_LEN(S); i++) hmmOut[i] += args[t].hmm[i];
        }
    }
    free(args);
    free(th);
}

/* ---- the reference's list post-processing (impl/pairwiseAligner.c:979-1003, :1519-1792), for tests of cpecan_b200/csrc/host/realign.c ---- */

int64_t orc_cigar_anchors(const int64_t *ops, int64_t nOps, int64_t start1, int64_t start2, int64_t trim, int64_t expansion, int64_t *out,
        int64_t cap) {
    /* ops: (type, length) with type 0 = match, 1 = indel X, 2 = indel Y */
    struct List opList;
    struct AlignmentOperation *store = st_malloc(sizeof(struct AlignmentOperation) * (nOps + 1));
    void **ptrs = st_malloc(sizeof(void *) * (nOps + 1));
    struct PairwiseAlignment pA;
    memset(&pA, 0, sizeof(pA));
    pA.start1 = pA.end1 = start1;
    pA.start2 = pA.end2 = start2;
    pA.strand1 = pA.strand2 = 1;
    for (int64_t i = 0; i < nOps; i++) {
        store[i].opType = ops[2 * i] == 0 ? PAIRWISE_MATCH : (ops[2 * i] == 1 ? PAIRWISE_INDEL_X : PAIRWISE_INDEL_Y);
        store[i].length = ops[2 * i + 1];
        store[i].score = 0;
        ptrs[i] = &store[i];
        if (ops[2 * i] != 2) pA.end1 += ops[2 * i + 1];
        if (ops[2 * i] != 1) pA.end2 += ops[2 * i + 1];
    }
    opList.list = ptrs;
    opList.length = nOps;
    pA.operationList = &opList;
    int64_t n = drain(convertPairwiseForwardStrandAlignmentToAnchorPairs(&pA, trim, expansion), out, cap);
    free(store);
    free(ptrs);
    return n;
}

int64_t orc_reweight(const int64_t *pairs, int64_t n, int64_t lX, int64_t lY, double gapGamma, int64_t *out) {
    return drain(reweightAlignedPairs2(toAnchors(pairs, n), lX, lY, gapGamma), out, n);
}

void orc_scores(const char *sX, const char *sY, const int64_t *pairs, int64_t n, double *out4) {
    stList *l = toAnchors(pairs, n);
    out4[0] = scoreByIdentity((char *) sX, (char *) sY, strlen(sX), strlen(sY), l);
    out4[1] = scoreByIdentityIgnoringGaps((char *) sX, (char *) sY, l);
    out4[2] = scoreByPosteriorProbability(strlen(sX), strlen(sY), l);
    out4[3] = scoreByPosteriorProbabilityIgnoringGaps(l);
    stList_destruct(l);
}

int64_t orc_mea(const int64_t *pairs, int64_t n, const int64_t *gapX, int64_t nX, const int64_t *gapY, int64_t nY, int64_t lX, int64_t lY,
        double gapGamma, int64_t *out, double *score) {
    PairwiseAlignmentParameters *p = pairwiseAlignmentBandingParameters_construct();
    p->gapGamma = gapGamma;
    stList *a = toAnchors(pairs, n), *x = toAnchors(gapX, nX), *y = toAnchors(gapY, nY);
    int64_t k = drain(getMaximalExpectedAccuracyPairwiseAlignment(a, x, y, lX, lY, score, p), out, n);
    stList_destruct(a);
    stList_destruct(x);
    stList_destruct(y);
    pairwiseAlignmentBandingParameters_destruct(p);
    return k;
}

int64_t orc_left_shift(const int64_t *pairs, int64_t n, const char *sX, const char *sY, int64_t *out, int64_t cap) {
    stList *a = toAnchors(pairs, n);
    int64_t k = drain(leftShiftAlignment(a, (char *) sX, (char *) sY), out, cap);
    stList_destruct(a);
    return k;
}
