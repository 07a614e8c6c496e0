/*
 * test_cpecan_api.c -- exercises libcpecan.so through the reference's own C API (include/cpecan/*.h).
 *
 *   test_cpecan_api host              containers, Hmm file formats, parameter JSON, split points (no GPU needed)
 *   test_cpecan_api gpu               the reference's known-answer tests that reach the device:
 *                                     test_bands (tests/pairwiseAlignerTest.c:69-132), test_diagonalDPCalculations's four
 *                                     pairs (:311-322), test_getAlignedPairs / ...WithRaggedEnds properties (:649-715),
 *                                     test_computeForwardProbability ordering (:1157-1188), test_em likelihood (:1091-1155)
 *   test_cpecan_api run IN OUT        runs the problems in IN through the batched and the one-pair entry points and
 *                                     writes the results to OUT (tests/test_gpu_host_api.py compares them with the oracle)
 */
#include <inttypes.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "cpecan/pairwiseAligner.h"

static int g_failures = 0, g_checks = 0;
#define CHECK(cond)                                                                  \
    do {                                                                             \
        g_checks++;                                                                  \
        if (!(cond)) {                                                               \
            g_failures++;                                                            \
            fprintf(stderr, "FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond);          \
        }                                                                            \
    } while (0)

static uint64_t g_rng = 0x9E3779B97F4A7C15ull;
static uint64_t rnd(void) {
    uint64_t z = (g_rng += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static int64_t rnd_int(int64_t lo, int64_t hi) { return lo + (int64_t) (rnd() % (uint64_t) (hi - lo)); }

static char *random_sequence(int64_t length) {
    char *s = malloc((size_t) length + 1);
    for (int64_t i = 0; i < length; i++) s[i] = "ACGT"[rnd() & 3];
    s[length] = '\0';
    return s;
}

/* substitutions and short indels, in the spirit of evolveSequence (impl/randomSequences.c:50-73) */
static char *evolve(const char *s) {
    const size_t len = strlen(s);
    char *out = malloc(2 * len + 16);
    size_t k = 0;
    for (size_t i = 0; i < len; i++) {
        const uint64_t r = rnd() % 100;
        if (r < 2) continue;                                  /* deletion */
        if (r < 4) out[k++] = "ACGT"[rnd() & 3];              /* insertion */
        out[k++] = r < 14 ? "ACGT"[rnd() & 3] : s[i];         /* substitution */
    }
    out[k] = '\0';
    return out;
}

static int tuple_is(stIntTuple *t, int64_t a, int64_t b, int64_t c, int64_t d) {
    return stIntTuple_length(t) == 4 && stIntTuple_get(t, 0) == a && stIntTuple_get(t, 1) == b && stIntTuple_get(t, 2) == c &&
           stIntTuple_get(t, 3) == d;
}

/* ------------------------------------------------------------------------------------ host tests */

static void test_containers(void) {
    stList *l = stList_construct3(0, (void (*)(void *)) stIntTuple_destruct);
    for (int64_t i = 0; i < 100; i++) stList_append(l, stIntTuple_construct3(100 - i, i, 7));
    CHECK(stList_length(l) == 100);
    stList_sort(l, stIntTuple_cmpFn);
    CHECK(stIntTuple_get(stList_get(l, 0), 0) == 1 && stIntTuple_get(stList_get(l, 99), 0) == 100);
    stList_reverse(l);
    CHECK(stIntTuple_get(stList_get(l, 0), 0) == 100);
    stIntTuple *t = stList_pop(l);
    CHECK(stIntTuple_get(t, 0) == 1 && stList_length(l) == 99);
    stIntTuple *u = stIntTuple_construct3(1, 99, 7);
    CHECK(stIntTuple_equalsFn(t, u));
    stIntTuple_destruct(t);
    stIntTuple_destruct(u);
    stList_destruct(l);
}

/* test_hmm_* (tests/pairwiseAlignerTest.c:997-1089): write -> load round trip, normalisation */
static void test_hmm(StateMachineType type) {
    Hmm *hmm = hmm_constructEmpty(0.0, type);
    const int64_t S = hmm->stateNumber;
    CHECK(S == (type == fiveState || type == fiveStateAsymmetric ? 5 : 3));
    for (int64_t from = 0; from < S; from++)
        for (int64_t to = 0; to < S; to++) hmm_addToTransitionExpectation(hmm, from, to, (double) (from * S + to));
    for (int64_t s = 0; s < S; s++)
        for (int x = 0; x < 4; x++)
            for (int y = 0; y < 4; y++) hmm_addToEmissionsExpectation(hmm, s, (Symbol) x, (Symbol) y, (double) (s * 16 + x * 4 + y));
    hmm->likelihood = -123.5;
    char path[] = "/tmp/cpecan_hmm_XXXXXX";
    const int fd = mkstemp(path);
    CHECK(fd >= 0);
    FILE *fh = fdopen(fd, "w");
    hmm_write(hmm, fh);
    fclose(fh);
    Hmm *back = hmm_loadFromFile(path);
    remove(path);
    CHECK(back->type == type && back->stateNumber == S && back->likelihood == -123.5);
    for (int64_t i = 0; i < S * S; i++) CHECK(back->transitions[i] == (double) i);
    for (int64_t i = 0; i < S * 16; i++) CHECK(back->emissions[i] == (double) i);
    hmm_normalise(back);
    for (int64_t from = 0; from < S; from++) {
        double z = 0.0;
        for (int64_t to = 0; to < S; to++) z += (double) (from * S + to);
        for (int64_t to = 0; to < S; to++) CHECK(fabs(hmm_getTransition(back, from, to) - (double) (from * S + to) / z) < 1e-12);
    }
    for (int64_t s = 0; s < S; s++) {
        double z = 0.0;
        for (int i = 0; i < 16; i++) z += (double) (s * 16 + i);
        for (int x = 0; x < 4; x++)
            for (int y = 0; y < 4; y++)
                CHECK(fabs(hmm_getEmissionsExpectation(back, s, (Symbol) x, (Symbol) y) - (double) (s * 16 + x * 4 + y) / z) < 1e-12);
    }
    hmm_destruct(back);
    hmm_randomise(hmm);
    double z = 0.0;
    for (int64_t to = 0; to < S; to++) z += hmm_getTransition(hmm, 0, to);
    CHECK(fabs(z - 1.0) < 1e-9);
    StateMachine *sM = hmm_getStateMachine(hmm);
    CHECK(sM->type == type && sM->stateNumber == S && sM->matchState == 0);
    CHECK(sM->startStateProb(sM, 0) == 0.0 && sM->startStateProb(sM, 1) == LOG_ZERO);
    stateMachine_destruct(sM);
    hmm_destruct(hmm);
}

static void test_json(void) {
    char pj[] = "{\"threshold\": 0.25, \"diagonalExpansion\": 6, \"dynamicAnchorExpansion\": true, \"gapGamma\": 0.25, "
                "\"splitMatrixBiggerThanThis\": 100, \"alignAmbiguityCharacters\": false}";
    PairwiseAlignmentParameters *p = pairwiseAlignmentParameters_jsonParse(pj, strlen(pj));
    CHECK(p->threshold == 0.25 && p->diagonalExpansion == 6 && p->dynamicAnchorExpansion && p->gapGamma == 0.25f);
    CHECK(p->splitMatrixBiggerThanThis == 100 && !p->alignAmbiguityCharacters);
    CHECK(p->minDiagsBetweenTraceBack == 1000 && p->traceBackDiagonals == 40 && p->constraintDiagonalTrim == 14); /* defaults kept */
    pairwiseAlignmentBandingParameters_destruct(p);
    p = pairwiseAlignmentBandingParameters_construct();
    CHECK(p->threshold == 0.01 && p->diagonalExpansion == 20 && p->anchorMatrixBiggerThanThis == 500 * 500);
    CHECK(p->splitMatrixBiggerThanThis == 3000 * 3000 && p->gapGamma == 0.5f && !p->dynamicAnchorExpansion);
    pairwiseAlignmentBandingParameters_destruct(p);

    char hj[512] = "{\"type\": 2, \"transitions\": [";
    for (int i = 0; i < 9; i++) sprintf(hj + strlen(hj), "%s%d.5", i ? ", " : "", i);
    strcat(hj, "], \"emissions\": [");
    for (int i = 0; i < 48; i++) sprintf(hj + strlen(hj), "%s%d", i ? "," : "", i);
    strcat(hj, "], \"likelihood\": -7.25}");
    Hmm *hmm = hmm_jsonParse(hj, strlen(hj));
    CHECK(hmm->type == threeState && hmm->stateNumber == 3 && hmm->likelihood == -7.25);
    CHECK(hmm->transitions[8] == 8.5 && hmm->emissions[47] == 47.0);
    hmm_destruct(hmm);
}

/* test_getSplitPoints, tests/pairwiseAlignerTest.c:578-647 */
static void test_split_points(void) {
    const int64_t matrixSize = 2000 * 2000;
    stList *anchorPairs = stList_construct3(0, (void (*)(void *)) stIntTuple_destruct);
    int64_t lX = 3000, lY = 1000;
    stList *sp = getSplitPoints(anchorPairs, lX, lY, matrixSize, 0, 0);
    CHECK(stList_length(sp) == 1 && tuple_is(stList_get(sp, 0), 0, 0, lX, lY));
    stList_destruct(sp);
    lX = 20000;
    lY = 25000;
    sp = getSplitPoints(anchorPairs, lX, lY, matrixSize, 1, 1);
    CHECK(stList_length(sp) == 0);
    stList_destruct(sp);
    sp = getSplitPoints(anchorPairs, lX, lY, matrixSize, 1, 0);
    CHECK(stList_length(sp) == 1 && tuple_is(stList_get(sp, 0), 18000, 23000, lX, lY));
    stList_destruct(sp);
    sp = getSplitPoints(anchorPairs, lX, lY, matrixSize, 0, 1);
    CHECK(stList_length(sp) == 1 && tuple_is(stList_get(sp, 0), 0, 0, 2000, 2000));
    stList_destruct(sp);
    sp = getSplitPoints(anchorPairs, lX, lY, matrixSize, 0, 0);
    CHECK(stList_length(sp) == 2 && tuple_is(stList_get(sp, 0), 0, 0, 2000, 2000) && tuple_is(stList_get(sp, 1), 18000, 23000, lX, lY));
    stList_destruct(sp);
    const int64_t pts[8][2] = { { 2000, 2000 }, { 4002, 4001 }, { 5000, 5000 }, { 8000, 6000 }, { 9000, 9000 }, { 10000, 14000 }, { 15000, 15000 },
                                { 16000, 16000 } };
    for (int i = 0; i < 8; i++) stList_append(anchorPairs, stIntTuple_construct2(pts[i][0], pts[i][1]));
    sp = getSplitPoints(anchorPairs, lX, lY, matrixSize, 0, 0);
    CHECK(stList_length(sp) == 5);
    if (stList_length(sp) == 5) {
        CHECK(tuple_is(stList_get(sp, 0), 0, 0, 3001, 3001));
        CHECK(tuple_is(stList_get(sp, 1), 3002, 3001, 9500, 11001));
        CHECK(tuple_is(stList_get(sp, 2), 9501, 12000, 12001, 14500));
        CHECK(tuple_is(stList_get(sp, 3), 13000, 14501, 18000, 18001));
        CHECK(tuple_is(stList_get(sp, 4), 18001, 23000, 20000, 25000));
    }
    stList_destruct(sp);
    stList_destruct(anchorPairs);
}

static void test_diagonal_and_symbols(void) {
    /* test_diagonal, tests/pairwiseAlignerTest.c:17-59 */
    const Diagonal d = diagonal_construct(3, -1, 1);
    CHECK(diagonal_getXay(d) == 3 && diagonal_getMinXmy(d) == -1 && diagonal_getMaxXmy(d) == 1 && diagonal_getWidth(d) == 2);
    CHECK(diagonal_getXCoordinate(3, -1) == 1 && diagonal_getYCoordinate(3, -1) == 2);
    CHECK(diagonal_equals(d, diagonal_construct(3, -1, 1)) && !diagonal_equals(d, diagonal_construct(3, -1, 3)));
    /* test_symbol, :146-153 */
    Symbol *s = symbol_convertStringToSymbols("AaCcGgTtNn-", 11);
    const Symbol want[11] = { a, a, c, c, g, g, t, t, n, n, n };
    for (int i = 0; i < 11; i++) CHECK(s[i] == want[i]);
    free(s);
    CHECK(symbol_convertSymbolToChar(g) == 'G' && symbol_convertSymbolToChar(n) == 'N');
}

/* test_logAdd, tests/pairwiseAlignerTest.c:134-144: within 1e-3 of the exact value; plus the contract's edge cases */
static void test_log_add(void) {
    for (int i = 0; i < 100000; i++) {
        const double x = (double) rnd_int(-1000000, 1000000) / 10000.0, y = x + (double) rnd_int(-120000, 120000) / 10000.0;
        const double exact = (x > y ? x : y) + log1p(exp(-fabs(x - y)));
        CHECK(fabs(logAdd(x, y) - exact) < 0.001);
        CHECK(logAdd(x, y) == logAdd(y, x));
    }
    CHECK(logAdd(LOG_ZERO, 3.0) == 3.0 && logAdd(-2.0, LOG_ZERO) == -2.0 && logAdd(LOG_ZERO, LOG_ZERO) == LOG_ZERO);
    CHECK(logAdd(0.0, 7.5) == 7.5 && logAdd(0.0, 7.4999) > 7.4999 && logAdd(1.0, 1.0) == 1.0 + (double) 0.693203116424741f);
}

static void test_symbol_string(void) {
    SymbolString s = symbolString_construct("AcgTnX", 6);
    const Symbol want[6] = { a, c, g, t, n, n };
    CHECK(s.length == 6);
    for (int i = 0; i < 6; i++) CHECK(s.sequence[i] == want[i]);
    symbolString_destruct(s);
    s = symbolString_construct("", 0);
    CHECK(s.length == 0);
    symbolString_destruct(s);
}

/* ------------------------------------------------------------------------------------- gpu tests */

static int diag_is(Diagonal d, int64_t xay, int64_t l, int64_t r) { return d.xay == xay && d.xmyL == l && d.xmyR == r; }

static void test_bands(void) {
    stList *anchorPairs = stList_construct3(0, (void (*)(void *)) stIntTuple_destruct);
    stList_append(anchorPairs, stIntTuple_construct2(1, 0));
    stList_append(anchorPairs, stIntTuple_construct2(2, 1));
    stList_append(anchorPairs, stIntTuple_construct2(3, 3));
    Band *band = band_construct(anchorPairs, 6, 5, 2);
    BandIterator *it = bandIterator_construct(band);
    const int64_t want[12][3] = { { 0, 0, 0 }, { 1, -1, 1 }, { 2, -2, 2 }, { 3, -1, 3 }, { 4, -2, 4 }, { 5, -1, 3 },
                                  { 6, -2, 4 }, { 7, -3, 3 }, { 8, -2, 2 }, { 9, -1, 3 }, { 10, 0, 2 }, { 11, 1, 1 } };
    for (int i = 0; i < 12; i++) CHECK(diag_is(bandIterator_getNext(it), want[i][0], want[i][1], want[i][2]));
    CHECK(diag_is(bandIterator_getNext(it), 11, 1, 1)); /* clamps at the end */
    for (int i = 11; i >= 7; i--) CHECK(diag_is(bandIterator_getPrevious(it), want[i][0], want[i][1], want[i][2]));
    for (int i = 7; i <= 10; i++) CHECK(diag_is(bandIterator_getNext(it), want[i][0], want[i][1], want[i][2]));
    for (int i = 10; i >= 0; i--) CHECK(diag_is(bandIterator_getPrevious(it), want[i][0], want[i][1], want[i][2]));
    CHECK(diag_is(bandIterator_getPrevious(it), 0, 0, 0)); /* clamps at the start */
    CHECK(diag_is(bandIterator_getNext(it), 0, 0, 0));
    CHECK(diag_is(bandIterator_getNext(it), 1, -1, 1));
    bandIterator_destruct(it);
    band_destruct(band);
    stList_destruct(anchorPairs);
}

/* the four pairs of test_diagonalDPCalculations (tests/pairwiseAlignerTest.c:242-324) through the public wrapper */
static void test_kat(void) {
    StateMachine *sM = stateMachine5_construct(fiveState);
    PairwiseAlignmentParameters *p = pairwiseAlignmentBandingParameters_construct();
    p->threshold = 0.2;
    stList *anchors = stList_construct();
    stList *pairs = getAlignedPairsUsingAnchors(sM, "AGCG", "AGTTCG", anchors, p, 0, 0);
    stList_sort(pairs, stIntTuple_cmpFn);
    CHECK(stList_length(pairs) == 4);
    const int64_t want[4][3] = { { 8665179, 2, 4 }, { 9259684, 1, 1 }, { 9893294, 3, 5 }, { 9944673, 0, 0 } };
    for (int64_t i = 0; i < 4 && i < stList_length(pairs); i++) {
        stIntTuple *t = stList_get(pairs, i);
        CHECK(stIntTuple_get(t, 0) == want[i][0] && stIntTuple_get(t, 1) == want[i][1] && stIntTuple_get(t, 2) == want[i][2]);
    }
    stList_destruct(pairs);
    char sx[] = "AGCG", sy[] = "AGTTCG";
    const double lp = computeForwardProbability(sx, sy, anchors, p, sM, 0, 0);
    CHECK(fabs(lp - (-17.519321161239)) < 1e-9); /* SURVEY.md appendix A7 */
    stList_destruct(anchors);
    pairwiseAlignmentBandingParameters_destruct(p);
    stateMachine_destruct(sM);
}

/* checkAlignedPairs, tests/pairwiseAlignerTest.c:344-381: range, score bounds, uniqueness */
static void check_aligned_pairs(stList *pairs, int64_t lX, int64_t lY) {
    stList_sort(pairs, stIntTuple_cmpFn);
    char *seen = calloc((size_t) ((lX + 1) * (lY + 1)), 1);
    for (int64_t i = 0; i < stList_length(pairs); i++) {
        stIntTuple *t = stList_get(pairs, i);
        CHECK(stIntTuple_length(t) == 3);
        const int64_t score = stIntTuple_get(t, 0), x = stIntTuple_get(t, 1), y = stIntTuple_get(t, 2);
        CHECK(score > 0 && score <= PAIR_ALIGNMENT_PROB_1);
        CHECK(x >= 0 && x < lX && y >= 0 && y < lY);
        if (x >= 0 && x < lX && y >= 0 && y < lY) {
            CHECK(!seen[x * (lY + 1) + y]);
            seen[x * (lY + 1) + y] = 1;
        }
    }
    free(seen);
}

static void test_get_aligned_pairs(void) {
    StateMachine *machines[2] = { stateMachine5_construct(fiveState), stateMachine3_construct(threeState) };
    PairwiseAlignmentParameters *p = pairwiseAlignmentBandingParameters_construct();
    /* batched and one-pair forms agree, and every list is valid */
    enum { N = 24 };
    char *sX[N], *sY[N];
    for (int i = 0; i < N; i++) {
        sX[i] = random_sequence(rnd_int(0, 100));
        sY[i] = evolve(sX[i]);
    }
    for (int m = 0; m < 2; m++) {
        stList **batch = getAlignedPairsUsingAnchorsBatch(machines[m], N, (const char *const *) sX, (const char *const *) sY, NULL, p, NULL, NULL);
        for (int i = 0; i < N; i++) {
            check_aligned_pairs(batch[i], (int64_t) strlen(sX[i]), (int64_t) strlen(sY[i]));
            stList *one = getAlignedPairs(machines[m], sX[i], sY[i], p, 0, 0);
            stList_sort(one, stIntTuple_cmpFn);
            CHECK(stList_length(one) == stList_length(batch[i]));
            for (int64_t k = 0; k < stList_length(one) && k < stList_length(batch[i]); k++) CHECK(stIntTuple_equalsFn(stList_get(one, k), stList_get(batch[i], k)));
            stList_destruct(one);
            stList_destruct(batch[i]);
        }
        free(batch);
    }
    /* test_getAlignedPairsWithRaggedEnds (:676-715): a 100 bp core inside random flanks aligns on the shifted diagonal */
    for (int rep = 0; rep < 5; rep++) {
        char *core = random_sequence(100), *left = random_sequence(100), *right = random_sequence(100);
        char *x = malloc(301);
        snprintf(x, 301, "%s%s%s", left, core, right);
        stList *pairs = getAlignedPairs(machines[0], x, core, p, 1, 1);
        check_aligned_pairs(pairs, 300, 100);
        int64_t onDiagonal = 0;
        for (int64_t i = 0; i < stList_length(pairs); i++) {
            stIntTuple *t = stList_get(pairs, i);
            if (stIntTuple_get(t, 1) == stIntTuple_get(t, 2) + 100 && stIntTuple_get(t, 0) > PAIR_ALIGNMENT_PROB_1 / 2) onDiagonal++;
        }
        CHECK(onDiagonal >= 95);
        stList_destruct(pairs);
        free(core);
        free(left);
        free(right);
        free(x);
    }
    /* test_computeForwardProbability (:1157-1188): LOG_ZERO < logP(x,y) <= logP(x,x) <= 0 */
    for (int i = 0; i < N; i++) {
        if (strlen(sX[i]) == 0) continue;
        stList *none = stList_construct();
        const double pxy = computeForwardProbability(sX[i], sY[i], none, p, machines[0], 0, 0);
        const double pxx = computeForwardProbability(sX[i], sX[i], none, p, machines[0], 0, 0);
        CHECK(pxy > LOG_ZERO && pxy <= pxx && pxx <= 0.0);
        stList_destruct(none);
    }
    /* indel lists: gap probabilities are valid and the match list equals the plain call */
    {
        stList *m, *gx, *gy;
        getAlignedPairsWithIndels(machines[1], sX[3], sY[3], p, &m, &gx, &gy, 0, 0);
        stList *plain = getAlignedPairs(machines[1], sX[3], sY[3], p, 0, 0);
        CHECK(stList_length(m) == stList_length(plain));
        for (int64_t i = 0; i < stList_length(gx); i++) CHECK(stIntTuple_get(stList_get(gx, i), 0) <= PAIR_ALIGNMENT_PROB_1);
        stList_destruct(m);
        stList_destruct(gx);
        stList_destruct(gy);
        stList_destruct(plain);
    }
    for (int i = 0; i < N; i++) {
        free(sX[i]);
        free(sY[i]);
    }
    pairwiseAlignmentBandingParameters_destruct(p);
    stateMachine_destruct(machines[0]);
    stateMachine_destruct(machines[1]);
}

/* test_em_* (tests/pairwiseAlignerTest.c:1091-1155): EM on one pair from a random model does not lose likelihood */
static void test_em(StateMachineType type) {
    char *sX = random_sequence(150), *sY = evolve(sX);
    PairwiseAlignmentParameters *p = pairwiseAlignmentBandingParameters_construct();
    Hmm *hmm = hmm_constructEmpty(0.0, type);
    hmm_randomise(hmm);
    double previous = -INFINITY;
    for (int iteration = 0; iteration < 10; iteration++) {
        StateMachine *sM = hmm_getStateMachine(hmm);
        hmm_destruct(hmm);
        hmm = hmm_constructEmpty(0.000000000001, type);
        getExpectations(sM, hmm, sX, sY, p, 0, 0);
        CHECK(isfinite(hmm->likelihood));
        for (int64_t i = 0; i < hmm->stateNumber * hmm->stateNumber; i++) CHECK(hmm->transitions[i] >= 0.0);
        CHECK(previous <= hmm->likelihood * 0.95); /* likelihoods are negative: allow the reference's 5 % slack */
        previous = hmm->likelihood;
        hmm_normalise(hmm);
        stateMachine_destruct(sM);
    }
    hmm_destruct(hmm);
    pairwiseAlignmentBandingParameters_destruct(p);
    free(sX);
    free(sY);
}

/* --------------------------------------------------------------------------------------- driver */

static char *read_line(FILE *f) {
    char *line = NULL;
    size_t cap = 0;
    const ssize_t len = getline(&line, &cap, f);
    if (len < 0) {
        fprintf(stderr, "unexpected end of input\n");
        exit(2);
    }
    if (len > 0 && line[len - 1] == '\n') line[len - 1] = '\0';
    return line;
}

static void write_list(FILE *out, const char *tag, stList *l) {
    fprintf(out, "%s %" PRIi64, tag, stList_length(l));
    for (int64_t i = 0; i < stList_length(l); i++) {
        stIntTuple *t = stList_get(l, i);
        fprintf(out, " %" PRIi64 " %" PRIi64 " %" PRIi64, stIntTuple_get(t, 0), stIntTuple_get(t, 1), stIntTuple_get(t, 2));
    }
    fprintf(out, "\n");
}

static void write_hmm(FILE *out, const char *tag, Hmm *hmm) {
    fprintf(out, "%s", tag);
    for (int64_t i = 0; i < hmm->stateNumber * hmm->stateNumber; i++) fprintf(out, " %a", hmm->transitions[i]);
    for (int64_t i = 0; i < hmm->stateNumber * 16; i++) fprintf(out, " %a", hmm->emissions[i]);
    fprintf(out, " %a\n", hmm->likelihood);
}

/* what the reference's own wrappers pass as coordinateCorrectionFn (impl/pairwiseAligner.c:1259-1271, :1411-1429): shift the region's
 * tuples and move them, last first, onto the result list */
static void correct_lists(int64_t offsetX, int64_t offsetY, void *extraArgs, int nLists) {
    for (int l = 0; l < nLists; l++) {
        stList *sub = ((void **) extraArgs)[2 * l], *all = ((void **) extraArgs)[2 * l + 1];
        while (stList_length(sub) > 0) {
            stIntTuple *t = stList_pop(sub);
            stList_append(all, stIntTuple_construct3(stIntTuple_get(t, 0), stIntTuple_get(t, 1) + offsetX, stIntTuple_get(t, 2) + offsetY));
            stIntTuple_destruct(t);
        }
    }
}
static void correct_one(int64_t offsetX, int64_t offsetY, void *extraArgs) { correct_lists(offsetX, offsetY, extraArgs, 1); }
static void correct_three(int64_t offsetX, int64_t offsetY, void *extraArgs) { correct_lists(offsetX, offsetY, extraArgs, 3); }

/* getPosteriorProbsWithBanding / ...SplittingAlignmentsByLargeGaps with each of the three callbacks (inc/pairwiseAligner.h:245, :264) */
static void write_callback_forms(FILE *out, StateMachine *sM, const char *sX, const char *sY, stList *anchors, PairwiseAlignmentParameters *p, bool rl,
                                 bool rr) {
    const int64_t lX = (int64_t) strlen(sX), lY = (int64_t) strlen(sY);
    /* one region, no splitting, lists in the callback's own order */
    SymbolString x = symbolString_construct(sX, lX), y = symbolString_construct(sY, lY);
    stList *lists[6];
    for (int i = 0; i < 6; i++) lists[i] = stList_construct3(0, (void (*)(void *)) stIntTuple_destruct);
    getPosteriorProbsWithBanding(sM, anchors, x, y, p, rl, rr, diagonalCalculationPosteriorMatchProbs, lists);
    write_list(out, "banding", lists[0]);
    stList_destruct(lists[0]);
    lists[0] = stList_construct3(0, (void (*)(void *)) stIntTuple_destruct);
    getPosteriorProbsWithBanding(sM, anchors, x, y, p, rl, rr, diagonalCalculationPosteriorProbs, lists);
    write_list(out, "bandingM", lists[0]);
    write_list(out, "bandingX", lists[2]);
    write_list(out, "bandingY", lists[4]);
    Hmm *h = hmm_constructEmpty(0.0, sM->type);
    getPosteriorProbsWithBanding(sM, anchors, x, y, p, rl, rr, diagonalCalculationExpectations, h);
    write_hmm(out, "bandingE", h);
    hmm_destruct(h);
    symbolString_destruct(x);
    symbolString_destruct(y);
    /* split at large gaps, with the reference's correction functions */
    for (int i = 0; i < 6; i++) {
        stList_destruct(lists[i]);
        lists[i] = stList_construct3(0, (void (*)(void *)) stIntTuple_destruct);
    }
    getPosteriorProbsWithBandingSplittingAlignmentsByLargeGaps(sM, anchors, sX, sY, lX, lY, p, rl, rr, diagonalCalculationPosteriorMatchProbs, correct_one,
                                                               lists);
    write_list(out, "splitting", lists[1]);
    stList_destruct(lists[1]);
    lists[1] = stList_construct3(0, (void (*)(void *)) stIntTuple_destruct);
    getPosteriorProbsWithBandingSplittingAlignmentsByLargeGaps(sM, anchors, sX, sY, lX, lY, p, rl, rr, diagonalCalculationPosteriorProbs, correct_three,
                                                               lists);
    write_list(out, "splittingM", lists[1]);
    write_list(out, "splittingX", lists[3]);
    write_list(out, "splittingY", lists[5]);
    h = hmm_constructEmpty(0.0, sM->type);
    getPosteriorProbsWithBandingSplittingAlignmentsByLargeGaps(sM, anchors, sX, sY, lX, lY, p, rl, rr, diagonalCalculationExpectations, NULL, h);
    write_hmm(out, "splittingE", h);
    hmm_destruct(h);
    for (int i = 0; i < 6; i++) stList_destruct(lists[i]);
}

/*
 * IN:  line 1: type hmmFile|- paramsJson...   (the rest of the line is the parameter JSON)
 *      line 2: n
 *      per problem: "raggedLeft raggedRight nAnchors", sX, sY ("-" = empty), "x y e x y e ..."
 */
static int run_file(const char *inPath, const char *outPath) {
    FILE *in = fopen(inPath, "r"), *out = fopen(outPath, "w");
    if (in == NULL || out == NULL) {
        fprintf(stderr, "cannot open %s / %s\n", inPath, outPath);
        return 2;
    }
    char *head = read_line(in);
    int type;
    char hmmFile[512];
    int used = 0;
    if (sscanf(head, "%d %511s %n", &type, hmmFile, &used) < 2) return 2;
    PairwiseAlignmentParameters *p = pairwiseAlignmentParameters_jsonParse(head + used, strlen(head + used));
    StateMachine *sM;
    if (strcmp(hmmFile, "-") != 0) {
        Hmm *model = hmm_loadFromFile(hmmFile);
        sM = hmm_getStateMachine(model);
        hmm_destruct(model);
    } else {
        sM = (type == fiveState || type == fiveStateAsymmetric) ? stateMachine5_construct((StateMachineType) type)
                                                                : stateMachine3_construct((StateMachineType) type);
    }
    free(head);
    char *line = read_line(in);
    const int64_t n = atoll(line);
    free(line);
    char **sX = malloc((size_t) (n + 1) * sizeof(char *)), **sY = malloc((size_t) (n + 1) * sizeof(char *));
    stList **anchors = malloc((size_t) (n + 1) * sizeof(stList *));
    bool *rl = malloc((size_t) n + 1), *rr = malloc((size_t) n + 1);
    for (int64_t i = 0; i < n; i++) {
        line = read_line(in);
        int l, r;
        long long nA;
        sscanf(line, "%d %d %lld", &l, &r, &nA);
        free(line);
        rl[i] = l != 0;
        rr[i] = r != 0;
        sX[i] = read_line(in);
        sY[i] = read_line(in);
        if (strcmp(sX[i], "-") == 0) sX[i][0] = '\0';
        if (strcmp(sY[i], "-") == 0) sY[i][0] = '\0';
        anchors[i] = stList_construct3(0, (void (*)(void *)) stIntTuple_destruct);
        line = read_line(in);
        char *q = line;
        for (long long k = 0; k < nA; k++) {
            long long x, y, e;
            int adv = 0;
            sscanf(q, "%lld %lld %lld %n", &x, &y, &e, &adv);
            q += adv;
            stList_append(anchors[i], stIntTuple_construct3(x, y, e));
        }
        free(line);
    }
    /* batched forms */
    stList **m, **gx, **gy;
    getAlignedPairsWithIndelsUsingAnchorsBatch(sM, n, (const char *const *) sX, (const char *const *) sY, anchors, p, &m, &gx, &gy, rl, rr);
    stList **only = getAlignedPairsUsingAnchorsBatch(sM, n, (const char *const *) sX, (const char *const *) sY, anchors, p, rl, rr);
    double *fwd = malloc((size_t) (n + 1) * sizeof(double));
    computeForwardProbabilityBatch(sM, n, (const char *const *) sX, (const char *const *) sY, anchors, p, rl, rr, fwd);
    Hmm *total = hmm_constructEmpty(0.0, sM->type);
    getExpectationsUsingAnchorsBatch(sM, total, n, (const char *const *) sX, (const char *const *) sY, anchors, p, rl, rr);
    for (int64_t i = 0; i < n; i++) {
        fprintf(out, "problem %" PRIi64 "\n", i);
        write_list(out, "match", m[i]);
        write_list(out, "gapX", gx[i]);
        write_list(out, "gapY", gy[i]);
        write_list(out, "only", only[i]);
        fprintf(out, "forward %a\n", fwd[i]);
        /* the reference's one-pair signatures */
        stList *one = getAlignedPairsUsingAnchors(sM, sX[i], sY[i], anchors[i], p, rl[i], rr[i]);
        write_list(out, "one", one);
        stList_destruct(one);
        Hmm *h = hmm_constructEmpty(0.0, sM->type);
        getExpectationsUsingAnchors(sM, h, sX[i], sY[i], anchors[i], p, rl[i], rr[i]);
        write_hmm(out, "expectations", h);
        hmm_destruct(h);
        fprintf(out, "forward1 %a\n", computeForwardProbability(sX[i], sY[i], anchors[i], p, sM, rl[i], rr[i]));
        write_callback_forms(out, sM, sX[i], sY[i], anchors[i], p, rl[i], rr[i]);
        stList_destruct(m[i]);
        stList_destruct(gx[i]);
        stList_destruct(gy[i]);
        stList_destruct(only[i]);
        stList_destruct(anchors[i]);
        free(sX[i]);
        free(sY[i]);
    }
    write_hmm(out, "total", total);
    hmm_destruct(total);
    fclose(in);
    fclose(out);
    cpecan_shutdown();
    return 0;
}

int main(int argc, char **argv) {
    if (argc >= 4 && strcmp(argv[1], "run") == 0) return run_file(argv[2], argv[3]);
    if (argc >= 2 && strcmp(argv[1], "host") == 0) {
        test_containers();
        for (int type = 0; type < 4; type++) test_hmm((StateMachineType) type);
        test_json();
        test_split_points();
        test_diagonal_and_symbols();
        test_log_add();
        test_symbol_string();
    } else if (argc >= 2 && strcmp(argv[1], "gpu") == 0) {
        test_bands();
        test_kat();
        test_get_aligned_pairs();
        test_em(fiveState);
        test_em(fiveStateAsymmetric);
        test_em(threeState);
        test_em(threeStateAsymmetric);
        cpecan_shutdown();
    } else {
        fprintf(stderr, "usage: %s host | gpu | run IN OUT\n", argv[0]);
        return 2;
    }
    printf("%d checks, %d failures\n", g_checks, g_failures);
    return g_failures == 0 ? 0 : 1;
}
