/*
 * bench_capi.c -- end-to-end time of the drop-in API: getAlignedPairsUsingAnchorsBatch (include/cpecan/pairwiseAligner.h) called the
 * way a cPecan caller would, with NUL-terminated strings and stLists of anchor tuples in, stLists of (pInt, x, y) tuples out, the
 * result lists walked once and destructed.  bench.py writes the synthetic workload to a file, runs this program and reports the
 * figure as `e2e_capi` beside the flat C-ABI `e2e`.
 *
 *   bench_capi WORKLOAD.bin STEPS WARMUP        -> one JSON line on stdout
 *   bench_capi WORKLOAD.bin STEPS WARMUP em     -> EM iterations over a resident batch instead (expectation pass + all-reduce)
 *   WORKLOAD.bin: int64 n; int64 xOff[n+1], yOff[n+1], aOff[n+1]; char seqX[xOff[n]], seqY[yOff[n]]; int64 anchors[3 * aOff[n]]
 */
#include <inttypes.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "cpecan/pairwiseAligner.h"

static double now(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double) ts.tv_sec + 1e-9 * (double) ts.tv_nsec;
}

static void *read_block(FILE *f, size_t bytes) {
    void *p = malloc(bytes ? bytes : 1);
    if (p == NULL || fread(p, 1, bytes, f) != bytes) {
        fprintf(stderr, "bench_capi: short read\n");
        exit(2);
    }
    return p;
}

int main(int argc, char **argv) {
    if (argc < 4) {
        fprintf(stderr, "usage: %s WORKLOAD.bin STEPS WARMUP\n", argv[0]);
        return 2;
    }
    FILE *f = fopen(argv[1], "rb");
    if (f == NULL) {
        perror(argv[1]);
        return 2;
    }
    const int steps = atoi(argv[2]), warmup = atoi(argv[3]);
    int64_t n;
    if (fread(&n, sizeof(n), 1, f) != 1) return 2;
    int64_t *xOff = read_block(f, (size_t) (n + 1) * 8), *yOff = read_block(f, (size_t) (n + 1) * 8), *aOff = read_block(f, (size_t) (n + 1) * 8);
    char *seqX = read_block(f, (size_t) xOff[n]), *seqY = read_block(f, (size_t) yOff[n]);
    int64_t *anchors = read_block(f, (size_t) aOff[n] * 24);
    fclose(f);

    /* the caller's data structures (not timed: they exist before the call) */
    char **sX = malloc((size_t) n * sizeof(char *)), **sY = malloc((size_t) n * sizeof(char *));
    stList **anchorLists = malloc((size_t) n * sizeof(stList *));
    for (int64_t i = 0; i < n; i++) {
        const int64_t lX = xOff[i + 1] - xOff[i], lY = yOff[i + 1] - yOff[i];
        sX[i] = malloc((size_t) lX + 1);
        sY[i] = malloc((size_t) lY + 1);
        memcpy(sX[i], seqX + xOff[i], (size_t) lX);
        memcpy(sY[i], seqY + yOff[i], (size_t) lY);
        sX[i][lX] = sY[i][lY] = '\0';
        anchorLists[i] = stList_construct3(0, (void (*)(void *)) stIntTuple_destruct);
        for (int64_t a = aOff[i]; a < aOff[i + 1]; a++) stList_append(anchorLists[i], stIntTuple_construct3(anchors[3 * a], anchors[3 * a + 1], anchors[3 * a + 2]));
    }
    StateMachine *sM = stateMachine5_construct(fiveState);
    PairwiseAlignmentParameters *p = pairwiseAlignmentBandingParameters_construct();

    if (argc >= 5 && strcmp(argv[4], "em") == 0) {
        /* EM iterations over a resident batch (cpecanResidentBatch_*): the alignments go to the devices once; an iteration is one
         * expectation pass per device and one NCCL all-reduce of the expectation totals ($CPECAN_DEVICES GPUs) */
        CpecanResidentBatch *rb = cpecanResidentBatch_construct(n, (const char *const *) sX, (const char *const *) sY, anchorLists, p, NULL, NULL);
        double tIter = 0.0, likelihood = 0.0, t00 = 0.0;
        for (int it = 0; it < warmup + steps; it++) {
            Hmm *hmm = hmm_constructEmpty(0.0, fiveState);
            const double t0 = now();
            cpecanResidentBatch_getExpectations(rb, sM, hmm, p);
            if (it >= warmup) tIter += now() - t0;
            likelihood = hmm->likelihood;
            t00 = hmm->transitions[0];
            hmm_destruct(hmm);
        }
        printf("{\"pairs\": %" PRIi64 ", \"devices\": %d, \"steps\": %d, \"s_per_em_iteration\": %.6f, \"likelihood\": %.17g, \"t00\": %.17g}\n", n,
               cpecan_getDeviceCount(), steps, tIter / steps, likelihood, t00);
        cpecanResidentBatch_destruct(rb);
        cpecan_shutdown();
        return 0;
    }
    int64_t tuples = 0, checksum = 0;
    double tCall = 0.0, tWalk = 0.0, tFree = 0.0;
    for (int it = 0; it < warmup + steps; it++) {
        const double t0 = now();
        stList **lists = getAlignedPairsUsingAnchorsBatch(sM, n, (const char *const *) sX, (const char *const *) sY, anchorLists, p, NULL, NULL);
        const double t1 = now();
        int64_t count = 0, sum = 0;
        for (int64_t i = 0; i < n; i++) { /* what a caller does first: look at every pair once */
            const int64_t len = stList_length(lists[i]);
            count += len;
            for (int64_t k = 0; k < len; k++) sum += stIntTuple_get(stList_get(lists[i], k), 0);
        }
        const double t2 = now();
        for (int64_t i = 0; i < n; i++) stList_destruct(lists[i]);
        free(lists);
        const double t3 = now();
        if (it >= warmup) {
            tCall += t1 - t0;
            tWalk += t2 - t1;
            tFree += t3 - t2;
        }
        tuples = count;
        checksum = sum;
    }
    printf("{\"pairs\": %" PRIi64 ", \"steps\": %d, \"tuples\": %" PRIi64 ", \"weight_sum\": %" PRIi64
           ", \"s_per_step\": %.6f, \"s_call\": %.6f, \"s_walk\": %.6f, \"s_destruct\": %.6f, \"devices\": %d}\n",
           n, steps, tuples, checksum, (tCall + tWalk + tFree) / steps, tCall / steps, tWalk / steps, tFree / steps, cpecan_getDeviceCount());
    cpecan_shutdown();
    return 0;
}
