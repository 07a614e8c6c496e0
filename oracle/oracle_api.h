/*
 * oracle/oracle_api.h -- TEST INFRASTRUCTURE ONLY.
 *
 * One flat C interface implemented twice:
 *   oracle/_ref/libcpecan_ref.so   the reference's own unmodified sources (built from /root/reference
 *                                  against oracle/shim by oracle/Makefile) behind oracle/ref_driver.c
 *   oracle/_build/libcpecan_oracle.so   this repo's plain-C restatement (oracle/pairhmm_oracle.c)
 * Tests load either through ctypes and compare it with the CUDA product.  Nothing under
 * cpecan_b200/ or include/ may include, link or call this.
 */
#ifndef ORACLE_API_H_
#define ORACLE_API_H_

#include <stdint.h>

typedef struct {
    double threshold;                   /* pairwiseAligner.h:29 */
    int64_t minDiagsBetweenTraceBack;   /* :30 */
    int64_t traceBackDiagonals;         /* :31 */
    int64_t diagonalExpansion;          /* :32 */
    int64_t constraintDiagonalTrim;     /* :33 */
    int64_t splitMatrixBiggerThanThis;  /* :36 */
    int64_t dynamicAnchorExpansion;     /* :39 */
} OrcParams;

typedef struct {
    int64_t type;     /* StateMachineType, stateMachine.h:28-33: 0 fiveState, 1 fiveStateAsymmetric, 2 threeState, 3 threeStateAsymmetric */
    int64_t fromHmm;  /* 0: stateMachine{5,3}_construct defaults; 1: hmm_getStateMachine(transitions, emissions) */
    double transitions[25]; /* row-major from*S+to, S*S used */
    double emissions[80];   /* state*16 + x*4 + y, S*16 used */
} OrcModel;

#define ORC_HMM_LEN(S) ((S) * (S) + (S) * 16 + 1) /* transitions, emissions, likelihood */

#ifdef __cplusplus
extern "C" {
#endif

void orc_default_params(OrcParams *p);

double orc_logadd(double x, double y);

/* writes (xay, xmyL, xmyR) for xay = 0..lX+lY; returns lX+lY+1 */
int64_t orc_band(const int64_t *anchors, int64_t nAnchors, int64_t lX, int64_t lY, int64_t expansion, int dynamic,
                 int64_t *out);

/* writes (x1,y1,x2,y2) per region; returns number of regions (may exceed cap; only cap are written) */
int64_t orc_split_points(const int64_t *anchors, int64_t nAnchors, int64_t lX, int64_t lY, int64_t splitMatrixBiggerThanThis,
                         int raggedLeft, int raggedRight, int64_t *out, int64_t cap);

/* transition / emission table of the model as seen by the DP:
 * out[0..S-1] start, [S..2S-1] ragged start, [2S..3S-1] end, [3S..4S-1] ragged end,
 * then for every (cX,cY) in 5x5 and every transition the model issues, in issue order:
 * (group, from, to, eP, tP) as 5 doubles; returns number of doubles written. */
int64_t orc_model_dump(const OrcModel *m, double *out, int64_t cap);

/* (pInt, x, y) triples in the reference's emission order; returns count (only cap written) */
int64_t orc_aligned_pairs(const OrcModel *m, const OrcParams *p, const char *sX, const char *sY, const int64_t *anchors,
                          int64_t nAnchors, int raggedLeft, int raggedRight, int64_t *out, int64_t cap);

/* three lists (match, gapX, gapY); counts[3] receives lengths */
void orc_aligned_pairs_with_indels(const OrcModel *m, const OrcParams *p, const char *sX, const char *sY,
                                   const int64_t *anchors, int64_t nAnchors, int raggedLeft, int raggedRight,
                                   int64_t *outMatch, int64_t *outGapX, int64_t *outGapY, int64_t cap, int64_t *counts);

/* hmm is ORC_HMM_LEN(S) doubles, accumulated into (+=) */
void orc_expectations(const OrcModel *m, const OrcParams *p, const char *sX, const char *sY, const int64_t *anchors,
                      int64_t nAnchors, int raggedLeft, int raggedRight, double *hmm);

double orc_forward_prob(const OrcModel *m, const OrcParams *p, const char *sX, const char *sY, const int64_t *anchors,
                        int64_t nAnchors, int raggedLeft, int raggedRight);

/* CPU baseline: nPairs independent problems farmed over nThreads pthreads.
 * mode 0: aligned pairs (counts[i] = #pairs, checksum[i] = sum of pInt*(x+1)+y);  mode 1: expectations into hmmOut (summed). */
void orc_batch(const OrcModel *m, const OrcParams *p, int64_t nPairs, const char *seqX, const int64_t *xOff, const char *seqY,
               const int64_t *yOff, const int64_t *anchors, const int64_t *aOff, const uint8_t *raggedLeft,
               const uint8_t *raggedRight, int mode, int nThreads, int64_t *counts, int64_t *checksum, double *hmmOut);

const char *orc_identity(void); /* "reference" or "port" */

#ifdef __cplusplus
}
#endif
#endif
