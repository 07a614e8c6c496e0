/*
 * cpecan/stateMachine.h -- the pair-HMM model interface of cPecan's hot path, served by libcpecan.so.
 *
 * Same names, argument meaning and file formats as the reference's inc/stateMachine.h (cited per item), so a
 * caller of cPecan's library compiles against this header unchanged.  Behind it the model is a flat, vtable-free
 * CpbModel (include/cpecan_b200.h) that is shipped to the device; the DP itself never runs on the host.
 */
#ifndef CPECAN_STATEMACHINE_H_
#define CPECAN_STATEMACHINE_H_

#include "cpecan/sonLibLite.h"

#ifdef __cplusplus
extern "C" {
#endif

#define SYMBOL_NUMBER 5
#define SYMBOL_NUMBER_NO_N 4

/* inc/stateMachine.h:16-22 */
typedef enum { a = 0, c = 1, g = 2, t = 3, n = 4 } Symbol;

/* inc/stateMachine.h:28-33 */
typedef enum { fiveState = 0, fiveStateAsymmetric = 1, threeState = 2, threeStateAsymmetric = 3 } StateMachineType;

typedef struct _stateMachine StateMachine;

/*
 * inc/stateMachine.h:37-55.  The public fields and the four boundary-vector callbacks keep their meaning.
 * cellCalculate exists for layout compatibility only: the per-cell recurrence (impl/stateMachine.c:450-480,
 * :689-714) is compiled into the CUDA kernels, so calling it aborts with a message instead of running a
 * host-side cell update.
 */
struct _stateMachine {
    StateMachineType type;
    int64_t stateNumber;
    int64_t matchState;
    int64_t gapXState;
    int64_t gapYState;
    double (*startStateProb)(StateMachine *sM, int64_t state);
    double (*endStateProb)(StateMachine *sM, int64_t state);
    double (*raggedEndStateProb)(StateMachine *sM, int64_t state);
    double (*raggedStartStateProb)(StateMachine *sM, int64_t state);
    void (*cellCalculate)(StateMachine *sM, double *current, double *lower, double *middle, double *upper, Symbol cX, Symbol cY,
                          void (*doTransition)(double *, double *, int64_t, int64_t, double, double, void *), void *extraArgs);
};

/* inc/stateMachine.h:61-67: model file contents / EM expectation accumulator */
typedef struct _hmm {
    StateMachineType type;
    double *transitions; /* [from * stateNumber + to] */
    double *emissions;   /* [state * 16 + x * 4 + y] */
    double likelihood;
    int64_t stateNumber;
} Hmm;

Hmm *hmm_constructEmpty(double pseudoExpectation, StateMachineType type); /* :69 */
void hmm_randomise(Hmm *hmm);                                             /* :71 */
void hmm_destruct(Hmm *hmmExpectations);                                  /* :73 */
void hmm_write(Hmm *hmmExpectations, FILE *fileHandle);                   /* :75, text format of impl/stateMachine.c:133-143 */
void hmm_addToTransitionExpectation(Hmm *hmmExpectations, int64_t from, int64_t to, double p);
double hmm_getTransition(Hmm *hmmExpectations, int64_t from, int64_t to);
void hmm_setTransition(Hmm *hmm, int64_t from, int64_t to, double p);
void hmm_addToEmissionsExpectation(Hmm *hmmExpectations, int64_t state, Symbol x, Symbol y, double p);
double hmm_getEmissionsExpectation(Hmm *hmm, int64_t state, Symbol x, Symbol y);
void hmm_setEmissionsExpectation(Hmm *hmm, int64_t state, Symbol x, Symbol y, double p);
Hmm *hmm_loadFromFile(const char *fileName); /* :89 */
Hmm *hmm_jsonParse(char *buf, size_t r);     /* :91, keys type / transitions / emissions / likelihood */
void hmm_normalise(Hmm *hmm);                /* :93 */

StateMachine *hmm_getStateMachine(Hmm *hmm);                  /* :95 */
StateMachine *stateMachine5_construct(StateMachineType type); /* :97 */
StateMachine *stateMachine3_construct(StateMachineType type); /* :99 */
void stateMachine_destruct(StateMachine *stateMachine);       /* :101 */

#ifdef __cplusplus
}
#endif
#endif /* CPECAN_STATEMACHINE_H_ */
