/*
 * host/stateMachine.c -- cPecan's model objects (Hmm, StateMachine) on the host side of libcpecan.so.
 *
 * The reference keeps a model as a struct of log-probabilities plus a vtable whose cellCalculate is called once
 * per DP cell (impl/stateMachine.c:377-521, :631-745).  Here a StateMachine is a thin wrapper around the flat
 * CpbModel image that cpb_batch_run ships to the device; the numbers in that image come from
 * cpb_model_default / cpb_model_from_hmm (cpecan_b200/csrc/model.c), which restate the reference's constructors
 * and Hmm loaders.  This file owns only the Hmm container, its text / JSON formats, and the wrapper.
 */
#include <inttypes.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "cpecan/stateMachine.h"
#include "cpecan_b200.h"
#include "host_internal.h"
#include "minijson.h"

/* ---------------------------------------------------------------------------------------------- Hmm */

static int64_t states_of(StateMachineType type) {
    switch (type) {
    case fiveState:
    case fiveStateAsymmetric:
        return 5;
    case threeState:
    case threeStateAsymmetric:
        return 3;
    default:
        st_errAbort("Unrecognised state type: %i\n", (int) type);
    }
    return 0;
}

Hmm *hmm_constructEmpty(double pseudoExpectation, StateMachineType type) {
    const int64_t S = states_of(type);
    Hmm *hmm = cpecan_malloc(sizeof(Hmm));
    hmm->type = type;
    hmm->stateNumber = S;
    hmm->transitions = cpecan_malloc((size_t) (S * S) * sizeof(double));
    hmm->emissions = cpecan_malloc((size_t) (S * 16) * sizeof(double));
    for (int64_t i = 0; i < S * S; i++) hmm->transitions[i] = pseudoExpectation;
    for (int64_t i = 0; i < S * 16; i++) hmm->emissions[i] = pseudoExpectation;
    hmm->likelihood = 0.0;
    return hmm;
}

void hmm_destruct(Hmm *hmm) {
    if (hmm == NULL) return;
    free(hmm->transitions);
    free(hmm->emissions);
    free(hmm);
}

double hmm_getTransition(Hmm *hmm, int64_t from, int64_t to) { return hmm->transitions[from * hmm->stateNumber + to]; }
void hmm_setTransition(Hmm *hmm, int64_t from, int64_t to, double p) { hmm->transitions[from * hmm->stateNumber + to] = p; }
void hmm_addToTransitionExpectation(Hmm *hmm, int64_t from, int64_t to, double p) { hmm->transitions[from * hmm->stateNumber + to] += p; }

static double *emission_slot(Hmm *hmm, int64_t state, Symbol x, Symbol y) {
    return &hmm->emissions[state * 16 + (int64_t) x * SYMBOL_NUMBER_NO_N + (int64_t) y];
}
double hmm_getEmissionsExpectation(Hmm *hmm, int64_t state, Symbol x, Symbol y) { return *emission_slot(hmm, state, x, y); }
void hmm_setEmissionsExpectation(Hmm *hmm, int64_t state, Symbol x, Symbol y, double p) { *emission_slot(hmm, state, x, y) = p; }
void hmm_addToEmissionsExpectation(Hmm *hmm, int64_t state, Symbol x, Symbol y, double p) { *emission_slot(hmm, state, x, y) += p; }

/* impl/stateMachine.c:88-112: every transition row and every state's emission table is scaled to sum to one */
void hmm_normalise(Hmm *hmm) {
    const int64_t S = hmm->stateNumber;
    for (int64_t from = 0; from < S; from++) {
        double total = 0.0;
        for (int64_t to = 0; to < S; to++) total += hmm->transitions[from * S + to];
        for (int64_t to = 0; to < S; to++) hmm->transitions[from * S + to] = hmm->transitions[from * S + to] / total;
    }
    for (int64_t state = 0; state < S; state++) {
        double total = 0.0;
        for (int64_t i = 0; i < 16; i++) total += hmm->emissions[state * 16 + i];
        for (int64_t i = 0; i < 16; i++) hmm->emissions[state * 16 + i] = hmm->emissions[state * 16 + i] / total;
    }
}

/* impl/stateMachine.c:114-131; the reference draws from sonLib's st_random(), here from drand48() */
void hmm_randomise(Hmm *hmm) {
    const int64_t S = hmm->stateNumber;
    for (int64_t i = 0; i < S * S; i++) hmm->transitions[i] = drand48();
    for (int64_t i = 0; i < S * 16; i++) hmm->emissions[i] = drand48();
    hmm_normalise(hmm);
}

/* text format of impl/stateMachine.c:133-143: "type \t S*S transitions \t likelihood \n 16*S emissions \t \n", all %f */
void hmm_write(Hmm *hmm, FILE *fileHandle) {
    const int64_t S = hmm->stateNumber;
    fprintf(fileHandle, "%i\t", (int) hmm->type);
    for (int64_t i = 0; i < S * S; i++) fprintf(fileHandle, "%f\t", hmm->transitions[i]);
    fprintf(fileHandle, "%f\n", hmm->likelihood);
    for (int64_t i = 0; i < S * 16; i++) fprintf(fileHandle, "%f\t", hmm->emissions[i]);
    fprintf(fileHandle, "\n");
}

/* one line of the file as whitespace-separated tokens; returns the token count, tokens point into *line */
static int64_t read_tokens(FILE *fH, char **line, char ***tokens) {
    size_t cap = 0;
    *line = NULL;
    if (getline(line, &cap, fH) < 0) {
        free(*line);
        *line = NULL;
        *tokens = NULL;
        return 0;
    }
    int64_t n = 0, tcap = 64;
    char **tok = cpecan_malloc((size_t) tcap * sizeof(char *));
    char *save = NULL;
    for (char *w = strtok_r(*line, " \t\r\n", &save); w != NULL; w = strtok_r(NULL, " \t\r\n", &save)) {
        if (n == tcap) {
            tcap *= 2;
            tok = realloc(tok, (size_t) tcap * sizeof(char *));
            if (tok == NULL) st_errAbort("cpecan: out of memory reading a model file");
        }
        tok[n++] = w;
    }
    *tokens = tok;
    return n;
}

Hmm *hmm_loadFromFile(const char *fileName) {
    FILE *fH = fopen(fileName, "r");
    if (fH == NULL) st_errAbort("Could not open the input state machine file %s\n", fileName);
    char *line, **tok;
    int64_t nTok = read_tokens(fH, &line, &tok);
    if (nTok < 2) st_errAbort("Got an empty line in the input state machine file %s\n", fileName);
    int type;
    if (sscanf(tok[0], "%i", &type) != 1) st_errAbort("Failed to parse state number (int) from string: %s\n", tok[0]);
    Hmm *hmm = hmm_constructEmpty(0.0, (StateMachineType) type);
    const int64_t S = hmm->stateNumber;
    if (nTok != S * S + 2)
        st_errAbort("Got the wrong number of transitions in the input state machine file %s, got %" PRIi64 " instead of %" PRIi64 "\n", fileName,
                    nTok, S * S + 2);
    for (int64_t i = 0; i < S * S; i++) {
        if (sscanf(tok[i + 1], "%lf", &hmm->transitions[i]) != 1) st_errAbort("Failed to parse transition prob (float) from string: %s\n", tok[i + 1]);
    }
    if (sscanf(tok[nTok - 1], "%lf", &hmm->likelihood) != 1) st_errAbort("Failed to parse likelihood (float) from string: %s\n", tok[nTok - 1]);
    free(tok);
    free(line);

    nTok = read_tokens(fH, &line, &tok);
    if (nTok != S * 16)
        st_errAbort("Got the wrong number of emissions in the input state machine file %s, got %" PRIi64 " instead of %" PRIi64 "\n", fileName, nTok,
                    S * 16);
    for (int64_t i = 0; i < S * 16; i++) {
        if (sscanf(tok[i], "%lf", &hmm->emissions[i]) != 1) st_errAbort("Failed to parse emission prob (float) from string: %s\n", tok[i]);
    }
    free(tok);
    free(line);
    fclose(fH);
    return hmm;
}

/* JSON form (impl/stateMachine.c:204-253): {"type": T, "transitions": [...], "emissions": [...], "likelihood": L}; "type" first */
typedef struct {
    Hmm *hmm;
    int gotTransitions, gotEmissions;
} HmmJsonState;

static int hmm_member(MiniJson *j, const char *key, void *extra) {
    HmmJsonState *st = extra;
    if (strcmp(key, "type") == 0) {
        double v;
        if (minijson_number(j, &v) != 0) return -1;
        if (st->hmm != NULL) st_errAbort("ERROR: duplicate type key in hmm json\n");
        st->hmm = hmm_constructEmpty(0.0, (StateMachineType) (int) v);
        return 0;
    }
    if (st->hmm == NULL) st_errAbort("ERROR: Unrecognised key in polish params json: %s\n", key); /* the reference insists on "type" first */
    const int64_t S = st->hmm->stateNumber;
    if (strcmp(key, "transitions") == 0) {
        st->gotTransitions = 1;
        return minijson_number_array(j, st->hmm->transitions, S * S);
    }
    if (strcmp(key, "emissions") == 0) {
        st->gotEmissions = 1;
        return minijson_number_array(j, st->hmm->emissions, S * 16);
    }
    if (strcmp(key, "likelihood") == 0) return minijson_number(j, &st->hmm->likelihood);
    st_errAbort("ERROR: Unrecognised key in hmm json: %s\n", key);
    return -1;
}

Hmm *hmm_jsonParse(char *buf, size_t r) {
    MiniJson j;
    minijson_init(&j, buf, r);
    HmmJsonState st = { NULL, 0, 0 };
    if (minijson_object(&j, hmm_member, &st) != 0) st_errAbort("ERROR: could not parse hmm json: %s\n", j.error);
    if (st.hmm == NULL) st_errAbort("ERROR: too few tokens to parse in hmm json\n");
    if (!st.gotEmissions) st_errAbort("ERROR: Did not find emissions specified in json HMM\n");
    if (!st.gotTransitions) st_errAbort("ERROR: Did not find transitions specified in json HMM\n");
    return st.hmm;
}

/* ------------------------------------------------------------------------------------- StateMachine */

static double start_prob(StateMachine *sM, int64_t state) { return cpecan_model_of(sM)->start[state]; }
static double end_prob(StateMachine *sM, int64_t state) { return cpecan_model_of(sM)->end[state]; }
static double ragged_end_prob(StateMachine *sM, int64_t state) { return cpecan_model_of(sM)->raggedEnd[state]; }
static double ragged_start_prob(StateMachine *sM, int64_t state) { return cpecan_model_of(sM)->raggedStart[state]; }

static void cell_calculate_unavailable(StateMachine *sM, double *current, double *lower, double *middle, double *upper, Symbol cX, Symbol cY,
                                       void (*doTransition)(double *, double *, int64_t, int64_t, double, double, void *), void *extraArgs) {
    (void) sM, (void) current, (void) lower, (void) middle, (void) upper, (void) cX, (void) cY, (void) doTransition, (void) extraArgs;
    st_errAbort("cpecan: StateMachine.cellCalculate is not callable in this library: the cell recurrence runs inside the CUDA kernels "
                "(use getAlignedPairs*, getExpectations* or computeForwardProbability)");
}

static StateMachine *wrap(const CpbModel *m) {
    CpecanStateMachine *w = cpecan_malloc(sizeof(*w));
    w->model = *m;
    StateMachine *sM = &w->base;
    sM->type = (StateMachineType) m->type;
    sM->stateNumber = m->stateNumber;
    /* state numbering of impl/stateMachine.c:261-263 (five-state: match, shortGapX, shortGapY, longGapX, longGapY) and :625-629 */
    sM->matchState = 0;
    sM->gapXState = 1;
    sM->gapYState = 2;
    sM->startStateProb = start_prob;
    sM->endStateProb = end_prob;
    sM->raggedEndStateProb = ragged_end_prob;
    sM->raggedStartStateProb = ragged_start_prob;
    sM->cellCalculate = cell_calculate_unavailable;
    return sM;
}

const CpbModel *cpecan_model_of(StateMachine *sM) { return &((CpecanStateMachine *) sM)->model; }

StateMachine *stateMachine5_construct(StateMachineType type) {
    if (type != fiveState && type != fiveStateAsymmetric) st_errAbort("stateMachine5_construct: type %i is not a five-state type\n", (int) type);
    CpbModel m;
    if (cpb_model_default((int) type, &m) != CPB_OK) st_errAbort("cpecan: %s", cpb_last_error());
    return wrap(&m);
}

StateMachine *stateMachine3_construct(StateMachineType type) {
    if (type != threeState && type != threeStateAsymmetric) st_errAbort("stateMachine3_construct: type %i is not a three-state type\n", (int) type);
    CpbModel m;
    if (cpb_model_default((int) type, &m) != CPB_OK) st_errAbort("cpecan: %s", cpb_last_error());
    return wrap(&m);
}

/* impl/stateMachine.c:797-819 */
StateMachine *hmm_getStateMachine(Hmm *hmm) {
    CpbModel m;
    if (cpb_model_from_hmm((int) hmm->type, hmm->transitions, hmm->emissions, &m) != CPB_OK) st_errAbort("cpecan: %s", cpb_last_error());
    return wrap(&m);
}

void stateMachine_destruct(StateMachine *stateMachine) { free(stateMachine); }
