/* host_internal.h -- private to the C host layer of libcpecan.so */
#ifndef CPECAN_HOST_INTERNAL_H_
#define CPECAN_HOST_INTERNAL_H_

#include <stddef.h>

#include "cpecan/stateMachine.h"
#include "cpecan_b200.h"

/* a StateMachine as this library allocates it: the public struct first (callers only see that), then the flat device image */
typedef struct {
    StateMachine base;
    CpbModel model;
} CpecanStateMachine;

const CpbModel *cpecan_model_of(StateMachine *sM);
void *cpecan_malloc(size_t bytes); /* aborts on failure, like st_malloc */

/* containers.c: n (pInt, x, y) triples as a list of 3-tuples (destructor stIntTuple_destruct), the tuples cut from one slab */
stList *cpecan_tripleList_construct(const int32_t *triples, int64_t n);

/* pairwiseAligner.c: fn(first, last, arg) over [0, n) on up to $CPECAN_HOST_THREADS (default: the online CPUs, at most 32) threads;
 * the ranges are contiguous and balanced by weight[i+1] - weight[i] (a prefix-sum array of n + 1 entries) when weight is not NULL */
void cpecan_parallel_for(int64_t n, const int64_t *weight, void (*fn)(int64_t first, int64_t last, void *arg), void *arg);

#endif
