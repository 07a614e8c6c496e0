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

#endif
