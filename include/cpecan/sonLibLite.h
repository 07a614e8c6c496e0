/*
 * sonLibLite.h -- the handful of sonLib container entry points that cross cPecan's pair-HMM API.
 *
 * cPecan's public functions take and return sonLib containers (stList of stIntTuple; inc/pairwiseAligner.h:56-108
 * of the reference).  sonLib itself is not part of cPecan's tree (include.mk:2 points at a sibling checkout), so
 * libcpecan.so ships its own implementation of exactly the functions a caller of the hot path touches, under the
 * names sonLib gives them.  A build against a real sonLib defines CPECAN_USE_SONLIB and includes "sonLib.h"
 * instead; the containers are then sonLib's and host/containers.c is left out of the link.
 */
#ifndef CPECAN_SONLIB_LITE_H_
#define CPECAN_SONLIB_LITE_H_

#ifdef CPECAN_USE_SONLIB
#include "sonLib.h"
#else

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct _stList stList;
typedef struct _stIntTuple stIntTuple;

/* lists of owned or borrowed pointers */
stList *stList_construct(void);
stList *stList_construct3(int64_t size, void (*destructElement)(void *));
void stList_destruct(stList *list);
void stList_setDestructor(stList *list, void (*destructElement)(void *));
int64_t stList_length(stList *list);
void *stList_get(stList *list, int64_t index);
void stList_set(stList *list, int64_t index, void *item);
void stList_append(stList *list, void *item);
void stList_appendAll(stList *list, stList *other);
void *stList_pop(stList *list);
void stList_reverse(stList *list);
void stList_sort(stList *list, int (*cmpFn)(const void *a, const void *b));

/* immutable tuples of int64 */
stIntTuple *stIntTuple_construct2(int64_t a, int64_t b);
stIntTuple *stIntTuple_construct3(int64_t a, int64_t b, int64_t c);
stIntTuple *stIntTuple_construct4(int64_t a, int64_t b, int64_t c, int64_t d);
void stIntTuple_destruct(stIntTuple *tuple);
int64_t stIntTuple_length(stIntTuple *tuple);
int64_t stIntTuple_get(stIntTuple *tuple, int64_t index);
int stIntTuple_cmpFn(const void *a, const void *b);
int stIntTuple_equalsFn(const void *a, const void *b);

/* sonLib's error convention: print and exit(1) */
void st_errAbort(const char *format, ...);

#ifdef __cplusplus
}
#endif
#endif /* CPECAN_USE_SONLIB */
#endif /* CPECAN_SONLIB_LITE_H_ */
