/*
 * oracle/shim/sonLib.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Minimal stand-in for the (un-vendored) sonLib headers, just enough for the
 * reference's own impl/pairwiseAligner.c, impl/stateMachine.c and
 * impl/randomSequences.c to compile *unmodified* from /root/reference into
 * oracle/_ref/ (see oracle/Makefile).  sonLib only contributes containers,
 * logging and string helpers to the pair-HMM path; all arithmetic lives in the
 * reference's own two files + libm, so this shim cannot change results.
 *
 * Nothing in the product (cpecan_b200/, include/) includes this file.
 */
#ifndef ORACLE_SHIM_SONLIB_H_
#define ORACLE_SHIM_SONLIB_H_

#include <assert.h>
#include <inttypes.h>
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifndef TRUE
#define TRUE 1
#endif
#ifndef FALSE
#define FALSE 0
#endif

#define LOG_ONE 0.0

/* ---- memory / errors / logging ---- */
void *st_malloc(size_t size);
void *st_calloc(size_t n, size_t size);
void st_errAbort(const char *fmt, ...);
void st_errnoAbort(const char *fmt, ...);
void st_logDebug(const char *fmt, ...);
void st_logInfo(const char *fmt, ...);
void st_logCritical(const char *fmt, ...);
int64_t st_system(const char *fmt, ...);

/* exceptions: the only throw on the path is an invalid Diagonal; the shim aborts */
void stThrowNew(const char *id, const char *fmt, ...);

/* ---- random ---- */
double st_random(void);
int64_t st_randomInt(int64_t min, int64_t maxPlusOne);
void st_randomSeed(uint64_t seed);

/* ---- stList ---- */
typedef struct _stList stList;
stList *stList_construct(void);
stList *stList_construct3(int64_t size, void (*destructElement)(void *));
void stList_destruct(stList *list);
int64_t stList_length(stList *list);
void *stList_get(stList *list, int64_t index);
void stList_set(stList *list, int64_t index, void *item);
void stList_append(stList *list, void *item);
void stList_appendAll(stList *to, stList *from);
void *stList_pop(stList *list);
void stList_reverse(stList *list);
void stList_sort(stList *list, int (*cmpFn)(const void *a, const void *b));
void stList_setDestructor(stList *list, void (*destructElement)(void *));

/* ---- stIntTuple ---- */
typedef struct _stIntTuple stIntTuple;
stIntTuple *stIntTuple_construct2(int64_t a, int64_t b);
stIntTuple *stIntTuple_construct3(int64_t a, int64_t b, int64_t c);
stIntTuple *stIntTuple_construct4(int64_t a, int64_t b, int64_t c, int64_t d);
void stIntTuple_destruct(stIntTuple *t);
int64_t stIntTuple_get(stIntTuple *t, int64_t index);
int64_t stIntTuple_length(stIntTuple *t);
int stIntTuple_cmpFn(const void *a, const void *b);
int stIntTuple_equalsFn(const void *a, const void *b);

/* ---- stSortedSet (only used by filterToRemoveOverlap, off the hot path) ---- */
typedef struct _stSortedSet stSortedSet;
stSortedSet *stSortedSet_construct3(int (*cmpFn)(const void *, const void *), void (*destructElement)(void *));
void stSortedSet_insert(stSortedSet *set, void *item);
void *stSortedSet_search(stSortedSet *set, void *item);
void stSortedSet_destruct(stSortedSet *set);

/* ---- strings / files ---- */
char *stString_copy(const char *s);
char *stString_print(const char *fmt, ...);
char *stString_getSubString(const char *s, int64_t start, int64_t length);
char *stString_replace(const char *original, const char *toReplace, const char *replacement);
stList *stString_split(const char *s);
char *stFile_getLineFromFile(FILE *fh);

/* ---- JSON (jsmn) : not needed by the oracle; stubs abort ---- */
typedef struct {
    int type, start, end, size;
} jsmntok_t;
int64_t stJson_setupParser(char *buf, size_t r, jsmntok_t **tokens, char **js);
char *stJson_token_tostr(char *js, jsmntok_t *t);
double stJson_parseFloat(char *js, jsmntok_t *tokens, int64_t tokenIndex);
int64_t stJson_parseInt(char *js, jsmntok_t *tokens, int64_t tokenIndex);
bool stJson_parseBool(char *js, jsmntok_t *tokens, int64_t tokenIndex);
int64_t stJson_parseFloatArray(double *out, int64_t n, char *js, jsmntok_t *tokens, int64_t tokenIndex);

#endif
