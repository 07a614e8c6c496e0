/*
 * minijson.h -- the small part of JSON the model and parameter files use: one flat object whose values are
 * numbers, booleans, strings or arrays of numbers (the reference parses the same files with sonLib's stJson_*
 * wrappers over jsmn, impl/pairwiseAligner.c:1354-1409 and impl/stateMachine.c:204-253).
 */
#ifndef CPECAN_MINIJSON_H_
#define CPECAN_MINIJSON_H_

#include <stddef.h>
#include <stdint.h>

typedef struct {
    const char *p, *end;
    char error[128];
} MiniJson;

/* callbacks receive the key (NUL-terminated copy) and the parser positioned at the value */
typedef int (*MiniJsonMember)(MiniJson *j, const char *key, void *extra);

void minijson_init(MiniJson *j, const char *buf, size_t len);
/* parses { "key": value, ... } calling member() per key; returns 0 on success, -1 on error (message in j->error) */
int minijson_object(MiniJson *j, MiniJsonMember member, void *extra);
int minijson_number(MiniJson *j, double *out);
/* accepts true / false and, as jsmn-based parsers do, the numbers 0 / 1 */
int minijson_bool(MiniJson *j, int *out);
/* array of exactly n numbers */
int minijson_number_array(MiniJson *j, double *out, int64_t n);
int minijson_skip_value(MiniJson *j);

#endif
