/*
 * containers.c -- own implementation of the sonLib list / int-tuple calls that cross cPecan's API
 * (include/cpecan/sonLibLite.h).  Left out of the link when building against a real sonLib.
 */
#ifndef CPECAN_USE_SONLIB
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "cpecan/sonLibLite.h"

struct _stList {
    void **items;
    int64_t length, capacity;
    void (*destructElement)(void *);
};

struct _stIntTuple {
    int64_t length;
    int64_t values[];
};

void st_errAbort(const char *format, ...) {
    va_list ap;
    va_start(ap, format);
    vfprintf(stderr, format, ap);
    va_end(ap);
    fputc('\n', stderr);
    exit(1);
}

static void *xmalloc(size_t bytes) {
    void *p = malloc(bytes ? bytes : 1);
    if (p == NULL) st_errAbort("cpecan: out of memory allocating %zu bytes", bytes);
    return p;
}

stList *stList_construct3(int64_t size, void (*destructElement)(void *)) {
    stList *l = xmalloc(sizeof(*l));
    l->capacity = size > 4 ? size : 4;
    l->length = size > 0 ? size : 0;
    l->items = xmalloc((size_t) l->capacity * sizeof(void *));
    memset(l->items, 0, (size_t) l->capacity * sizeof(void *));
    l->destructElement = destructElement;
    return l;
}

stList *stList_construct(void) { return stList_construct3(0, NULL); }

void stList_destruct(stList *l) {
    if (l == NULL) return;
    if (l->destructElement != NULL) {
        for (int64_t i = 0; i < l->length; i++) {
            if (l->items[i] != NULL) l->destructElement(l->items[i]);
        }
    }
    free(l->items);
    free(l);
}

void stList_setDestructor(stList *l, void (*destructElement)(void *)) { l->destructElement = destructElement; }
int64_t stList_length(stList *l) { return l == NULL ? 0 : l->length; }

void *stList_get(stList *l, int64_t i) {
    if (i < 0 || i >= l->length) st_errAbort("stList_get: index %lld out of range (length %lld)", (long long) i, (long long) l->length);
    return l->items[i];
}

void stList_set(stList *l, int64_t i, void *item) {
    if (i < 0 || i >= l->length) st_errAbort("stList_set: index %lld out of range (length %lld)", (long long) i, (long long) l->length);
    l->items[i] = item;
}

static void reserve(stList *l, int64_t n) {
    if (n <= l->capacity) return;
    int64_t cap = l->capacity;
    while (cap < n) cap = cap * 2;
    void **items = realloc(l->items, (size_t) cap * sizeof(void *));
    if (items == NULL) st_errAbort("cpecan: out of memory growing a list to %lld items", (long long) cap);
    l->items = items;
    l->capacity = cap;
}

void stList_append(stList *l, void *item) {
    reserve(l, l->length + 1);
    l->items[l->length++] = item;
}

void stList_appendAll(stList *l, stList *other) {
    reserve(l, l->length + other->length);
    memcpy(l->items + l->length, other->items, (size_t) other->length * sizeof(void *));
    l->length += other->length;
}

void *stList_pop(stList *l) {
    if (l->length == 0) st_errAbort("stList_pop: empty list");
    return l->items[--l->length];
}

void stList_reverse(stList *l) {
    for (int64_t i = 0, j = l->length - 1; i < j; i++, j--) {
        void *t = l->items[i];
        l->items[i] = l->items[j];
        l->items[j] = t;
    }
}

static int (*g_cmp)(const void *, const void *);
static int cmp_indirect(const void *a, const void *b) { return g_cmp(*(void *const *) a, *(void *const *) b); }

void stList_sort(stList *l, int (*cmpFn)(const void *a, const void *b)) {
    g_cmp = cmpFn; /* cPecan's callers are single threaded */
    qsort(l->items, (size_t) l->length, sizeof(void *), cmp_indirect);
}

static stIntTuple *tuple_new(int64_t n) {
    stIntTuple *t = xmalloc(sizeof(*t) + (size_t) n * sizeof(int64_t));
    t->length = n;
    return t;
}

stIntTuple *stIntTuple_construct2(int64_t a, int64_t b) {
    stIntTuple *t = tuple_new(2);
    t->values[0] = a;
    t->values[1] = b;
    return t;
}

stIntTuple *stIntTuple_construct3(int64_t a, int64_t b, int64_t c) {
    stIntTuple *t = tuple_new(3);
    t->values[0] = a;
    t->values[1] = b;
    t->values[2] = c;
    return t;
}

stIntTuple *stIntTuple_construct4(int64_t a, int64_t b, int64_t c, int64_t d) {
    stIntTuple *t = tuple_new(4);
    t->values[0] = a;
    t->values[1] = b;
    t->values[2] = c;
    t->values[3] = d;
    return t;
}

void stIntTuple_destruct(stIntTuple *t) { free(t); }
int64_t stIntTuple_length(stIntTuple *t) { return t->length; }

int64_t stIntTuple_get(stIntTuple *t, int64_t i) {
    if (i < 0 || i >= t->length) st_errAbort("stIntTuple_get: index %lld out of range (length %lld)", (long long) i, (long long) t->length);
    return t->values[i];
}

int stIntTuple_cmpFn(const void *a, const void *b) {
    const stIntTuple *x = a, *y = b;
    const int64_t n = x->length < y->length ? x->length : y->length;
    for (int64_t i = 0; i < n; i++) {
        if (x->values[i] != y->values[i]) return x->values[i] < y->values[i] ? -1 : 1;
    }
    return x->length == y->length ? 0 : (x->length < y->length ? -1 : 1);
}

int stIntTuple_equalsFn(const void *a, const void *b) { return stIntTuple_cmpFn(a, b) == 0; }

#endif /* !CPECAN_USE_SONLIB */
