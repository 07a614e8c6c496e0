/*
 * containers.c -- own implementation of the sonLib list / int-tuple calls that cross cPecan's API
 * (include/cpecan/sonLibLite.h).  Left out of the link when building against a real sonLib.
 */
#ifndef CPECAN_USE_SONLIB
#define _GNU_SOURCE /* qsort_r */
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "cpecan/sonLibLite.h"

struct _stList {
    void **items;
    int64_t length, capacity;
    void (*destructElement)(void *);
    void *slab; /* not NULL: every item is a tuple cut from this slab (cpecan_tripleList_construct) and the list still holds exactly those
                 * tuples, possibly reordered -- stList_destruct then gives them back with one subtraction instead of one call each.
                 * Cleared by every operation that adds, replaces or takes out an item. */
};

/*
 * A tuple is its length followed by its values.  Tuples the library returns by the million (one per aligned pair) do not come from
 * one malloc each: a whole list's tuples are cut from one slab (cpecan_tripleList_construct).  Such a tuple carries, above the low
 * byte of `length`, its distance from the start of the slab, so that stIntTuple_destruct -- which callers are entitled to call on any
 * single tuple, in any order -- finds the slab's count of live tuples and frees the slab with the last one.
 */
struct _stIntTuple {
    int64_t length; /* bits 0..7 the length; bit 62 set: cut from a slab, bits 8..61 = byte offset of this tuple in the slab */
    int64_t values[];
};
#define TUPLE_IN_SLAB ((int64_t) 1 << 62)
#define TUPLE_LENGTH(t) ((t)->length & 0xFF)
typedef struct {
    int64_t live;     /* tuples not yet destructed */
    int64_t reserved; /* keeps the tuples 16-byte aligned */
} TupleSlab;

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
    l->slab = NULL;
    return l;
}

stList *stList_construct(void) { return stList_construct3(0, NULL); }

static void slab_release(void *slab, int64_t n);

void stList_destruct(stList *l) {
    if (l == NULL) return;
    if (l->slab != NULL && l->destructElement == (void (*)(void *)) stIntTuple_destruct) {
        slab_release(l->slab, l->length); /* an untouched result list: its tuples go back together */
    } else if (l->destructElement != NULL) {
        for (int64_t i = 0; i < l->length; i++) {
            if (l->items[i] != NULL) l->destructElement(l->items[i]);
        }
    }
    free(l->items);
    free(l);
}

void stList_setDestructor(stList *l, void (*destructElement)(void *)) { l->destructElement = destructElement; } /* stList_destruct checks it against the mark */
int64_t stList_length(stList *l) { return l == NULL ? 0 : l->length; }

void *stList_get(stList *l, int64_t i) {
    if (i < 0 || i >= l->length) st_errAbort("stList_get: index %lld out of range (length %lld)", (long long) i, (long long) l->length);
    return l->items[i];
}

void stList_set(stList *l, int64_t i, void *item) {
    if (i < 0 || i >= l->length) st_errAbort("stList_set: index %lld out of range (length %lld)", (long long) i, (long long) l->length);
    l->slab = NULL;
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
    l->slab = NULL;
    reserve(l, l->length + 1);
    l->items[l->length++] = item;
}

void stList_appendAll(stList *l, stList *other) {
    l->slab = NULL;
    reserve(l, l->length + other->length);
    memcpy(l->items + l->length, other->items, (size_t) other->length * sizeof(void *));
    l->length += other->length;
}

void *stList_pop(stList *l) {
    if (l->length == 0) st_errAbort("stList_pop: empty list");
    l->slab = NULL;
    return l->items[--l->length];
}

void stList_reverse(stList *l) {
    for (int64_t i = 0, j = l->length - 1; i < j; i++, j--) {
        void *t = l->items[i];
        l->items[i] = l->items[j];
        l->items[j] = t;
    }
}

static int cmp_indirect(const void *a, const void *b, void *cmpFn) {
    return ((int (*)(const void *, const void *)) cmpFn)(*(void *const *) a, *(void *const *) b);
}

void stList_sort(stList *l, int (*cmpFn)(const void *a, const void *b)) {
    qsort_r(l->items, (size_t) l->length, sizeof(void *), cmp_indirect, (void *) cmpFn); /* no global: lists are sorted on several host threads */
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

void stIntTuple_destruct(stIntTuple *t) {
    if (t == NULL) return;
    if (t->length & TUPLE_IN_SLAB) {
        TupleSlab *slab = (TupleSlab *) ((char *) t - ((t->length & ~TUPLE_IN_SLAB) >> 8));
        if (--slab->live == 0) free(slab);
        return;
    }
    free(t);
}
int64_t stIntTuple_length(stIntTuple *t) { return TUPLE_LENGTH(t); }

int64_t stIntTuple_get(stIntTuple *t, int64_t i) {
    if (i < 0 || i >= TUPLE_LENGTH(t)) st_errAbort("stIntTuple_get: index %lld out of range (length %lld)", (long long) i, (long long) TUPLE_LENGTH(t));
    return t->values[i];
}

int stIntTuple_cmpFn(const void *a, const void *b) {
    const stIntTuple *x = a, *y = b;
    const int64_t lx = TUPLE_LENGTH(x), ly = TUPLE_LENGTH(y), n = lx < ly ? lx : ly;
    for (int64_t i = 0; i < n; i++) {
        if (x->values[i] != y->values[i]) return x->values[i] < y->values[i] ? -1 : 1;
    }
    return lx == ly ? 0 : (lx < ly ? -1 : 1);
}

/* n (pInt, x, y) triples as a list of 3-tuples with destructor stIntTuple_destruct: two allocations instead of n + 2 */
stList *cpecan_tripleList_construct(const int32_t *triples, int64_t n) {
    stList *l = stList_construct3(0, (void (*)(void *)) stIntTuple_destruct);
    if (n <= 0) return l;
    const size_t stride = sizeof(stIntTuple) + 3 * sizeof(int64_t); /* 32 bytes */
    TupleSlab *slab = xmalloc(sizeof(TupleSlab) + (size_t) n * stride);
    slab->live = n;
    slab->reserved = 0;
    reserve(l, n);
    char *at = (char *) (slab + 1);
    for (int64_t i = 0; i < n; i++, at += stride) {
        stIntTuple *t = (stIntTuple *) at;
        t->length = 3 | TUPLE_IN_SLAB | ((int64_t) (at - (char *) slab) << 8);
        t->values[0] = triples[3 * i];
        t->values[1] = triples[3 * i + 1];
        t->values[2] = triples[3 * i + 2];
        l->items[i] = t;
    }
    l->length = n;
    l->slab = slab;
    return l;
}

static void slab_release(void *slab, int64_t n) {
    TupleSlab *s = slab;
    s->live -= n;
    if (s->live == 0) free(s);
}

int stIntTuple_equalsFn(const void *a, const void *b) { return stIntTuple_cmpFn(a, b) == 0; }

#else /* CPECAN_USE_SONLIB: sonLib owns the tuple layout, so the triples become tuples one allocation at a time */
#include "sonLib.h"
stList *cpecan_tripleList_construct(const int32_t *triples, int64_t n) {
    stList *l = stList_construct3(0, (void (*)(void *)) stIntTuple_destruct);
    for (int64_t i = 0; i < n; i++) stList_append(l, stIntTuple_construct3(triples[3 * i], triples[3 * i + 1], triples[3 * i + 2]));
    return l;
}
#endif /* !CPECAN_USE_SONLIB */
