/*
 * oracle/shim/shim.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Implementation of the handful of sonLib symbols declared in ./sonLib.h so the
 * reference's pair-HMM sources can be compiled unmodified into oracle/_ref/.
 * Containers, logging and string helpers only: no arithmetic of the path lives here.
 */
#include <errno.h>
#include <stdarg.h>
#include <ctype.h>

#include "sonLib.h"
#include "bioioC.h"

/* ---------------- memory / errors ---------------- */
void *st_malloc(size_t size) {
    void *p = malloc(size ? size : 1);
    if (p == NULL) {
        fprintf(stderr, "oracle shim: malloc of %zu failed\n", size);
        abort();
    }
    return p;
}

void *st_calloc(size_t n, size_t size) {
    void *p = calloc(n ? n : 1, size ? size : 1);
    if (p == NULL) {
        fprintf(stderr, "oracle shim: calloc failed\n");
        abort();
    }
    return p;
}

void st_errAbort(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vfprintf(stderr, fmt, ap);
    va_end(ap);
    fputc('\n', stderr);
    exit(1);
}

void st_errnoAbort(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vfprintf(stderr, fmt, ap);
    va_end(ap);
    fprintf(stderr, " (errno %d)\n", errno);
    exit(1);
}

static int shimLogLevel = 0;
void st_logDebug(const char *fmt, ...) {
    if (shimLogLevel < 2) return;
    va_list ap;
    va_start(ap, fmt);
    vfprintf(stderr, fmt, ap);
    va_end(ap);
}
void st_logInfo(const char *fmt, ...) {
    if (shimLogLevel < 1) return;
    va_list ap;
    va_start(ap, fmt);
    vfprintf(stderr, fmt, ap);
    va_end(ap);
}
void st_logCritical(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vfprintf(stderr, fmt, ap);
    va_end(ap);
}

int64_t st_system(const char *fmt, ...) {
    char buf[4096];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    return system(buf);
}

void stThrowNew(const char *id, const char *fmt, ...) {
    va_list ap;
    fprintf(stderr, "oracle shim: uncaught exception %s: ", id);
    va_start(ap, fmt);
    vfprintf(stderr, fmt, ap);
    va_end(ap);
    fputc('\n', stderr);
    abort();
}

/* ---------------- random (splitmix64; seedable so fixtures are reproducible) ---------------- */
static uint64_t rngState = 0x9E3779B97F4A7C15ULL;
void st_randomSeed(uint64_t seed) { rngState = seed; }
static uint64_t nextU64(void) {
    uint64_t z = (rngState += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
double st_random(void) { return (double) (nextU64() >> 11) * (1.0 / 9007199254740992.0); }
int64_t st_randomInt(int64_t min, int64_t maxPlusOne) {
    assert(maxPlusOne > min);
    return min + (int64_t) (nextU64() % (uint64_t) (maxPlusOne - min));
}

/* ---------------- stList ---------------- */
struct _stList {
    void **items;
    int64_t length, capacity;
    void (*destructElement)(void *);
};

stList *stList_construct3(int64_t size, void (*destructElement)(void *)) {
    stList *l = st_malloc(sizeof(stList));
    l->capacity = size > 8 ? size : 8;
    l->items = st_calloc(l->capacity, sizeof(void *));
    l->length = size;
    l->destructElement = destructElement;
    return l;
}
stList *stList_construct(void) { return stList_construct3(0, NULL); }
void stList_destruct(stList *l) {
    if (l->destructElement != NULL) {
        for (int64_t i = 0; i < l->length; i++) {
            if (l->items[i] != NULL) l->destructElement(l->items[i]);
        }
    }
    free(l->items);
    free(l);
}
int64_t stList_length(stList *l) { return l->length; }
void *stList_get(stList *l, int64_t i) {
    assert(i >= 0 && i < l->length);
    return l->items[i];
}
void stList_set(stList *l, int64_t i, void *item) {
    assert(i >= 0 && i < l->length);
    l->items[i] = item;
}
void stList_append(stList *l, void *item) {
    if (l->length == l->capacity) {
        l->capacity *= 2;
        l->items = realloc(l->items, sizeof(void *) * l->capacity);
        if (l->items == NULL) abort();
    }
    l->items[l->length++] = item;
}
void stList_appendAll(stList *to, stList *from) {
    for (int64_t i = 0; i < from->length; i++) stList_append(to, from->items[i]);
}
void *stList_pop(stList *l) {
    assert(l->length > 0);
    return l->items[--l->length];
}
void stList_reverse(stList *l) {
    for (int64_t i = 0, j = l->length - 1; i < j; i++, j--) {
        void *t = l->items[i];
        l->items[i] = l->items[j];
        l->items[j] = t;
    }
}
static int (*sortCmp)(const void *, const void *);
static int sortTrampoline(const void *a, const void *b) { return sortCmp(*(void *const *) a, *(void *const *) b); }
void stList_sort(stList *l, int (*cmpFn)(const void *, const void *)) {
    sortCmp = cmpFn;
    qsort(l->items, l->length, sizeof(void *), sortTrampoline);
}
void stList_setDestructor(stList *l, void (*d)(void *)) { l->destructElement = d; }

/* ---------------- stIntTuple ---------------- */
struct _stIntTuple {
    int64_t n;
    int64_t v[4];
};
static stIntTuple *tupleN(int64_t n, int64_t a, int64_t b, int64_t c, int64_t d) {
    stIntTuple *t = st_malloc(sizeof(stIntTuple));
    t->n = n;
    t->v[0] = a;
    t->v[1] = b;
    t->v[2] = c;
    t->v[3] = d;
    return t;
}
stIntTuple *stIntTuple_construct2(int64_t a, int64_t b) { return tupleN(2, a, b, 0, 0); }
stIntTuple *stIntTuple_construct3(int64_t a, int64_t b, int64_t c) { return tupleN(3, a, b, c, 0); }
stIntTuple *stIntTuple_construct4(int64_t a, int64_t b, int64_t c, int64_t d) { return tupleN(4, a, b, c, d); }
void stIntTuple_destruct(stIntTuple *t) { free(t); }
int64_t stIntTuple_get(stIntTuple *t, int64_t i) {
    assert(i >= 0 && i < t->n);
    return t->v[i];
}
int64_t stIntTuple_length(stIntTuple *t) { return t->n; }
int stIntTuple_cmpFn(const void *a, const void *b) {
    const stIntTuple *x = a, *y = b;
    int64_t n = x->n < y->n ? x->n : y->n;
    for (int64_t i = 0; i < n; i++) {
        if (x->v[i] != y->v[i]) return x->v[i] < y->v[i] ? -1 : 1;
    }
    return x->n == y->n ? 0 : (x->n < y->n ? -1 : 1);
}
int stIntTuple_equalsFn(const void *a, const void *b) { return stIntTuple_cmpFn(a, b) == 0; }

/* ---------------- stSortedSet (linear; off the hot path) ---------------- */
struct _stSortedSet {
    stList *items;
    int (*cmpFn)(const void *, const void *);
};
stSortedSet *stSortedSet_construct3(int (*cmpFn)(const void *, const void *), void (*destructElement)(void *)) {
    stSortedSet *s = st_malloc(sizeof(stSortedSet));
    s->items = stList_construct3(0, destructElement);
    s->cmpFn = cmpFn;
    return s;
}
void *stSortedSet_search(stSortedSet *s, void *item) {
    for (int64_t i = 0; i < stList_length(s->items); i++) {
        if (s->cmpFn(stList_get(s->items, i), item) == 0) return stList_get(s->items, i);
    }
    return NULL;
}
void stSortedSet_insert(stSortedSet *s, void *item) {
    if (stSortedSet_search(s, item) == NULL) stList_append(s->items, item);
}
void stSortedSet_destruct(stSortedSet *s) {
    stList_destruct(s->items);
    free(s);
}

/* ---------------- strings / files ---------------- */
char *stString_copy(const char *s) {
    char *c = st_malloc(strlen(s) + 1);
    strcpy(c, s);
    return c;
}
char *stString_print(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    int n = vsnprintf(NULL, 0, fmt, ap);
    va_end(ap);
    char *buf = st_malloc((size_t) n + 1);
    va_start(ap, fmt);
    vsnprintf(buf, (size_t) n + 1, fmt, ap);
    va_end(ap);
    return buf;
}
char *stString_getSubString(const char *s, int64_t start, int64_t length) {
    char *c = st_malloc((size_t) length + 1);
    memcpy(c, s + start, (size_t) length);
    c[length] = '\0';
    return c;
}
char *stString_replace(const char *original, const char *toReplace, const char *replacement) {
    size_t lo = strlen(original), lt = strlen(toReplace), lr = strlen(replacement);
    assert(lt > 0);
    size_t cap = lo + 1, len = 0;
    char *out = st_malloc(cap);
    const char *p = original;
    while (*p != '\0') {
        const char *hit = strstr(p, toReplace);
        size_t chunk = hit == NULL ? strlen(p) : (size_t) (hit - p);
        size_t need = len + chunk + (hit ? lr : 0) + 1;
        if (need > cap) {
            while (cap < need) cap *= 2;
            out = realloc(out, cap);
            if (out == NULL) abort();
        }
        memcpy(out + len, p, chunk);
        len += chunk;
        if (hit == NULL) break;
        memcpy(out + len, replacement, lr);
        len += lr;
        p = hit + lt;
    }
    out[len] = '\0';
    return out;
}
stList *stString_split(const char *s) {
    stList *l = stList_construct3(0, free);
    const char *p = s;
    while (*p != '\0') {
        while (*p != '\0' && isspace((unsigned char) *p)) p++;
        if (*p == '\0') break;
        const char *q = p;
        while (*q != '\0' && !isspace((unsigned char) *q)) q++;
        stList_append(l, stString_getSubString(p, 0, q - p));
        p = q;
    }
    return l;
}
char *stFile_getLineFromFile(FILE *fh) {
    size_t cap = 256, len = 0;
    char *buf = st_malloc(cap);
    int ch;
    while ((ch = fgetc(fh)) != EOF && ch != '\n') {
        if (len + 2 > cap) {
            cap *= 2;
            buf = realloc(buf, cap);
            if (buf == NULL) abort();
        }
        buf[len++] = (char) ch;
    }
    if (ch == EOF && len == 0) {
        free(buf);
        return NULL;
    }
    buf[len] = '\0';
    return buf;
}

/* ---------------- JSON + bioio stubs (never reached by the oracle drivers) ---------------- */
static void unreachable(const char *what) {
    fprintf(stderr, "oracle shim: %s is not implemented (out of the hot path)\n", what);
    abort();
}
int64_t stJson_setupParser(char *buf, size_t r, jsmntok_t **tokens, char **js) {
    (void) buf; (void) r; (void) tokens; (void) js;
    unreachable("stJson_setupParser");
    return 0;
}
char *stJson_token_tostr(char *js, jsmntok_t *t) {
    (void) js; (void) t;
    unreachable("stJson_token_tostr");
    return NULL;
}
double stJson_parseFloat(char *js, jsmntok_t *tokens, int64_t i) {
    (void) js; (void) tokens; (void) i;
    unreachable("stJson_parseFloat");
    return 0;
}
int64_t stJson_parseInt(char *js, jsmntok_t *tokens, int64_t i) {
    (void) js; (void) tokens; (void) i;
    unreachable("stJson_parseInt");
    return 0;
}
bool stJson_parseBool(char *js, jsmntok_t *tokens, int64_t i) {
    (void) js; (void) tokens; (void) i;
    unreachable("stJson_parseBool");
    return 0;
}
int64_t stJson_parseFloatArray(double *out, int64_t n, char *js, jsmntok_t *tokens, int64_t i) {
    (void) out; (void) n; (void) js; (void) tokens; (void) i;
    unreachable("stJson_parseFloatArray");
    return 0;
}
char *getTempFile(void) {
    unreachable("getTempFile");
    return NULL;
}
void fastaWrite(char *sequence, char *header, FILE *file) {
    (void) sequence; (void) header; (void) file;
    unreachable("fastaWrite");
}
struct PairwiseAlignment *cigarRead(FILE *fileHandle) {
    (void) fileHandle;
    unreachable("cigarRead");
    return NULL;
}
void destructPairwiseAlignment(struct PairwiseAlignment *pA) {
    (void) pA;
    unreachable("destructPairwiseAlignment");
}
