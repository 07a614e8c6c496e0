/*
 * realignJobs.h -- what cPecanRealign and cPecanEm share: progress logging, the table of input sequences, and the preparation of
 * one input alignment (a cigar) for the device pass: sub-sequences on the forward strand, anchors from the cigar's match columns
 * (cPecanRealign.c:242-274, :220-240, :509-528 of the reference).  Static functions: each program is one translation unit.
 */
#ifndef CPECAN_REALIGN_JOBS_H_
#define CPECAN_REALIGN_JOBS_H_

#include <ctype.h>
#include <inttypes.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "cpecan/pairwiseAligner.h"
#include "cpecan/pairwiseAlignment.h"

/* ---- logging: -a INFO / DEBUG prints progress to stderr, as sonLib's st_logInfo would ---- */

static int logLevel = 0; /* 0 off, 1 info, 2 debug */

static void set_log_level(const char *s) {
    if (s == NULL) return;
    char u[16];
    size_t n = 0;
    for (; s[n] != '\0' && n + 1 < sizeof(u); n++) u[n] = (char) toupper((unsigned char) s[n]);
    u[n] = '\0';
    if (strcmp(u, "INFO") == 0) logLevel = 1;
    else if (strcmp(u, "DEBUG") == 0) logLevel = 2;
    else logLevel = 0;
}

static void log_info(const char *format, ...) {
    if (logLevel < 1) return;
    va_list ap;
    va_start(ap, format);
    vfprintf(stderr, format, ap);
    va_end(ap);
}

static void *xmalloc(size_t bytes) {
    void *p = malloc(bytes ? bytes : 1);
    if (p == NULL) st_errAbort("out of memory allocating %zu bytes", bytes);
    return p;
}

/* ---- sequences by the first word of their FASTA header (cPecanRealign.c:242-274) ---- */

typedef struct {
    char *name, *seq;
    int64_t length;
} NamedSeq;
static NamedSeq *seqs = NULL; /* open-addressing table, capSeqs a power of two, name == NULL: free slot */
static int64_t nSeqs = 0, capSeqs = 0;

static uint64_t name_hash(const char *s) {
    uint64_t h = 1469598103934665603ULL; /* FNV-1a */
    for (; *s != '\0'; s++) h = (h ^ (unsigned char) *s) * 1099511628211ULL;
    return h;
}

/* the slot holding `name`, or the free slot where it would go */
static NamedSeq *sequence_slot(NamedSeq *table, int64_t cap, const char *name) {
    for (uint64_t i = name_hash(name) & (uint64_t) (cap - 1);; i = (i + 1) & (uint64_t) (cap - 1)) {
        if (table[i].name == NULL || strcmp(table[i].name, name) == 0) return &table[i];
    }
}

static NamedSeq *find_sequence(const char *name) {
    if (capSeqs == 0) return NULL;
    NamedSeq *slot = sequence_slot(seqs, capSeqs, name);
    return slot->name != NULL ? slot : NULL;
}

static void add_sequence(const char *header, const char *sequence, int64_t length) {
    size_t n = 0;
    while (header[n] != '\0' && !isspace((unsigned char) header[n])) n++;
    char *name = stString_getSubString(header, 0, (int64_t) n);
    NamedSeq *old = find_sequence(name);
    if (old != NULL) {
        log_info("Got a repeat header: %s with sequence length: %" PRIi64 " vs. the existing sequence of length: %" PRIi64 ", complete header: %s\n", name,
                 length, old->length, header);
        if (length > old->length) { /* a more complete version of the same sequence (overlapping fragments) */
            log_info("Replacing sequence\n");
            free(old->seq);
            old->seq = stString_copy(sequence);
            old->length = length;
        }
        free(name);
        return;
    }
    log_info("Adding sequence for header: %s, with length %" PRIi64 ", complete header: %s\n", name, length, header);
    if (2 * (nSeqs + 1) > capSeqs) {
        const int64_t cap = capSeqs ? 2 * capSeqs : 1024;
        NamedSeq *table = xmalloc((size_t) cap * sizeof(NamedSeq));
        memset(table, 0, (size_t) cap * sizeof(NamedSeq));
        for (int64_t i = 0; i < capSeqs; i++) {
            if (seqs[i].name != NULL) *sequence_slot(table, cap, seqs[i].name) = seqs[i];
        }
        free(seqs);
        seqs = table;
        capSeqs = cap;
    }
    NamedSeq *slot = sequence_slot(seqs, capSeqs, name);
    slot->name = name;
    slot->seq = stString_copy(sequence);
    slot->length = length;
    nSeqs++;
}

/* ---- coordinates (cPecanRealign.c:220-240, :292-298) ---- */

static void rebase(int64_t *start, int64_t *end, int64_t *strand, int64_t shift, bool flip) {
    *start += shift;
    *end += shift;
    if (flip) {
        *strand = *strand ? 0 : 1;
        const int64_t t = *end;
        *end = *start;
        *start = t;
    }
}

static char *sub_sequence(const char *seq, int64_t start, int64_t end, bool strand) {
    if (strand) return stString_getSubString(seq, start, end - start);
    char *fwd = stString_getSubString(seq, end, start - end);
    char *rc = stString_reverseComplementString(fwd);
    free(fwd);
    return rc;
}

/* ---- one input alignment, prepared for the device pass ---- */

typedef struct {
    struct PairwiseAlignment *pA; /* rebased to the forward strand, starting at 0 */
    char *subX, *subY;
    bool flip1, flip2;
    int64_t shift1, shift2;
    stList *anchors;  /* every column of the input alignment (after trim) */
    stList *filtered; /* those whose two bases agree (and are not N): what constrains the band */
} Job;

static bool columns_match(const Job *j, stIntTuple *t) {
    const int x = toupper((unsigned char) j->subX[stIntTuple_get(t, 0)]), y = toupper((unsigned char) j->subY[stIntTuple_get(t, 1)]);
    return x == y && x != 'N';
}

static void job_prepare(Job *j, struct PairwiseAlignment *pA, const PairwiseAlignmentParameters *p) {
    log_info("Processing alignment for sequences: %s and %s\n", pA->contig1, pA->contig2);
    const NamedSeq *sX = find_sequence(pA->contig1), *sY = find_sequence(pA->contig2);
    if (sX == NULL || sY == NULL) st_errAbort("no sequence named %s in the input files", sX == NULL ? pA->contig1 : pA->contig2);
    const int64_t hi1 = pA->strand1 ? pA->end1 : pA->start1, hi2 = pA->strand2 ? pA->end2 : pA->start2;
    if (hi1 > sX->length || hi2 > sY->length || pA->start1 < 0 || pA->end1 < 0 || pA->start2 < 0 || pA->end2 < 0) {
        st_errAbort("alignment of %s and %s reaches beyond the end of a sequence", pA->contig1, pA->contig2);
    }
    j->pA = pA;
    j->flip1 = !pA->strand1;
    j->flip2 = !pA->strand2;
    j->shift1 = pA->strand1 ? pA->start1 : pA->end1;
    j->shift2 = pA->strand2 ? pA->start2 : pA->end2;
    j->subX = sub_sequence(sX->seq, pA->start1, pA->end1, pA->strand1 != 0);
    j->subY = sub_sequence(sY->seq, pA->start2, pA->end2, pA->strand2 != 0);
    rebase(&pA->start1, &pA->end1, &pA->strand1, -j->shift1, j->flip1);
    rebase(&pA->start2, &pA->end2, &pA->strand2, -j->shift2, j->flip2);
    checkPairwiseAlignment(pA);
    j->anchors = convertPairwiseForwardStrandAlignmentToAnchorPairs(pA, p->constraintDiagonalTrim, p->diagonalExpansion);
    j->filtered = stList_construct(); /* borrows the tuples of j->anchors */
    for (int64_t i = 0; i < stList_length(j->anchors); i++) {
        stIntTuple *t = stList_get(j->anchors, i);
        if (columns_match(j, t)) stList_append(j->filtered, t);
    }
}

static void job_release(Job *j) {
    destructPairwiseAlignment(j->pA);
    stList_destruct(j->filtered);
    stList_destruct(j->anchors);
    free(j->subX);
    free(j->subY);
}

static int64_t parse_int(const char *arg, const char *what) {
    char *end;
    const long long v = strtoll(arg, &end, 10);
    if (end == arg || *end != '\0' || v < 0) st_errAbort("%s needs a non-negative integer, got '%s'", what, arg);
    return (int64_t) v;
}

static float parse_float(const char *arg, const char *what) {
    char *end;
    const float v = strtof(arg, &end);
    if (end == arg || *end != '\0' || !(v >= 0.0f)) st_errAbort("%s needs a non-negative number, got '%s'", what, arg);
    return v;
}

static void read_sequence_files(int argc, char *argv[], int first) {
    for (int i = first; i < argc; i++) {
        FILE *f = fopen(argv[i], "r");
        if (f == NULL) st_errAbort("cannot read %s", argv[i]);
        fastaReadToFunction(f, add_sequence);
        fclose(f);
    }
}

static void free_sequences(void) {
    for (int64_t i = 0; i < capSeqs; i++) {
        free(seqs[i].name);
        free(seqs[i].seq);
    }
    free(seqs);
    seqs = NULL;
    nSeqs = capSeqs = 0;
}

#endif /* CPECAN_REALIGN_JOBS_H_ */
