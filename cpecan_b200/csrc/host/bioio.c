/*
 * bioio.c -- cigar / FASTA I/O and the PairwiseAlignment record of include/cpecan/pairwiseAlignment.h: what
 * cPecanRealign reads and writes around the device pass (cPecanRealign.c:498-523, :591-599 of the reference, where
 * sonLib's bioioC provides them).  Left out of the link when building against a real sonLib.
 */
#ifndef CPECAN_USE_SONLIB
#include <ctype.h>
#include <inttypes.h>
#include <stdlib.h>
#include <string.h>

#include "cpecan/pairwiseAlignment.h"
#include "host_internal.h"

/* ---- strings ---- */

char *stString_copy(const char *s) {
    if (s == NULL) return NULL;
    const size_t n = strlen(s);
    char *c = cpecan_malloc(n + 1);
    memcpy(c, s, n + 1);
    return c;
}

char *stString_getSubString(const char *s, int64_t start, int64_t length) {
    char *c = cpecan_malloc((size_t) length + 1);
    memcpy(c, s + start, (size_t) length);
    c[length] = '\0';
    return c;
}

static char complement(char c) {
    switch (c) {
    case 'A': return 'T';  case 'a': return 't';
    case 'C': return 'G';  case 'c': return 'g';
    case 'G': return 'C';  case 'g': return 'c';
    case 'T': return 'A';  case 't': return 'a';
    /* two- and three-base IUPAC codes; N, S, W and anything else are their own complement */
    case 'R': return 'Y';  case 'r': return 'y';
    case 'Y': return 'R';  case 'y': return 'r';
    case 'K': return 'M';  case 'k': return 'm';
    case 'M': return 'K';  case 'm': return 'k';
    case 'B': return 'V';  case 'b': return 'v';
    case 'V': return 'B';  case 'v': return 'b';
    case 'D': return 'H';  case 'd': return 'h';
    case 'H': return 'D';  case 'h': return 'd';
    default: return c;
    }
}

char *stString_reverseComplementString(const char *s) {
    const size_t n = strlen(s);
    char *c = cpecan_malloc(n + 1);
    for (size_t i = 0; i < n; i++) c[i] = complement(s[n - 1 - i]);
    c[n] = '\0';
    return c;
}

/* ---- struct List ---- */

struct List *constructEmptyList(int64_t length, void (*destructElement)(void *)) {
    struct List *l = cpecan_malloc(sizeof(*l));
    l->length = length;
    l->maxLength = length > 8 ? length : 8;
    l->list = cpecan_malloc((size_t) l->maxLength * sizeof(void *));
    memset(l->list, 0, (size_t) l->maxLength * sizeof(void *));
    l->destructElement = destructElement;
    return l;
}

void listAppend(struct List *l, void *item) {
    if (l->length == l->maxLength) {
        l->maxLength *= 2;
        l->list = realloc(l->list, (size_t) l->maxLength * sizeof(void *));
        if (l->list == NULL) st_errAbort("cpecan: out of memory growing a list to %" PRIi64 " items", l->maxLength);
    }
    l->list[l->length++] = item;
}

void destructList(struct List *l) {
    if (l == NULL) return;
    if (l->destructElement != NULL) {
        for (int64_t i = 0; i < l->length; i++) l->destructElement(l->list[i]);
    }
    free(l->list);
    free(l);
}

/* ---- alignments ---- */

struct AlignmentOperation *constructAlignmentOperation(int64_t opType, int64_t length, double score) {
    struct AlignmentOperation *op = cpecan_malloc(sizeof(*op));
    op->opType = opType;
    op->length = length;
    op->score = score;
    return op;
}

void destructAlignmentOperation(struct AlignmentOperation *op) { free(op); }

struct PairwiseAlignment *constructPairwiseAlignment(const char *contig1, int64_t start1, int64_t end1, int64_t strand1, const char *contig2,
                                                     int64_t start2, int64_t end2, int64_t strand2, double score, struct List *operationList) {
    struct PairwiseAlignment *pA = cpecan_malloc(sizeof(*pA));
    pA->contig1 = stString_copy(contig1);
    pA->start1 = start1;
    pA->end1 = end1;
    pA->strand1 = strand1;
    pA->contig2 = stString_copy(contig2);
    pA->start2 = start2;
    pA->end2 = end2;
    pA->strand2 = strand2;
    pA->score = score;
    pA->operationList = operationList;
    return pA;
}

void destructPairwiseAlignment(struct PairwiseAlignment *pA) {
    if (pA == NULL) return;
    destructList(pA->operationList);
    free(pA->contig1);
    free(pA->contig2);
    free(pA);
}

void checkPairwiseAlignment(struct PairwiseAlignment *pA) {
    int64_t l1 = 0, l2 = 0;
    for (int64_t i = 0; i < pA->operationList->length; i++) {
        const struct AlignmentOperation *op = pA->operationList->list[i];
        if (op->length <= 0) st_errAbort("cigar %s/%s: operation %" PRIi64 " has length %" PRIi64, pA->contig1, pA->contig2, i, op->length);
        if (op->opType != PAIRWISE_INDEL_Y) l1 += op->length;
        if (op->opType != PAIRWISE_INDEL_X) l2 += op->length;
    }
    const int64_t span1 = pA->strand1 ? pA->end1 - pA->start1 : pA->start1 - pA->end1;
    const int64_t span2 = pA->strand2 ? pA->end2 - pA->start2 : pA->start2 - pA->end2;
    if (span1 < 0 || span2 < 0 || span1 != l1 || span2 != l2) {
        st_errAbort("cigar %s [%" PRIi64 ", %" PRIi64 ") %c / %s [%" PRIi64 ", %" PRIi64 ") %c: operations cover %" PRIi64 " and %" PRIi64
                    " bases",
                    pA->contig1, pA->start1, pA->end1, pA->strand1 ? '+' : '-', pA->contig2, pA->start2, pA->end2, pA->strand2 ? '+' : '-', l1,
                    l2);
    }
}

/* ---- lines ---- */

/* next line of the stream without its newline, in a buffer the caller owns and reuses; NULL at end of file */
static char *read_line(FILE *f, char **buf, size_t *cap) {
    const ssize_t got = getline(buf, cap, f); /* a character at a time through fgetc was a third of cPecanRealign's start-up on a 40 MB FASTA file */
    if (got < 0) return NULL;
    size_t n = (size_t) got;
    if (n > 0 && (*buf)[n - 1] == '\n') n--;
    if (n > 0 && (*buf)[n - 1] == '\r') n--;
    (*buf)[n] = '\0';
    return *buf;
}

/* ---- cigars ---- */

static int strand_of(const char *token, const char *line) {
    if (strcmp(token, "+") == 0) return 1;
    if (strcmp(token, "-") == 0) return 0;
    st_errAbort("cigar line has strand '%s': %s", token, line);
    return 0;
}

static int64_t int_of(const char *token, const char *line) {
    char *end;
    const long long v = strtoll(token, &end, 10);
    if (end == token || *end != '\0') st_errAbort("cigar line has '%s' where an integer is expected: %s", token, line);
    return (int64_t) v;
}

struct PairwiseAlignment *cigarRead(FILE *fileHandle) {
    char *line = NULL;
    size_t cap = 0;
    struct PairwiseAlignment *pA = NULL;
    while (pA == NULL && read_line(fileHandle, &line, &cap) != NULL) {
        if (strncmp(line, "cigar:", 6) != 0) continue;
        char *copy = stString_copy(line), *save = NULL;
        char *tok[10];
        int n = 0;
        while (n < 9 && (tok[n] = strtok_r(n == 0 ? copy + 6 : NULL, " \t", &save)) != NULL) n++;
        if (n < 9) st_errAbort("cigar line has %d of the 9 leading fields: %s", n, line);
        char *endp;
        const double score = strtod(tok[8], &endp);
        if (endp == tok[8]) st_errAbort("cigar line has score '%s': %s", tok[8], line);
        struct List *ops = constructEmptyList(0, (void (*)(void *)) destructAlignmentOperation);
        for (char *t = strtok_r(NULL, " \t", &save); t != NULL; t = strtok_r(NULL, " \t", &save)) {
            int64_t type;
            if (strcmp(t, "M") == 0) type = PAIRWISE_MATCH;
            else if (strcmp(t, "D") == 0) type = PAIRWISE_INDEL_X;
            else if (strcmp(t, "I") == 0) type = PAIRWISE_INDEL_Y;
            else {
                st_errAbort("cigar line has operation '%s': %s", t, line);
                return NULL;
            }
            char *len = strtok_r(NULL, " \t", &save);
            if (len == NULL) st_errAbort("cigar line ends after operation '%s': %s", t, line);
            listAppend(ops, constructAlignmentOperation(type, int_of(len, line), 0.0));
        }
        pA = constructPairwiseAlignment(tok[4], int_of(tok[5], line), int_of(tok[6], line), strand_of(tok[7], line), tok[0], int_of(tok[1], line),
                                        int_of(tok[2], line), strand_of(tok[3], line), score, ops);
        free(copy);
    }
    free(line);
    return pA;
}

void cigarWrite(FILE *fileHandle, struct PairwiseAlignment *pA, int64_t withProbs) {
    fprintf(fileHandle, "cigar: %s %" PRIi64 " %" PRIi64 " %c %s %" PRIi64 " %" PRIi64 " %c %f", pA->contig2, pA->start2, pA->end2,
            pA->strand2 ? '+' : '-', pA->contig1, pA->start1, pA->end1, pA->strand1 ? '+' : '-', pA->score);
    for (int64_t i = 0; i < pA->operationList->length; i++) {
        const struct AlignmentOperation *op = pA->operationList->list[i];
        const char c = op->opType == PAIRWISE_MATCH ? 'M' : (op->opType == PAIRWISE_INDEL_X ? 'D' : 'I');
        if (withProbs) fprintf(fileHandle, " %c %" PRIi64 " %f", c, op->length, op->score);
        else fprintf(fileHandle, " %c %" PRIi64, c, op->length);
    }
    fputc('\n', fileHandle);
}

/* ---- FASTA ---- */

void fastaReadToFunction(FILE *fastaFile, void (*addSeq)(const char *header, const char *sequence, int64_t length)) {
    char *line = NULL, *header = NULL, *seq = NULL;
    size_t cap = 0, seqCap = 0, seqLen = 0;
    for (;;) {
        char *l = read_line(fastaFile, &line, &cap);
        if (l == NULL || l[0] == '>') {
            if (header != NULL) {
                if (seq == NULL) seq = cpecan_malloc(1);
                seq[seqLen] = '\0';
                addSeq(header, seq, (int64_t) seqLen);
                free(header);
                header = NULL;
            }
            if (l == NULL) break;
            header = stString_copy(l + 1);
            seqLen = 0;
            continue;
        }
        if (header == NULL) continue; /* text before the first record */
        for (const char *c = l; *c != '\0'; c++) {
            if (isspace((unsigned char) *c)) continue;
            if (seqLen + 2 > seqCap) {
                seqCap = seqCap ? seqCap * 2 : 1024;
                seq = realloc(seq, seqCap);
                if (seq == NULL) st_errAbort("cpecan: out of memory reading a sequence");
            }
            seq[seqLen++] = *c;
        }
    }
    free(line);
    free(seq);
}

void fastaWrite(const char *sequence, const char *header, FILE *file) {
    fprintf(file, ">%s\n", header);
    const size_t n = strlen(sequence);
    for (size_t i = 0; i < n; i += 100) fprintf(file, "%.*s\n", (int) (n - i < 100 ? n - i : 100), sequence + i);
}
#endif /* CPECAN_USE_SONLIB */
