/*
 * cpecan/pairwiseAlignment.h -- cigar-style pairwise alignments and the sequence / cigar file I/O that
 * cPecanRealign needs (cPecanRealign.c:50-92, :509-523, :591-599 of the reference).
 *
 * In the reference these types and functions come from sonLib (pairwiseAlignment.h, bioioC.h, commonC.h), which is
 * not part of cPecan's tree (include.mk:2).  libcpecan.so carries its own implementation under the same names, with
 * the fields the reference's callers touch (pA->contig1, ->start1, ->end1, ->strand1, ->score, ->operationList->length,
 * ->operationList->list[i], op->opType, op->length, op->score).  A build against a real sonLib defines
 * CPECAN_USE_SONLIB and takes sonLib's own headers instead.
 *
 * File format (what LASTZ --format=cigar emits and cigarRead / cigarWrite of sonLib exchange), one alignment per line:
 *     cigar: <contig2> <start2> <end2> <strand2> <contig1> <start1> <end1> <strand1> <score> {<op> <length>}*
 * query (sequence 2) first, then target (sequence 1); strand '+' / '-'; op 'M' = match, 'D' = PAIRWISE_INDEL_X
 * (bases of sequence 1 only), 'I' = PAIRWISE_INDEL_Y (bases of sequence 2 only).  The reference's own use pins the
 * field order: impl/pairwiseAligner.c:1025-1047 runs `lastz <a> <b>` and asserts contig1 == "a", contig2 == "b".
 */
#ifndef CPECAN_PAIRWISEALIGNMENT_H_
#define CPECAN_PAIRWISEALIGNMENT_H_

#include "cpecan/sonLibLite.h"

#ifndef CPECAN_USE_SONLIB
#ifdef __cplusplus
extern "C" {
#endif

#define PAIRWISE_MATCH 0
#define PAIRWISE_INDEL_X 1
#define PAIRWISE_INDEL_Y 2

/* commonC.h's growable pointer array, as far as the alignment code uses it */
struct List {
    int64_t length;
    int64_t maxLength;
    void **list;
    void (*destructElement)(void *);
};
struct List *constructEmptyList(int64_t length, void (*destructElement)(void *));
void listAppend(struct List *list, void *item);
void destructList(struct List *list);

struct AlignmentOperation {
    int64_t opType;
    int64_t length;
    double score;
};
struct AlignmentOperation *constructAlignmentOperation(int64_t opType, int64_t length, double score);
void destructAlignmentOperation(struct AlignmentOperation *op);

struct PairwiseAlignment {
    char *contig1;
    int64_t start1, end1, strand1; /* strand 1 = '+': start <= end; 0 = '-': start >= end */
    char *contig2;
    int64_t start2, end2, strand2;
    double score;
    struct List *operationList;
};
/* copies the contig names, takes ownership of operationList */
struct PairwiseAlignment *constructPairwiseAlignment(const char *contig1, int64_t start1, int64_t end1, int64_t strand1, const char *contig2,
                                                     int64_t start2, int64_t end2, int64_t strand2, double score, struct List *operationList);
void destructPairwiseAlignment(struct PairwiseAlignment *pA);
/* aborts (st_errAbort) unless the operation lengths add up to |end - start| on both sequences and all lengths are positive */
void checkPairwiseAlignment(struct PairwiseAlignment *pA);

/* next alignment of the stream, or NULL at end of file; lines that do not start with "cigar:" are skipped */
struct PairwiseAlignment *cigarRead(FILE *fileHandle);
void cigarWrite(FILE *fileHandle, struct PairwiseAlignment *pA, int64_t withProbs);

/* calls addSeq(header line without '>', sequence without white space, length) for every record */
void fastaReadToFunction(FILE *fastaFile, void (*addSeq)(const char *header, const char *sequence, int64_t length));
void fastaWrite(const char *sequence, const char *header, FILE *file);

char *stString_copy(const char *s);
char *stString_getSubString(const char *s, int64_t start, int64_t length);
char *stString_reverseComplementString(const char *s);

#ifdef __cplusplus
}
#endif
#endif /* CPECAN_USE_SONLIB */
#endif /* CPECAN_PAIRWISEALIGNMENT_H_ */
