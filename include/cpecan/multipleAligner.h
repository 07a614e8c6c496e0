/*
 * cpecan/multipleAligner.h -- the one entry point of the reference's inc/multipleAligner.h that cPecanRealign needs.
 * The poset multiple aligner itself (impl/multipleAligner.c) is a caller of the pairwise path and is not part of this
 * library.
 */
#ifndef CPECAN_MULTIPLEALIGNER_H_
#define CPECAN_MULTIPLEALIGNER_H_

#include "cpecan/sonLibLite.h"

#ifdef __cplusplus
extern "C" {
#endif

/* inc/multipleAligner.h:67 -- the heaviest chain of pairs, strictly increasing in x and y, among the pairs whose weight is at least
 * matchGamma * PAIR_ALIGNMENT_PROB_1 (and positive).  Consumes alignedPairs; returns a new list of (weight, x, y), x ascending. */
stList *filterPairwiseAlignmentToMakePairsOrdered(stList *alignedPairs, const char *seqX, const char *seqY, float matchGamma);

#ifdef __cplusplus
}
#endif
#endif
