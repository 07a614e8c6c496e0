/*
 * oracle/shim/bioioC.h -- TEST INFRASTRUCTURE ONLY (see sonLib.h in this directory).
 * The anchoring subprocess path (getBlastPairs, impl/pairwiseAligner.c:1005-1080) is out of
 * scope; these exist only so the file links.
 */
#ifndef ORACLE_SHIM_BIOIOC_H_
#define ORACLE_SHIM_BIOIOC_H_

#include <stdio.h>
#include "pairwiseAlignment.h"

char *getTempFile(void);
void fastaWrite(char *sequence, char *header, FILE *file);
struct PairwiseAlignment *cigarRead(FILE *fileHandle);

#endif
