/*
 * host/logAdd.c -- logAdd (inc/pairwiseAligner.h:167 of the reference), the one scalar of the recurrence the reference's
 * header exposes.  Nothing in this library calls it: the DP runs on the device with its own copy (csrc/kernels.cuh log_add).
 * It is here so that callers and the reference's test_logAdd (tests/pairwiseAlignerTest.c:134-144) link.
 *
 * log(e^x + e^y) as the smaller operand plus a cubic in the difference d = |x - y| (impl/pairwiseAligner.c:287-307): four
 * segments, d <= 1, <= 2.5, <= 4.5 and above, Horner form, coefficients that are FLOAT literals in the reference (so they are
 * rounded to float before they are promoted); from d = 7.5 on, and when one side is LOG_ZERO, the larger operand is returned.
 * Compiled with -ffp-contract=off: every multiply and add is rounded separately, as in the reference's x86-64 build.
 */
#include <math.h>

#include "cpecan/pairwiseAligner.h"

static double cubic(double d) {
    static const float k[4][4] = { { -0.009350833524763f, 0.130659527668286f, 0.498799810682272f, 0.693203116424741f },
                                   { -0.014532321752540f, 0.139942324101744f, 0.495635523139337f, 0.692140569840976f },
                                   { -0.004605031767994f, 0.063427417320019f, 0.695956496475118f, 0.514272634594009f },
                                   { -0.000458661602210f, 0.009695946122598f, 0.930734667215156f, 0.168037164329057f } };
    const float *q = k[d <= 1.0 ? 0 : (d <= 2.5 ? 1 : (d <= 4.5 ? 2 : 3))];
    return (((double) q[0] * d + (double) q[1]) * d + (double) q[2]) * d + (double) q[3];
}

double logAdd(double x, double y) {
    const double lo = x < y ? x : y, hi = x < y ? y : x;
    if (lo == LOG_ZERO || hi - lo >= 7.5) return hi;
    return cubic(hi - lo) + lo;
}
