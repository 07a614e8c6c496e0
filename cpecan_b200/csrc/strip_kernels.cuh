/*
 * strip_kernels.cuh -- barrier-free forward / backward wavefronts: one warp per region (forward) or per
 * traceback block (backward), no shared-memory window, no band-width limit.
 *
 * Lane l of the warp owns matrix row x = 32*s + l of the current row strip s and walks the anti-diagonals
 * d that touch the strip (k_band records [dFirst, dLast] per strip).  On diagonal d it computes the cell
 * (x, d-x); everything outside the band is LOG_ZERO (the reference never creates those cells, and logAdd
 * with LOG_ZERO is the identity, so the arithmetic that remains is the reference's, bit for bit):
 *   forward : lower (x-1,y) and middle (x-1,y-1) are lane l-1's outputs one and two steps ago -> one
 *             shuffle-up of the previous output per step; upper (x,y-1) is the lane's own previous output;
 *   backward: (x+1,y) and (x+1,y+1) come from lane l+1 by shuffle-down, (x,y+1) is the lane's own.
 * Strips are processed in row order (forward) / reverse row order (backward); the edge row of a strip is
 * handed to the next strip through a small per-warp boundary array indexed by diagonal (L2-resident).
 * Warps fetch work items from a global counter, so long and short regions balance across the chip.
 */
#pragma once
#include "kernels.cuh"

namespace cpb {

struct StripArgs {
    const StripRec *strips;   /* per region: (lX>>5)+1 records at RegionDev.stripBase */
    double *boundary;         /* per warp slot: 2 x planes x bndStride doubles */
    int64_t bndStride;        /* ring size: power of two >= longest strip diagonal range + 4 */
    unsigned int *counter;    /* work-fetch counter (zeroed before the launch) */
    int32_t nItems;
    int32_t pad_;
};

struct StripTables {
    double ctab[16];
    double eGapX[5], eGapY[5], eMatch[25];
    double startv[5], rstartv[5], endv[5], rendv[5];
};

__device__ __forceinline__ void fill_strip_tables(StripTables &t, const CpbModel &m, int tid) {
    fill_coefficients(t.ctab, tid);
    if (tid < 5) {
        t.eGapX[tid] = m.eGapX[tid];
        t.eGapY[tid] = m.eGapY[tid];
        t.startv[tid] = m.start[tid];
        t.rstartv[tid] = m.raggedStart[tid];
        t.endv[tid] = m.end[tid];
        t.rendv[tid] = m.raggedEnd[tid];
    }
    if (tid < 25) t.eMatch[tid] = m.eMatch[tid];
}

__device__ __forceinline__ double shfl_up_f64(double v) { return __shfl_up_sync(0xFFFFFFFFu, v, 1); }
__device__ __forceinline__ double shfl_down_f64(double v) { return __shfl_down_sync(0xFFFFFFFFu, v, 1); }

/* ---------------------------------------------------------------------------------------------
 * k_forward_strip
 * ------------------------------------------------------------------------------------------- */
template <int S, int WPC>
__global__ void __launch_bounds__(32 * WPC) k_forward_strip(const DpArgs a, const CpbModel model, const StripArgs sa) {
    __shared__ __align__(16) StripTables tab;
    fill_strip_tables(tab, model, threadIdx.x);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int slot = blockIdx.x * WPC + (threadIdx.x >> 5);
    double *bnd = sa.boundary + (int64_t) slot * 2 * S * sa.bndStride;
    const int64_t bs = sa.bndStride;
    const int rm = (int) sa.bndStride - 1; /* ring mask */

    for (;;) {
        unsigned item = 0;
        if (lane == 0) item = atomicAdd(sa.counter, 1u);
        item = __shfl_sync(0xFFFFFFFFu, item, 0);
        if (item >= (unsigned) sa.nItems) break;
        const int regionId = a.list[item];
        const RegionDev R = a.regions[regionId];
        const int N = R.lX + R.lY;
        const DiagRec *dg = a.diags + R.diagBase;
        const uint8_t *sx = a.symX + R.xBase, *sy = a.symY + R.yBase;
        double *pf = a.planesF + R.cellBase;
        double *aux = a.aux + R.auxBase;
        const StripRec *strips = sa.strips + R.stripBase;
        const int nStrips = (R.lX >> 5) + 1;
        const double *startVec = R.raggedL ? tab.rstartv : tab.startv;
        int prevFirst = 1, prevLast = 0; /* diagonal range the previous strip wrote to the boundary */

        for (int s = 0; s < nStrips; s++) {
            const StripRec sr = strips[s];
            if (sr.dLast < sr.dFirst) {
                prevFirst = 1;
                prevLast = 0;
                continue;
            }
            const int x = 32 * s + lane;
            const int cX = (x > 0 && x <= R.lX) ? sx[x - 1] : 4;
            const double eX = tab.eGapX[cX];
            double *bOut = bnd + (int64_t) (s & 1) * S * bs;
            const double *bIn = bnd + (int64_t) ((s & 1) ^ 1) * S * bs;

            double outPrev[S], recvPrev[S], bNext[S];
#pragma unroll
            for (int k = 0; k < S; k++) {
                outPrev[k] = CPB_NEG_INF;
                recvPrev[k] = CPB_NEG_INF;
                bNext[k] = CPB_NEG_INF;
            }
            if (lane == 0) {
                /* row x-1 of the previous strip: diagonal dFirst-2 seeds "middle", dFirst-1 is prefetched for the first step */
                const int d2 = sr.dFirst - 2, d1 = sr.dFirst - 1;
                if (d2 >= prevFirst && d2 <= prevLast) {
#pragma unroll
                    for (int k = 0; k < S; k++) recvPrev[k] = __ldcg(bIn + k * bs + (d2 & rm));
                }
                if (d1 >= prevFirst && d1 <= prevLast) {
#pragma unroll
                    for (int k = 0; k < S; k++) bNext[k] = __ldcg(bIn + k * bs + (d1 & rm));
                }
            }
            DiagRec cur = dg[sr.dFirst];
            int yn = sr.dFirst - x;
            int cYn = (yn > 0 && yn <= R.lY) ? sy[yn - 1] : 4;

            for (int d = sr.dFirst; d <= sr.dLast; d++) {
                const DiagRec nxt = dg[d + 1]; /* prefetch (record N+1 is a sentinel) */
                const int cY = cYn;
                {
                    /* prefetch the next step's column symbol and boundary cell */
                    const int y1 = d + 1 - x;
                    cYn = (y1 > 0 && y1 <= R.lY) ? sy[y1 - 1] : 4;
                }
                double recvNow[S];
#pragma unroll
                for (int k = 0; k < S; k++) recvNow[k] = shfl_up_f64(outPrev[k]);
                if (lane == 0) {
#pragma unroll
                    for (int k = 0; k < S; k++) recvNow[k] = bNext[k];
                    const bool have = d >= prevFirst && d <= prevLast; /* boundary of diagonal d feeds step d+1 */
#pragma unroll
                    for (int k = 0; k < S; k++) bNext[k] = have ? __ldcg(bIn + k * bs + (d & rm)) : CPB_NEG_INF;
                }
                const int xlo = (d + cur.xmyL) >> 1;
                const bool inBand = x >= xlo && x < xlo + cur.width;

                double tl[Shape<S>::NL], tm[Shape<S>::NM], tu[Shape<S>::NU], out[S];
                {
                    const double eM = tab.eMatch[cX * 5 + cY], eY = tab.eGapY[cY];
#pragma unroll
                    for (int k = 0; k < Shape<S>::NL; k++) tl[k] = eX + model.tLower[k];
#pragma unroll
                    for (int k = 0; k < Shape<S>::NM; k++) tm[k] = eM + model.tMiddle[k];
#pragma unroll
                    for (int k = 0; k < Shape<S>::NU; k++) tu[k] = eY + model.tUpper[k];
                }
                cell_forward<S>(out, recvNow, recvPrev, outPrev, tl, tm, tu, tab.ctab);
                if (d == 0) {
                    /* the single cell (0,0) holds the start vector (impl/pairwiseAligner.c:776-777) */
#pragma unroll
                    for (int k = 0; k < S; k++) out[k] = startVec[k];
                }
#pragma unroll
                for (int k = 0; k < S; k++) out[k] = inBand ? out[k] : CPB_NEG_INF;

                if (inBand) {
                    const int64_t cell = (int64_t) cur.coff + (x - xlo);
#pragma unroll
                    for (int k = 0; k < S; k++) {
                        if (k < a.nPlanes) pf[(int64_t) k * a.planeStride + cell] = out[k];
                    }
                    if (a.auxF != 0 && cur.aoff != NO_AUX) {
#pragma unroll
                        for (int k = 0; k < S; k++) aux[(int64_t) cur.aoff + (int64_t) k * cur.width + (x - xlo)] = out[k];
                    }
                    if (a.forwardOut != nullptr && d == N && N > 0) {
                        /* computeForwardProbability: the last cell dotted with the end vector (impl/pairwiseAligner.c:910-916) */
                        const double *ev = R.raggedR ? tab.rendv : tab.endv;
                        double v = out[0] + ev[0];
#pragma unroll
                        for (int k = 1; k < S; k++) v = log_add(v, out[k] + ev[k], tab.ctab);
                        a.forwardOut[regionId] = v;
                    }
                }
                if (lane == 31) {
#pragma unroll
                    for (int k = 0; k < S; k++) __stcg(bOut + k * bs + (d & rm), out[k]);
                }
#pragma unroll
                for (int k = 0; k < S; k++) {
                    recvPrev[k] = recvNow[k];
                    outPrev[k] = out[k];
                }
                cur = nxt;
            }
            prevFirst = sr.dFirst;
            prevLast = sr.dLast;
            __syncwarp();
        }
        if (a.forwardOut != nullptr && N == 0 && lane == 0) a.forwardOut[regionId] = 0.0; /* LOG_ONE for the empty problem */
    }
}

/* ---------------------------------------------------------------------------------------------
 * k_backward_strip : one warp per traceback block, strips in descending row order
 * ------------------------------------------------------------------------------------------- */
template <int S> struct BwdShare; /* states of (x+1, .) that row x needs: M (for the middle step) and the gap-X states */
template <> struct BwdShare<5> {
    static constexpr int N = 3;
    __device__ static constexpr int state(int k) { return k == 0 ? 0 : (k == 1 ? 1 : 3); }
};
template <> struct BwdShare<3> {
    static constexpr int N = 2;
    __device__ static constexpr int state(int k) { return k; }
};

template <int S, int WPC>
__global__ void __launch_bounds__(32 * WPC) k_backward_strip(const DpArgs a, const CpbModel model, const StripArgs sa) {
    __shared__ __align__(16) StripTables tab;
    fill_strip_tables(tab, model, threadIdx.x);
    __syncthreads();
    constexpr int NB = BwdShare<S>::N;
    const int lane = threadIdx.x & 31;
    const int slot = blockIdx.x * WPC + (threadIdx.x >> 5);
    double *bnd = sa.boundary + (int64_t) slot * 2 * NB * sa.bndStride;
    const int64_t bs = sa.bndStride;
    const int rm = (int) sa.bndStride - 1; /* ring mask */
    const int nF = a.auxF;

    for (;;) {
        unsigned item = 0;
        if (lane == 0) item = atomicAdd(sa.counter, 1u);
        item = __shfl_sync(0xFFFFFFFFu, item, 0);
        if (item >= (unsigned) sa.nItems) break;
        const BlockRec K = a.blocks[a.list[item]];
        const RegionDev R = a.regions[K.region];
        const DiagRec *dg = a.diags + R.diagBase;
        const uint8_t *sx = a.symX + R.xBase, *sy = a.symY + R.yBase;
        const double *pf = a.planesF + R.cellBase;
        double *pb = a.planesB + R.cellBase;
        double *aux = a.aux + R.auxBase;
        const StripRec *strips = sa.strips + R.stripBase;
        const int nStrips = (R.lX >> 5) + 1;
        const int top = K.top, T = K.T, from = K.from;
        const double *endVec = (K.atEnd && R.raggedR) ? tab.rendv : tab.endv;
        int prevHi = 0, prevLo = 1; /* diagonal range the previously processed (higher) strip wrote */

        for (int s = nStrips - 1; s >= 0; s--) {
            const StripRec sr = strips[s];
            const int dHi = min(sr.dLast, top), dLo = max(sr.dFirst, T + 1);
            if (dHi < dLo) { /* no row of this strip is inside the band on the block's diagonals */
                prevHi = 0;
                prevLo = 1;
                continue;
            }
            const int x = 32 * s + lane;
            const int cX = x < R.lX ? sx[x] : 4; /* symbol of row x+1 */
            const double eX = tab.eGapX[cX];
            double *bOut = bnd + (int64_t) (s & 1) * NB * bs;
            const double *bIn = bnd + (int64_t) ((s & 1) ^ 1) * NB * bs;

            double outPrev[S], recvPrevM = CPB_NEG_INF, bNext[NB];
#pragma unroll
            for (int k = 0; k < S; k++) outPrev[k] = CPB_NEG_INF;
#pragma unroll
            for (int k = 0; k < NB; k++) bNext[k] = CPB_NEG_INF;
            if (lane == 31) {
                /* row x+1 belongs to the strip processed before this one */
                const int d2 = dHi + 2, d1 = dHi + 1;
                if (d2 >= prevLo && d2 <= prevHi) recvPrevM = __ldcg(bIn + (d2 & rm));
                if (d1 >= prevLo && d1 <= prevHi) {
#pragma unroll
                    for (int k = 0; k < NB; k++) bNext[k] = __ldcg(bIn + k * bs + (d1 & rm));
                }
            }
            DiagRec cur = dg[dHi];
            int yn = dHi - x;
            int cYn = (yn >= 0 && yn < R.lY) ? sy[yn] : 4; /* symbol of column y+1 */

            for (int d = dHi; d >= dLo; d--) {
                const DiagRec nxt = dg[d - 1]; /* d-1 >= T >= 0 */
                const int cY = cYn;
                {
                    const int y1 = d - 1 - x;
                    cYn = (y1 >= 0 && y1 < R.lY) ? sy[y1] : 4;
                }
                double recvNow[NB];
#pragma unroll
                for (int k = 0; k < NB; k++) recvNow[k] = shfl_down_f64(outPrev[BwdShare<S>::state(k)]);
                if (lane == 31) {
#pragma unroll
                    for (int k = 0; k < NB; k++) recvNow[k] = bNext[k];
                    const bool have = d >= prevLo && d <= prevHi;
#pragma unroll
                    for (int k = 0; k < NB; k++) bNext[k] = have ? __ldcg(bIn + k * bs + (d & rm)) : CPB_NEG_INF;
                }
                const int xlo = (d + cur.xmyL) >> 1;
                const bool inBand = x >= xlo && x < xlo + cur.width;
                const int i = x - xlo;

                double out[S];
                {
                    double tl[Shape<S>::NL], tm[Shape<S>::NM], tu[Shape<S>::NU];
                    const double eM = tab.eMatch[cX * 5 + cY], eY = tab.eGapY[cY];
#pragma unroll
                    for (int k = 0; k < Shape<S>::NL; k++) tl[k] = eX + model.tLower[k];
#pragma unroll
                    for (int k = 0; k < Shape<S>::NM; k++) tm[k] = eM + model.tMiddle[k];
#pragma unroll
                    for (int k = 0; k < Shape<S>::NU; k++) tu[k] = eY + model.tUpper[k];
                    /* cell_backward reads toU[2], toU[4] (own previous output) and toL[1], toL[3] (row x+1) */
                    double toL[S];
#pragma unroll
                    for (int k = 0; k < S; k++) toL[k] = CPB_NEG_INF;
#pragma unroll
                    for (int k = 1; k < NB; k++) toL[BwdShare<S>::state(k)] = recvNow[k];
                    cell_backward<S>(out, recvPrevM, outPrev, toL, tm, tu, tl, tab.ctab);
                }
                if (d == top) {
#pragma unroll
                    for (int k = 0; k < S; k++) out[k] = endVec[k];
                }
#pragma unroll
                for (int k = 0; k < S; k++) out[k] = inBand ? out[k] : CPB_NEG_INF;

                if (inBand) {
                    const int64_t cell = (int64_t) cur.coff + i;
                    if (d <= from) {
#pragma unroll
                        for (int k = 0; k < S; k++) {
                            if (k < a.nPlanes) pb[(int64_t) k * a.planeStride + cell] = out[k];
                        }
                        if (cur.aoff != NO_AUX) {
                            /* cell_dotProduct(F[d], B[d]) (impl/pairwiseAligner.c:402-408); folded over the cells by k_totals */
                            double f[S];
                            if (nF != 0) {
#pragma unroll
                                for (int k = 0; k < S; k++) f[k] = aux[(int64_t) cur.aoff + (int64_t) k * cur.width + i];
                            } else {
#pragma unroll
                                for (int k = 0; k < S; k++) f[k] = pf[(int64_t) k * a.planeStride + cell];
                            }
                            double t = f[0] + out[0];
#pragma unroll
                            for (int k = 1; k < S; k++) t = log_add(t, f[k] + out[k], tab.ctab);
                            aux[(int64_t) cur.aoff + (int64_t) nF * cur.width + i] = t;
                        }
                    }
                    if (d - 1 > T && d - 1 <= from && nxt.aoff != NO_AUX) {
                        /* diagonal d-1 is a total diagonal: its second term is the fold of F[d].M + B[d].M (:643-651) */
                        aux[(int64_t) nxt.aoff + (int64_t) (nF + 1) * nxt.width + i] = pf[cell] + out[0];
                    }
                }
                if (lane == 0) {
#pragma unroll
                    for (int k = 0; k < NB; k++) __stcg(bOut + k * bs + (d & rm), out[BwdShare<S>::state(k)]);
                }
                recvPrevM = recvNow[0];
#pragma unroll
                for (int k = 0; k < S; k++) outPrev[k] = out[k];
                cur = nxt;
            }
            prevHi = dHi;
            prevLo = dLo;
            __syncwarp();
        }
    }
}

} /* namespace cpb */
