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
 * handed to the next strip through a small per-warp ring of 64-byte records indexed by diagonal
 * (L2-resident, read one step ahead).  Warps fetch work items from a global counter, so long and short
 * regions balance across the chip.  The step loop is unrolled by two with ping-pong register sets so no
 * state is copied between steps.
 */
#pragma once
#include "kernels.cuh"

namespace cpb {

constexpr int BND_REC = 8; /* doubles per boundary record (S <= 5 used, padded to 64 bytes) */

struct StripArgs {
    const StripRec *strips;   /* per region: (lX>>5)+1 records at RegionDev.stripBase */
    double *boundary;         /* per warp slot: 2 rings of ringSize records */
    int32_t ringSize;         /* power of two >= longest strip diagonal range + 4 */
    int32_t nItems;
    unsigned int *counter;    /* work-fetch counter (zeroed before the launch) */
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

template <int S> __device__ __forceinline__ void load_record(double *v, const double *rec) {
    /* 64-byte aligned record, L2 only (another lane of this warp wrote it) */
    const double2 a = __ldcg(reinterpret_cast<const double2 *>(rec));
    v[0] = a.x;
    v[1] = a.y;
    if (S > 2) {
        const double2 b = __ldcg(reinterpret_cast<const double2 *>(rec) + 1);
        v[2] = b.x;
        if (S > 3) v[3] = b.y;
    }
    if (S > 4) v[4] = __ldcg(rec + 4);
}
template <int S> __device__ __forceinline__ void store_record(double *rec, const double *v) {
    __stcg(reinterpret_cast<double2 *>(rec), make_double2(v[0], v[1]));
    if (S > 2) __stcg(reinterpret_cast<double2 *>(rec) + 1, make_double2(v[2], S > 3 ? v[3] : 0.0));
    if (S > 4) __stcg(rec + 4, v[4]);
}

/* ---------------------------------------------------------------------------------------------
 * k_forward_strip<S, NP, WPC>: NP = state planes written to HBM (0 forward-only, 1 match, 3 match+gaps, S all)
 * ------------------------------------------------------------------------------------------- */
template <int S, int NP, int WPC>
__global__ void __launch_bounds__(32 * WPC) k_forward_strip(const DpArgs a, const CpbModel model, const StripArgs sa) {
    __shared__ __align__(16) StripTables tab;
    fill_strip_tables(tab, model, threadIdx.x);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int slot = blockIdx.x * WPC + (threadIdx.x >> 5);
    const int rm = sa.ringSize - 1;
    double *ring0 = sa.boundary + (size_t) slot * 2 * sa.ringSize * BND_REC;
    double *ring1 = ring0 + (size_t) sa.ringSize * BND_REC;
    const bool keepFull = a.auxF != 0; /* full cells of total diagonals go to the aux records (posterior modes) */

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
        int prevFirst = 1, prevLast = 0; /* diagonal range the previous strip wrote to its boundary ring */

        for (int s = 0; s < nStrips; s++) {
            const StripRec sr = strips[s];
            if (sr.dLast < sr.dFirst) {
                prevFirst = 1;
                prevLast = 0;
                continue;
            }
            const int x = 32 * s + lane;
            const int cX5 = ((x > 0 && x <= R.lX) ? sx[x - 1] : 4) * 5;
            const double eX = tab.eGapX[cX5 / 5];
            double *bOut = (s & 1) ? ring1 : ring0;
            const double *bIn = (s & 1) ? ring0 : ring1;

            /* two register sets, used alternately: own previous output, and lane-1's output received one step ago */
            double ownA[S], ownB[S], rcvA[S], rcvB[S], bNext[S];
#pragma unroll
            for (int k = 0; k < S; k++) {
                ownA[k] = ownB[k] = CPB_NEG_INF;
                rcvA[k] = rcvB[k] = CPB_NEG_INF;
                bNext[k] = CPB_NEG_INF;
            }
            int d0 = sr.dFirst;
            if (d0 == 0) {
                /* diagonal 0: the single cell (0,0) holds the start vector (impl/pairwiseAligner.c:776-777) */
                if (lane == 0) {
                    const double *sv = R.raggedL ? tab.rstartv : tab.startv;
#pragma unroll
                    for (int k = 0; k < S; k++) {
                        ownA[k] = sv[k];
                        if (k < NP) pf[(int64_t) k * a.planeStride] = sv[k];
                    }
                }
                if (lane == 31) store_record<S>(bOut, ownA);
                d0 = 1;
            } else if (lane == 0) {
                /* row x-1 of the previous strip: diagonal d0-2 seeds "middle", d0-1 is the first step's "lower" */
                const int d2 = d0 - 2, d1 = d0 - 1;
                if (d2 >= prevFirst && d2 <= prevLast) load_record<S>(rcvA, bIn + (size_t) (d2 & rm) * BND_REC);
                if (d1 >= prevFirst && d1 <= prevLast) load_record<S>(bNext, bIn + (size_t) (d1 & rm) * BND_REC);
            }
            DiagRec cur = dg[d0 <= N ? d0 : N];
            int cY = 4;
            {
                const int y0 = d0 - x;
                if (y0 > 0 && y0 <= R.lY) cY = sy[y0 - 1];
            }

            /* one step: reads (own, rcvOld), writes (ownNew, rcvNew) */
            auto step = [&](const int d, const double(&own)[S], const double(&rcvOld)[S], double(&ownNew)[S], double(&rcvNew)[S]) {
                const DiagRec nxt = dg[d + 1]; /* prefetch (record N+1 is a sentinel) */
                const int cYnow = cY;
                {
                    const int y1 = d + 1 - x; /* next step's column symbol */
                    cY = (y1 > 0 && y1 <= R.lY) ? sy[y1 - 1] : 4;
                }
#pragma unroll
                for (int k = 0; k < S; k++) rcvNew[k] = shfl_up_f64(own[k]);
                if (lane == 0) {
#pragma unroll
                    for (int k = 0; k < S; k++) rcvNew[k] = bNext[k];
                }
                {
                    /* boundary cell of diagonal d feeds step d+1 (lane 0 only; everyone else keeps LOG_ZERO) */
                    const bool have = lane == 0 && d >= prevFirst && d <= prevLast;
                    if (have) load_record<S>(bNext, bIn + (size_t) (d & rm) * BND_REC);
                    else {
#pragma unroll
                        for (int k = 0; k < S; k++) bNext[k] = CPB_NEG_INF;
                    }
                }
                const int xlo = (d + cur.xmyL) >> 1;
                const int i = x - xlo;
                const bool inBand = i >= 0 && i < cur.width;

                double tl[Shape<S>::NL], tm[Shape<S>::NM], tu[Shape<S>::NU], out[S];
                {
                    const double eM = tab.eMatch[cX5 + cYnow], eY = tab.eGapY[cYnow];
#pragma unroll
                    for (int k = 0; k < Shape<S>::NL; k++) tl[k] = eX + model.tLower[k];
#pragma unroll
                    for (int k = 0; k < Shape<S>::NM; k++) tm[k] = eM + model.tMiddle[k];
#pragma unroll
                    for (int k = 0; k < Shape<S>::NU; k++) tu[k] = eY + model.tUpper[k];
                }
                cell_forward<S>(out, rcvNew, rcvOld, own, tl, tm, tu, tab.ctab);
#pragma unroll
                for (int k = 0; k < S; k++) ownNew[k] = inBand ? out[k] : CPB_NEG_INF;

                if (inBand) {
                    const int cell = (int) cur.coff + i; /* < 2^31 cells per region is enforced on the host */
#pragma unroll
                    for (int k = 0; k < NP; k++) pf[(int64_t) k * a.planeStride + cell] = out[k];
                    if (keepFull && cur.aoff != NO_AUX) {
#pragma unroll
                        for (int k = 0; k < S; k++) aux[(size_t) cur.aoff + (size_t) k * cur.width + i] = out[k];
                    }
                }
                if (lane == 31) store_record<S>(bOut + (size_t) (d & rm) * BND_REC, ownNew);
                cur = nxt;
            };

            int d = d0;
            for (; d + 1 <= sr.dLast; d += 2) {
                step(d, ownA, rcvA, ownB, rcvB);
                step(d + 1, ownB, rcvB, ownA, rcvA);
            }
            if (d <= sr.dLast) {
                step(d, ownA, rcvA, ownB, rcvB);
#pragma unroll
                for (int k = 0; k < S; k++) ownA[k] = ownB[k];
            }
            /* ownA now holds the outputs of the strip's last diagonal */
            if (NP == 0 && a.forwardOut != nullptr && s == nStrips - 1 && sr.dLast == N && N > 0 && lane == (R.lX & 31)) {
                /* computeForwardProbability: the last cell dotted with the end vector (impl/pairwiseAligner.c:910-916) */
                const double *ev = R.raggedR ? tab.rendv : tab.endv;
                double v = ownA[0] + ev[0];
#pragma unroll
                for (int k = 1; k < S; k++) v = log_add(v, ownA[k] + ev[k], tab.ctab);
                a.forwardOut[regionId] = v;
            }
            prevFirst = sr.dFirst;
            prevLast = sr.dLast;
            __syncwarp();
        }
        if (NP == 0 && a.forwardOut != nullptr && N == 0 && lane == 0) a.forwardOut[regionId] = 0.0; /* LOG_ONE for the empty problem */
    }
}

/* ---------------------------------------------------------------------------------------------
 * k_backward_strip<S, NP, WPC> : one warp per traceback block, strips in descending row order
 * ------------------------------------------------------------------------------------------- */
template <int S> struct BwdShare; /* states of row x+1 that row x needs: M (for the middle step) and the gap-X states */
template <> struct BwdShare<5> {
    static constexpr int N = 3;
    __host__ __device__ static constexpr int state(int k) { return k == 0 ? 0 : (k == 1 ? 1 : 3); }
};
template <> struct BwdShare<3> {
    static constexpr int N = 2;
    __host__ __device__ static constexpr int state(int k) { return k; }
};

template <int S, int NP, int WPC>
__global__ void __launch_bounds__(32 * WPC) k_backward_strip(const DpArgs a, const CpbModel model, const StripArgs sa) {
    __shared__ __align__(16) StripTables tab;
    fill_strip_tables(tab, model, threadIdx.x);
    __syncthreads();
    constexpr int NB = BwdShare<S>::N;
    const int lane = threadIdx.x & 31;
    const int slot = blockIdx.x * WPC + (threadIdx.x >> 5);
    const int rm = sa.ringSize - 1;
    double *ring0 = sa.boundary + (size_t) slot * 2 * sa.ringSize * BND_REC;
    double *ring1 = ring0 + (size_t) sa.ringSize * BND_REC;
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
            const int cX5 = (x < R.lX ? sx[x] : 4) * 5; /* symbol of row x+1 */
            const double eX = tab.eGapX[cX5 / 5];
            double *bOut = (s & 1) ? ring1 : ring0;
            const double *bIn = (s & 1) ? ring0 : ring1;

            double ownA[S], ownB[S], bNext[NB], recvM = CPB_NEG_INF;
#pragma unroll
            for (int k = 0; k < S; k++) ownA[k] = ownB[k] = CPB_NEG_INF;
#pragma unroll
            for (int k = 0; k < NB; k++) bNext[k] = CPB_NEG_INF;
            if (lane == 31) {
                /* row x+1 belongs to the strip processed before this one */
                const int d2 = dHi + 2, d1 = dHi + 1;
                if (d2 >= prevLo && d2 <= prevHi) recvM = __ldcg(bIn + (size_t) (d2 & rm) * BND_REC);
                if (d1 >= prevLo && d1 <= prevHi) load_record<NB>(bNext, bIn + (size_t) (d1 & rm) * BND_REC);
            }
            DiagRec cur = dg[dHi];
            int cY = 4;
            {
                const int y0 = dHi - x;
                if (y0 >= 0 && y0 < R.lY) cY = sy[y0]; /* symbol of column y+1 */
            }

            auto step = [&](const int d, const double(&own)[S], double(&ownNew)[S]) {
                const DiagRec nxt = dg[d - 1]; /* d-1 >= T >= 0 */
                const int cYnow = cY;
                {
                    const int y1 = d - 1 - x;
                    cY = (y1 >= 0 && y1 < R.lY) ? sy[y1] : 4;
                }
                double rcv[NB];
#pragma unroll
                for (int k = 0; k < NB; k++) rcv[k] = shfl_down_f64(own[BwdShare<S>::state(k)]);
                if (lane == 31) {
#pragma unroll
                    for (int k = 0; k < NB; k++) rcv[k] = bNext[k];
                }
                {
                    const bool have = lane == 31 && d >= prevLo && d <= prevHi;
                    if (have) load_record<NB>(bNext, bIn + (size_t) (d & rm) * BND_REC);
                    else {
#pragma unroll
                        for (int k = 0; k < NB; k++) bNext[k] = CPB_NEG_INF;
                    }
                }
                const int xlo = (d + cur.xmyL) >> 1;
                const int i = x - xlo;
                const bool inBand = i >= 0 && i < cur.width;

                double out[S];
                {
                    double tl[Shape<S>::NL], tm[Shape<S>::NM], tu[Shape<S>::NU];
                    const double eM = tab.eMatch[cX5 + cYnow], eY = tab.eGapY[cYnow];
#pragma unroll
                    for (int k = 0; k < Shape<S>::NL; k++) tl[k] = eX + model.tLower[k];
#pragma unroll
                    for (int k = 0; k < Shape<S>::NM; k++) tm[k] = eM + model.tMiddle[k];
#pragma unroll
                    for (int k = 0; k < Shape<S>::NU; k++) tu[k] = eY + model.tUpper[k];
                    /* cell_backward reads toU[2], toU[4] (own previous output: cell (x,y+1)) and toL[1], toL[3] (row x+1) */
                    double toL[S];
#pragma unroll
                    for (int k = 0; k < S; k++) toL[k] = CPB_NEG_INF;
#pragma unroll
                    for (int k = 1; k < NB; k++) toL[BwdShare<S>::state(k)] = rcv[k];
                    cell_backward<S>(out, recvM, own, toL, tm, tu, tl, tab.ctab);
                }
                if (d == top) {
#pragma unroll
                    for (int k = 0; k < S; k++) out[k] = endVec[k];
                }
#pragma unroll
                for (int k = 0; k < S; k++) ownNew[k] = inBand ? out[k] : CPB_NEG_INF;

                if (inBand) {
                    const int cell = (int) cur.coff + i;
                    if (d <= from) {
#pragma unroll
                        for (int k = 0; k < NP; k++) pb[(int64_t) k * a.planeStride + cell] = out[k];
                        if (cur.aoff != NO_AUX) {
                            /* cell_dotProduct(F[d], B[d]) (impl/pairwiseAligner.c:402-408); folded over the cells by k_totals */
                            double f[S];
                            if (nF != 0) {
#pragma unroll
                                for (int k = 0; k < S; k++) f[k] = aux[(size_t) cur.aoff + (size_t) k * cur.width + i];
                            } else {
#pragma unroll
                                for (int k = 0; k < S; k++) f[k] = pf[(int64_t) k * a.planeStride + cell];
                            }
                            double t = f[0] + out[0];
#pragma unroll
                            for (int k = 1; k < S; k++) t = log_add(t, f[k] + out[k], tab.ctab);
                            aux[(size_t) cur.aoff + (size_t) nF * cur.width + i] = t;
                        }
                    }
                    if (d - 1 > T && d - 1 <= from && nxt.aoff != NO_AUX) {
                        /* diagonal d-1 is a total diagonal: its second term is the fold of F[d].M + B[d].M (:643-651) */
                        aux[(size_t) nxt.aoff + (size_t) (nF + 1) * nxt.width + i] = pf[cell] + out[0];
                    }
                }
                if (lane == 0) {
                    double rec[NB];
#pragma unroll
                    for (int k = 0; k < NB; k++) rec[k] = ownNew[BwdShare<S>::state(k)];
                    store_record<NB>(bOut + (size_t) (d & rm) * BND_REC, rec);
                }
                recvM = rcv[0]; /* M of (x+1, .) on diagonal d+1 is the middle neighbour of the next step */
                cur = nxt;
            };

            int d = dHi;
            for (; d - 1 >= dLo; d -= 2) {
                step(d, ownA, ownB);
                step(d - 1, ownB, ownA);
            }
            if (d >= dLo) step(d, ownA, ownB);
            prevHi = dHi;
            prevLo = dLo;
            __syncwarp();
        }
    }
}

} /* namespace cpb */
