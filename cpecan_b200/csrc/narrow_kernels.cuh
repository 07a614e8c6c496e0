/*
 * narrow_kernels.cuh -- forward / backward wavefronts for narrow bands: several regions (forward) or traceback blocks (backward)
 * per warp.
 *
 * The strip kernels give a warp one region and a lane one matrix row; on an anti-diagonal of a band that is W diagonals wide only
 * ~W/2 cells exist, so with cPecanRealign's own band (expansion 4: 9 diagonals, <= 5 cells per anti-diagonal) or cPecanEm's
 * (expansion 10: <= 11 cells) most lanes of the warp idle.  Here a GROUP of G = 8 or 16 lanes owns a region and lane i of the group
 * owns cell i of the current anti-diagonal (requires every diagonal of the region to have at most G cells: RegionDev.maxW <= G);
 * the 32/G groups of a warp walk their own regions in lockstep, one diagonal per step each, and fetch a new region from the work
 * counter as they finish.  A whole diagonal lives in the group's registers, so there are no strips and no boundary rings.
 *
 * Same arithmetic as the strip kernels, in the same order: a cell, once computed, folds the transition groups of its three
 * successors (middle_fold / lower_folds / upper_folds, the reference's order, impl/stateMachine.c:454-479, :695-713) and the
 * successor picks the finished sums up with an indexed shuffle: from cell i-ish of the previous diagonal (lower, upper) and of the
 * one before (middle), the index shifting with the band's left edge.  Cells outside the band are LOG_ZERO.  Planes and aux records are written
 * exactly where the strip kernels write them, so k_totals / k_posterior / k_expect do not care which kernel ran.
 */
#pragma once
#include "strip_kernels.cuh"

namespace cpb {

struct NarrowArgs {
    int32_t nItems;
    int32_t pad_;
    unsigned int *counter; /* work-fetch counter (zeroed before the launch) */
};

/* the upper group folded by the source cell (x, y) for the cell (x, y+1); tu = gap-Y emission of column y+1 + upper transitions */
template <int S> __device__ __forceinline__ void upper_folds(double *u, const double *c, const double *tu, const LaTable la) {
    if constexpr (S == 5) {
        u[0] = log_add(c[0] + tu[0], c[2] + tu[1], la); /* -> shortGapY */
        u[1] = log_add(c[0] + tu[2], c[4] + tu[3], la); /* -> longGapY */
    } else {
        u[0] = log_add(log_add(c[0] + tu[0], c[2] + tu[1], la), c[1] + tu[2], la); /* -> gapY */
    }
}

template <int S> struct NarrowShape;
template <> struct NarrowShape<5> {
    static constexpr int NG = 2; /* gap states per side */
    __host__ __device__ static constexpr int gap_x(int k) { return k == 0 ? 1 : 3; }
    __host__ __device__ static constexpr int gap_y(int k) { return k == 0 ? 2 : 4; }
};
template <> struct NarrowShape<3> {
    static constexpr int NG = 1;
    __host__ __device__ static constexpr int gap_x(int) { return 1; }
    __host__ __device__ static constexpr int gap_y(int) { return 2; }
};

/* ---------------------------------------------------------------------------------------------
 * k_forward_narrow<S, NP, G, WPC>: one group of G lanes per region, diagonals 0 .. lX+lY.
 * A step needs what the cells of the two previous diagonals folded for their successors; the three sets (d, d-1, d-2) and the
 * three diagonal records in flight rotate through fixed registers, three steps per loop iteration, so nothing is copied.
 * ------------------------------------------------------------------------------------------- */
template <int S> struct NarrowFwdSet {
    double g[NarrowShape<S>::NG], u[NarrowShape<S>::NG], m; /* lower folds, upper folds and middle fold of this lane's cell */
    int xmyL, w;                                            /* left edge and width of the diagonal */
};

template <int S, int NP, int G, int WPC>
__global__ void __launch_bounds__(32 * WPC, CPB_STRIP_MIN_BLOCKS) k_forward_narrow(const DpArgs a, const CpbModel model, const NarrowArgs na) {
    __shared__ __align__(128) StripTables<S> tab;
    fill_strip_tables<S>(tab, model, threadIdx.x, 32 * WPC);
    __syncthreads();
    constexpr int NG = NarrowShape<S>::NG, NL = Shape<S>::NL, NM = Shape<S>::NM, NU = Shape<S>::NU;
    const LaTable la = logadd_lane_table(tab.la);
    const int l16 = threadIdx.x & 15, lane = threadIdx.x & 31, gl = lane & (G - 1), gbase = lane & ~(G - 1);
    const unsigned gmask = G == 32 ? 0xFFFFFFFFu : (((1u << G) - 1u) << gbase);
    const bool keepFull = a.auxF != 0;

    bool active = false, exhausted = false;
    int d = 0, N = 0, lX = 0, lY = 0;
    const DiagRec *dg = nullptr;
    const uint8_t *sx = nullptr, *sy = nullptr;
    double *pf = nullptr, *aux = nullptr;
    /* diagonal records are fetched three steps ahead and the symbols of a diagonal's cells one step ahead (records N+1, N+2 are
     * sentinels): nothing a step needs is a load issued in that step */
    DiagRec R0, R1, R2;
    R0.xmyL = R1.xmyL = R2.xmyL = 0;
    R0.width = R1.width = R2.width = 0;
    R0.coff = R1.coff = R2.coff = 0;
    R0.aoff = R1.aoff = R2.aoff = NO_AUX;
    int cXn = 4, cYn = 4; /* symbols of row x+1 and column y+1 of this lane's cell on the current diagonal */
    NarrowFwdSet<S> M0, M1, M2;
    auto clear = [&](NarrowFwdSet<S> &q) {
#pragma unroll
        for (int k = 0; k < NG; k++) q.g[k] = q.u[k] = CPB_NEG_INF;
        q.m = CPB_NEG_INF;
        q.xmyL = 0;
        q.w = 0;
    };
    clear(M0);
    clear(M1);
    clear(M2);

    /* one diagonal: `rec` is its record (replaced by the one three diagonals on), `recNext` the next diagonal's */
    auto step = [&](DiagRec &rec, const DiagRec &recNext, const NarrowFwdSet<S> &in1, const NarrowFwdSet<S> &in2, NarrowFwdSet<S> &out) {
        const int width = active ? rec.width : 0;
        const bool valid = gl < width;
        const int xmy = rec.xmyL + 2 * gl;
        /* the cell: finished sums from the lower (x-1, y) and upper (x, y-1) neighbours on d-1 and the middle one (x-1, y-1) on d-2 */
        const int iL = (xmy - 1 - in1.xmyL) >> 1, iU = iL + 1, iM = (xmy - in2.xmyL) >> 1;
        double cell[S];
        {
            double g[NG], u[NG];
#pragma unroll
            for (int k = 0; k < NG; k++) {
                g[k] = __shfl_sync(0xFFFFFFFFu, in1.g[k], gbase + (iL & (G - 1)));
                u[k] = __shfl_sync(0xFFFFFFFFu, in1.u[k], gbase + (iU & (G - 1)));
            }
            const double m = __shfl_sync(0xFFFFFFFFu, in2.m, gbase + (iM & (G - 1)));
            const bool okL = valid && (unsigned) iL < (unsigned) in1.w, okU = valid && (unsigned) iU < (unsigned) in1.w;
            const bool okM = valid && (unsigned) iM < (unsigned) in2.w;
            cell[0] = okM ? m : CPB_NEG_INF;
#pragma unroll
            for (int k = 0; k < NG; k++) {
                cell[NarrowShape<S>::gap_x(k)] = okL ? g[k] : CPB_NEG_INF;
                cell[NarrowShape<S>::gap_y(k)] = okU ? u[k] : CPB_NEG_INF;
            }
        }
        if (valid) {
            const int c = (int) rec.coff + gl;
#pragma unroll
            for (int k = 0; k < NP; k++) pf[(int64_t) k * a.planeStride + c] = cell[k];
            if (keepFull && rec.aoff != NO_AUX) {
#pragma unroll
                for (int k = 0; k < S; k++) aux[(size_t) rec.aoff + (size_t) k * rec.width + gl] = cell[k];
            }
        }
        /* fold for the successors: (x+1, y) and (x, y+1) on d+1, (x+1, y+1) on d+2 */
        {
            double tlD[NL], tmD[NM], tu[NU];
            load_row<NL>(tlD, tab.tl[cXn]);
            const double eM = tab.eM[cXn * 6 + cYn][l16], eY = tab.eY[cYn][l16];
#pragma unroll
            for (int k = 0; k < NM; k++) tmD[k] = eM + model.tMiddle[k];
#pragma unroll
            for (int k = 0; k < NU; k++) tu[k] = eY + model.tUpper[k];
            out.m = middle_fold<S>(cell, tmD, la);
            lower_folds<S>(out.g, cell, tlD, la);
            upper_folds<S>(out.u, cell, tu, la);
        }
        out.xmyL = rec.xmyL;
        out.w = width;
        if (active) {
            if (++d > N) {
                active = false;
            } else {
                /* symbols for the cells of the next diagonal, and the record three diagonals on */
                const int xmyN = recNext.xmyL + 2 * gl, xN = (d + xmyN) >> 1, yN = (d - xmyN) >> 1;
                const bool validN = gl < recNext.width;
                cXn = (validN && xN < lX) ? sx[xN] : 4;
                cYn = (validN && yN < lY) ? sy[yN] : 4;
                rec = dg[min(d + 2, N + 2)];
            }
        }
    };

    for (;;) {
        if (!active && !exhausted) {
            unsigned item = 0;
            if (gl == 0) item = atomicAdd(na.counter, 1u);
            item = __shfl_sync(gmask, item, gbase);
            if (item >= (unsigned) na.nItems) {
                exhausted = true;
            } else {
                const RegionDev R = a.regions[a.list[item]];
                lX = R.lX;
                lY = R.lY;
                N = R.lX + R.lY;
                dg = a.diags + R.diagBase;
                sx = a.symX + R.xBase;
                sy = a.symY + R.yBase;
                pf = a.planesF + R.cellBase;
                aux = a.aux + R.auxBase;
                d = 0;
                active = true;
                R0 = dg[0];
                R1 = dg[1];
                R2 = dg[2];
                cXn = (gl == 0 && lX > 0) ? sx[0] : 4; /* diagonal 0 is the cell (0, 0) */
                cYn = (gl == 0 && lY > 0) ? sy[0] : 4;
                /* The start vector (impl/pairwiseAligner.c:776-777) enters as what two imaginary diagonals would have handed to the
                 * cell (0,0): its gap-X states from lane 0 and gap-Y states from lane 1 of a diagonal "-1" with left edge -1, its
                 * match state from lane 0 of a diagonal "-2".  The values pass through the shuffles untouched. */
                const double *sv = R.raggedL ? tab.rstartv : tab.startv;
                clear(M0);
                clear(M1);
                clear(M2);
#pragma unroll
                for (int k = 0; k < NG; k++) {
                    M2.g[k] = gl == 0 ? sv[NarrowShape<S>::gap_x(k)] : CPB_NEG_INF;
                    M2.u[k] = gl == 1 ? sv[NarrowShape<S>::gap_y(k)] : CPB_NEG_INF;
                }
                M2.xmyL = -1;
                M2.w = 2;
                M1.m = gl == 0 ? sv[0] : CPB_NEG_INF;
                M1.xmyL = 0;
                M1.w = 1;
            }
        }
        if (__all_sync(0xFFFFFFFFu, !active)) break;
        step(R0, R1, M2, M1, M0);
        step(R1, R2, M0, M2, M1);
        step(R2, R0, M1, M0, M2);
    }
}

/* ---------------------------------------------------------------------------------------------
 * k_backward_narrow<S, NP, ZSUM, G, WPC>: one group of G lanes per traceback block, diagonals top .. T+1, gather form
 * (cell_backward).  ZSUM: the planes written are F + B per state; otherwise raw B (expectations).
 * ------------------------------------------------------------------------------------------- */
template <int S> struct NarrowBwdSet {
    double b[S];  /* backward states of this lane's cell */
    int xmyL, w;  /* left edge and width of the diagonal */
};

template <int S, int NP, bool ZSUM, int G, int WPC>
__global__ void __launch_bounds__(32 * WPC, CPB_STRIP_MIN_BLOCKS) k_backward_narrow(const DpArgs a, const CpbModel model, const NarrowArgs na) {
    __shared__ __align__(128) StripTables<S> tab;
    fill_strip_tables<S>(tab, model, threadIdx.x, 32 * WPC);
    __syncthreads();
    constexpr int NG = NarrowShape<S>::NG, NL = Shape<S>::NL, NM = Shape<S>::NM, NU = Shape<S>::NU;
    constexpr int NFM = ZSUM && NP > 0 ? NP : 1;
    const LaTable la = logadd_lane_table(tab.la);
    const int l16 = threadIdx.x & 15, lane = threadIdx.x & 31, gl = lane & (G - 1), gbase = lane & ~(G - 1);
    const unsigned gmask = G == 32 ? 0xFFFFFFFFu : (((1u << G) - 1u) << gbase);
    const int nF = a.auxF;

    bool active = false, exhausted = false;
    int d = 0, T = 0, top = 0, from = 0, lX = 0, lY = 0;
    const double *endVec = tab.endv;
    const DiagRec *dg = nullptr;
    const uint8_t *sx = nullptr, *sy = nullptr;
    const double *pf = nullptr;
    double *pb = nullptr, *aux = nullptr;
    /* records of d, d-1, d-2 in flight; the symbols and F values of a diagonal's cells are fetched one step ahead */
    DiagRec R0, R1, R2;
    R0.xmyL = R1.xmyL = R2.xmyL = 0;
    R0.width = R1.width = R2.width = 0;
    R0.coff = R1.coff = R2.coff = 0;
    R0.aoff = R1.aoff = R2.aoff = NO_AUX;
    int cXn = 4, cYn = 4;
    double fNext[NFM];
#pragma unroll
    for (int k = 0; k < NFM; k++) fNext[k] = 0.0;
    NarrowBwdSet<S> B0, B1, B2; /* the cells of d, d+1, d+2, rotating through fixed registers (three steps per loop iteration) */
    auto clear = [&](NarrowBwdSet<S> &q) {
#pragma unroll
        for (int k = 0; k < S; k++) q.b[k] = CPB_NEG_INF;
        q.xmyL = 0;
        q.w = 0;
    };
    clear(B0);
    clear(B1);
    clear(B2);

    /* one diagonal: `rec` is its record (replaced by the one three diagonals down), `recNext` the record of d-1 */
    auto step = [&](DiagRec &rec, const DiagRec &recNext, const NarrowBwdSet<S> &in1, const NarrowBwdSet<S> &in2, NarrowBwdSet<S> &out) {
        const int width = active ? rec.width : 0;
        const bool valid = gl < width;
        const int xmy = rec.xmyL + 2 * gl;
        /* (x, y+1) and (x+1, y) on d+1, (x+1, y+1) on d+2 */
        const int iA = (xmy - 1 - in1.xmyL) >> 1, iB = iA + 1, iM = (xmy - in2.xmyL) >> 1;
        double cell[S];
        {
            double u[S], l[NG];
#pragma unroll
            for (int k = 0; k < S; k++) u[k] = CPB_NEG_INF;
            const bool okA = valid && (unsigned) iA < (unsigned) in1.w, okB = valid && (unsigned) iB < (unsigned) in1.w;
            const bool okM = valid && (unsigned) iM < (unsigned) in2.w;
#pragma unroll
            for (int k = 0; k < NG; k++) {
                const double vu = __shfl_sync(0xFFFFFFFFu, in1.b[NarrowShape<S>::gap_y(k)], gbase + (iA & (G - 1)));
                const double vl = __shfl_sync(0xFFFFFFFFu, in1.b[NarrowShape<S>::gap_x(k)], gbase + (iB & (G - 1)));
                u[NarrowShape<S>::gap_y(k)] = okA ? vu : CPB_NEG_INF;
                l[k] = okB ? vl : CPB_NEG_INF;
            }
            const double vm = __shfl_sync(0xFFFFFFFFu, in2.b[0], gbase + (iM & (G - 1)));
            const double t2m = okM ? vm : CPB_NEG_INF;
            double tl[NL], tm[NM], tu[NU];
            load_row<NL>(tl, tab.tl[cXn]);
            const double eM = tab.eM[cXn * 6 + cYn][l16], eY = tab.eY[cYn][l16];
#pragma unroll
            for (int k = 0; k < NM; k++) tm[k] = eM + model.tMiddle[k];
#pragma unroll
            for (int k = 0; k < NU; k++) tu[k] = eY + model.tUpper[k];
            cell_backward<S>(cell, t2m, u, l, tm, tu, tl, la);
        }
        if (d == top) {
            /* the block's top diagonal holds the end vector (impl/pairwiseAligner.c:798-799) */
#pragma unroll
            for (int k = 0; k < S; k++) cell[k] = valid ? endVec[k] : CPB_NEG_INF;
        }
        if (valid) {
            const int c = (int) rec.coff + gl;
            double fm[NFM];
#pragma unroll
            for (int k = 0; k < NFM; k++) fm[k] = fNext[k];
            const bool owned = d <= from;
            const bool feeds = d - 1 > T && d - 1 <= from && recNext.aoff != NO_AUX; /* d-1 is a total diagonal */
            double f0 = fm[0];
            if (!ZSUM && feeds) f0 = pf[c];
            if (owned) {
#pragma unroll
                for (int k = 0; k < NP; k++) pb[(int64_t) k * a.planeStride + c] = ZSUM ? fm[k] + cell[k] : cell[k];
                if (rec.aoff != NO_AUX) {
                    /* cell_dotProduct(F[d], B[d]) (impl/pairwiseAligner.c:402-408); folded over the cells by k_totals */
                    double f[S];
                    if (nF != 0) {
#pragma unroll
                        for (int k = 0; k < S; k++) f[k] = aux[(size_t) rec.aoff + (size_t) k * rec.width + gl];
                    } else {
#pragma unroll
                        for (int k = 0; k < S; k++) f[k] = pf[(int64_t) k * a.planeStride + c];
                    }
                    double t = f[0] + cell[0];
#pragma unroll
                    for (int k = 1; k < S; k++) t = log_add(t, f[k] + cell[k], la);
                    aux[(size_t) rec.aoff + (size_t) nF * rec.width + gl] = t;
                }
            }
            /* diagonal d-1 is a total diagonal: its second term is the fold of F[d].M + B[d].M (:643-651) */
            if (feeds) aux[(size_t) recNext.aoff + (size_t) (nF + 1) * recNext.width + gl] = f0 + cell[0];
        }
#pragma unroll
        for (int k = 0; k < S; k++) out.b[k] = cell[k];
        out.xmyL = rec.xmyL;
        out.w = width;
        if (active) {
            if (--d <= T) {
                active = false;
            } else {
                /* symbols and F values for the cells of the next diagonal, and the record three diagonals down */
                const int xmyN = recNext.xmyL + 2 * gl, xN = (d + xmyN) >> 1, yN = (d - xmyN) >> 1;
                const bool validN = gl < recNext.width;
                cXn = (validN && xN < lX) ? sx[xN] : 4;
                cYn = (validN && yN < lY) ? sy[yN] : 4;
                if (ZSUM && NP > 0 && validN) {
#pragma unroll
                    for (int k = 0; k < NFM; k++) fNext[k] = pf[(int64_t) k * a.planeStride + (int) recNext.coff + gl];
                }
                rec = dg[d >= 2 ? d - 2 : 0];
            }
        }
    };

    for (;;) {
        if (!active && !exhausted) {
            unsigned item = 0;
            if (gl == 0) item = atomicAdd(na.counter, 1u);
            item = __shfl_sync(gmask, item, gbase);
            if (item >= (unsigned) na.nItems) {
                exhausted = true;
            } else {
                const BlockRec K = a.blocks[a.list[item]];
                const RegionDev R = a.regions[K.region];
                lX = R.lX;
                lY = R.lY;
                T = K.T;
                top = K.top;
                from = K.from;
                endVec = (K.atEnd && R.raggedR) ? tab.rendv : tab.endv;
                dg = a.diags + R.diagBase;
                sx = a.symX + R.xBase;
                sy = a.symY + R.yBase;
                pf = a.planesF + R.cellBase;
                pb = a.planesB + R.cellBase;
                aux = a.aux + R.auxBase;
                d = top;
                active = d > T;
                R0 = dg[d];
                R1 = dg[d - 1]; /* d - 1 >= T >= 0 */
                R2 = dg[d >= 2 ? d - 2 : 0];
                clear(B0);
                clear(B1);
                clear(B2);
                const int xmy0 = R0.xmyL + 2 * gl, x0 = (d + xmy0) >> 1, y0 = (d - xmy0) >> 1;
                const bool valid0 = gl < R0.width;
                cXn = (valid0 && x0 < lX) ? sx[x0] : 4;
                cYn = (valid0 && y0 < lY) ? sy[y0] : 4;
                if (ZSUM && NP > 0 && valid0) {
#pragma unroll
                    for (int k = 0; k < NFM; k++) fNext[k] = pf[(int64_t) k * a.planeStride + (int) R0.coff + gl];
                }
            }
        }
        if (__all_sync(0xFFFFFFFFu, !active)) break;
        step(R0, R1, B2, B1, B0);
        step(R1, R2, B0, B2, B1);
        step(R2, R0, B1, B0, B2);
    }
}

} /* namespace cpb */
