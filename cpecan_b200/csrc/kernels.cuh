/*
 * kernels.cuh -- sm_100a device code of the banded pair-HMM forward-backward.
 *
 * Data flow for one chunk of regions (a region = one split sub-problem, see engine.cu):
 *   k_band      per region : anchors -> per-diagonal band records + traceback-block schedule
 *   k_forward   per region : anti-diagonal wavefront, rolling 2-diagonal window in shared memory,
 *                            selected state planes (and full cells of "total" diagonals) -> HBM
 *   k_backward  per block  : same wavefront downwards in gather form, B planes -> HBM,
 *                            per-cell F.B dot products of the total diagonals -> HBM
 *   k_totals    per decade : the reference's left-to-right logAdd folds that give totalProbability
 *   k_posterior per block  : exp(F+B-total) threshold scan, count pass then compacted write pass
 *   k_expect    per block  : expected transition / emission counts
 *
 * Arithmetic contract (bit-compatible with cPecan's C): all state is FP64, logAdd is the reference's
 * 4-segment cubic evaluated with separately rounded multiplies and adds (impl/pairwiseAligner.c:290-307),
 * every transition is from + (eP + tP) (:384,:394) and each cell accumulates its transitions in the
 * order the reference does (impl/stateMachine.c:454-479 / :695-713 forward; scatter order of
 * diagonalCalculationBackward re-expressed as a gather).
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#include "cpecan_b200.h"

namespace cpb {

constexpr uint32_t NO_AUX = 0xFFFFFFFFu;
constexpr int XMY_BIAS = 1 << 30; /* even; makes xmy + bias positive so >>1 is a floor division */

struct __align__(16) DiagRec {
    int32_t xmyL;  /* smallest x-y on this diagonal */
    int32_t width; /* cells on this diagonal, (xmyR-xmyL)/2+1 */
    uint32_t coff; /* region-relative cell offset of the diagonal's first cell */
    uint32_t aoff; /* region-relative offset (in doubles) of the aux record if this is a "total" diagonal, else NO_AUX */
};

struct RegionDev {
    int64_t xBase, yBase;  /* first symbol of the region in the symbol arrays */
    int64_t anchorBase;    /* first anchor triple */
    int64_t diagBase;      /* first DiagRec; lX+lY+2 records */
    int64_t blockBase;     /* first BlockRec slot */
    int64_t cellBase;      /* chunk-relative first cell (host, after planning) */
    int64_t auxBase;       /* chunk-relative first aux double (host, after planning) */
    int64_t maskBase;      /* chunk-relative first keep-mask word (host, after planning) */
    int64_t stripBase;     /* first StripRec of the region: (lX>>5)+1 records */
    int64_t cells;         /* out: band cells */
    int64_t auxDoubles;    /* out */
    int32_t lX, lY;
    int32_t nAnchors;
    int32_t pair;
    int32_t ox, oy;        /* region origin in pair coordinates */
    int32_t blockCap;
    int32_t nBlocks;       /* out */
    int32_t maxW;          /* out: widest diagonal */
    int32_t maxSpan;       /* out: window slots needed (see k_band) */
    int32_t maxStripRange; /* out: longest diagonal range of a 32-row strip */
    int32_t err;           /* out: 0 ok, 1 invalid diagonal, 2 block table overflow */
    uint8_t raggedL, raggedR;
    uint8_t pad_[6];
};

struct StripRec {
    int32_t dFirst, dLast; /* diagonals on which any row of the 32-row strip is inside the band (dLast < dFirst: never) */
};

struct BlockRec {
    int32_t region;
    int32_t top;   /* diagonal the traceback starts from */
    int32_t T;     /* tracedBackTo before this block: owned diagonals are (T, from] */
    int32_t from;  /* tracedBackFrom */
    int32_t maxSpan; /* window slots needed for the diagonals in (T, top] */
    int32_t atEnd;
    int64_t decadeBase; /* chunk-relative index of the block's first decade (host, after planning) */
};

/* Everything the DP kernels need besides the launch list. */
struct DpArgs {
    const RegionDev *regions;
    const BlockRec *blocks;
    const DiagRec *diags;
    const uint8_t *symX, *symY;
    double *planesF;       /* nPlanes planes of planeStride doubles */
    double *planesB;
    double *aux;
    double *totals;        /* per diagonal record (same indexing as diags) */
    int64_t planeStride;
    int32_t nPlanes;       /* 0 forward-only, 1 match, 3 match+gaps, S expectations */
    int32_t auxF;          /* planes of full forward state kept in aux records of total diagonals (S or 0) */
    const int32_t *list;   /* region ids (forward) or block ids (others) */
    double *forwardOut;    /* per region, forward-only mode */
};

/* ---------------------------------------------------------------------------------------------
 * logAdd, bit-compatible with impl/pairwiseAligner.c:290-307
 * ------------------------------------------------------------------------------------------- */
/* coefficient table laid out [segment][a,b,c,k]; the literals are floats promoted to double exactly as in C */
__constant__ double c_coefficients[16] = {
    (double) -0.009350833524763f, (double) 0.130659527668286f, (double) 0.498799810682272f, (double) 0.693203116424741f,
    (double) -0.014532321752540f, (double) 0.139942324101744f, (double) 0.495635523139337f, (double) 0.692140569840976f,
    (double) -0.004605031767994f, (double) 0.063427417320019f, (double) 0.695956496475118f, (double) 0.514272634594009f,
    (double) -0.000458661602210f, (double) 0.009695946122598f, (double) 0.930734667215156f, (double) 0.168037164329057f };

__constant__ float c_coefficients_f32[16] = {
    -0.009350833524763f, 0.130659527668286f, 0.498799810682272f, 0.693203116424741f,
    -0.014532321752540f, 0.139942324101744f, 0.495635523139337f, 0.692140569840976f,
    -0.004605031767994f, 0.063427417320019f, 0.695956496475118f, 0.514272634594009f,
    -0.000458661602210f, 0.009695946122598f, 0.930734667215156f, 0.168037164329057f };

__device__ __forceinline__ void fill_coefficients(double *ctab, int tid) {
#if CPB_COEF_MODE == 1
    if (tid < 16) reinterpret_cast<float *>(ctab)[tid] = c_coefficients_f32[tid];
#else
    if (tid < 16) ctab[tid] = c_coefficients[tid];
#endif
}

#ifndef CPB_COEF_MODE
#define CPB_COEF_MODE 0 /* 0: FP64 table in shared memory; 1: FP32 table in shared memory, widened per use; 2: FP64 table in constant memory */
#endif

__device__ __forceinline__ double log_add(double x, double y, const double *__restrict__ ctab) {
    const double diff = __dsub_rn(x, y);
    const bool xSmaller = __double2hiint(diff) < 0; /* sign of x-y; both -inf gives NaN, handled below */
    const double big = xSmaller ? y : x;
    const double small = xSmaller ? x : y;
    const double d = fabs(diff);
    const int seg = (d > 1.0) + (d > 2.5) + (d > 4.5);
#if CPB_COEF_MODE == 1
    /* the reference's coefficients are float literals: fetch 16 bytes and widen exactly */
    const float4 c4 = *reinterpret_cast<const float4 *>(reinterpret_cast<const float *>(ctab) + 4 * seg);
    const double2 ab = make_double2((double) c4.x, (double) c4.y), ck = make_double2((double) c4.z, (double) c4.w);
#elif CPB_COEF_MODE == 2
    const double2 ab = *reinterpret_cast<const double2 *>(c_coefficients + 4 * seg);
    const double2 ck = *reinterpret_cast<const double2 *>(c_coefficients + 4 * seg + 2);
#else
    const double2 ab = *reinterpret_cast<const double2 *>(ctab + 4 * seg);
    const double2 ck = *reinterpret_cast<const double2 *>(ctab + 4 * seg + 2);
#endif
    double r = __dadd_rn(__dmul_rn(ab.x, d), ab.y);
    r = __dadd_rn(__dmul_rn(r, d), ck.x);
    r = __dadd_rn(__dmul_rn(r, d), ck.y);
    r = __dadd_rn(r, small);
    /* d >= 7.5, d = +inf (one side is LOG_ZERO) and d = NaN (both are) all return the larger operand */
    return (d < 7.5) ? r : big;
}

/* ---------------------------------------------------------------------------------------------
 * per-CTA tables: eP + tP for every (symbol, transition)
 * ------------------------------------------------------------------------------------------- */
template <int S> struct Shape;
template <> struct Shape<5> { static constexpr int NL = 4, NM = 5, NU = 4; };
template <> struct Shape<3> { static constexpr int NL = 3, NM = 3, NU = 3; };

template <int S> struct Tables {
    double ctab[16];
    double tl[5][Shape<S>::NL];      /* [cX][k]  gap-X emission + lower transition */
    double tm[25][Shape<S>::NM];     /* [cX*5+cY][k] */
    double tu[5][Shape<S>::NU];      /* [cY][k] */
    double startv[S], rstartv[S], endv[S], rendv[S];
    double eGapX[5], eGapY[5], eMatch[25]; /* bare emissions: the DP kernels add the (uniform) transition per cell */
};

template <int S> __device__ __forceinline__ void fill_tables(Tables<S> &t, const CpbModel &m, int tid, int nthreads) {
    fill_coefficients(t.ctab, tid);
    for (int i = tid; i < 5 * Shape<S>::NL; i += nthreads) {
        const int c = i / Shape<S>::NL, k = i % Shape<S>::NL;
        t.tl[c][k] = m.eGapX[c] + m.tLower[k];
    }
    for (int i = tid; i < 25 * Shape<S>::NM; i += nthreads) {
        const int c = i / Shape<S>::NM, k = i % Shape<S>::NM;
        t.tm[c][k] = m.eMatch[c] + m.tMiddle[k];
    }
    for (int i = tid; i < 5 * Shape<S>::NU; i += nthreads) {
        const int c = i / Shape<S>::NU, k = i % Shape<S>::NU;
        t.tu[c][k] = m.eGapY[c] + m.tUpper[k];
    }
    if (tid < 5) {
        t.eGapX[tid] = m.eGapX[tid];
        t.eGapY[tid] = m.eGapY[tid];
    }
    if (tid < 25) t.eMatch[tid] = m.eMatch[tid];
    if (tid < S) {
        t.startv[tid] = m.start[tid];
        t.rstartv[tid] = m.raggedStart[tid];
        t.endv[tid] = m.end[tid];
        t.rendv[tid] = m.raggedEnd[tid];
    }
}


/* (from, to) state of the k-th transition of each neighbour group, in the reference's issue order
 * (impl/stateMachine.c:454-479 five-state, :695-713 three-state) */
template <int S> __host__ __device__ constexpr int lower_from(int k) { return S == 5 ? (k == 0 ? 0 : (k == 1 ? 1 : (k == 2 ? 0 : 3))) : k; }
template <int S> __host__ __device__ constexpr int lower_to(int k) { return S == 5 ? (k < 2 ? 1 : 3) : 1; }
template <int S> __host__ __device__ constexpr int middle_from(int k) { return k; }
template <int S> __host__ __device__ constexpr int upper_from(int k) { return S == 5 ? (k == 0 ? 0 : (k == 1 ? 2 : (k == 2 ? 0 : 4))) : (k == 0 ? 0 : (k == 1 ? 2 : 1)); }
template <int S> __host__ __device__ constexpr int upper_to(int k) { return S == 5 ? (k < 2 ? 2 : 4) : 2; }

template <int WARPS> __device__ __forceinline__ void cta_sync() {
    if (WARPS == 1) __syncwarp();
    else __syncthreads();
}

__device__ __forceinline__ int slot_of(int xmy, int mask) { return ((xmy + XMY_BIAS) >> 1) & mask; }

#define CPB_NEG_INF (__longlong_as_double(0xFFF0000000000000LL))

/* ---------------------------------------------------------------------------------------------
 * cell updates
 * ------------------------------------------------------------------------------------------- */
/* forward: lo = cell (x-1,y), mid = (x-1,y-1), up = (x,y-1); all S states each (absent => -inf) */
template <int S>
__device__ __forceinline__ void cell_forward(double *out, const double *lo, const double *mid, const double *up, const double *tl,
                                             const double *tm, const double *tu, const double *ctab) {
    if constexpr (S == 5) {
        out[1] = log_add(lo[0] + tl[0], lo[1] + tl[1], ctab);
        out[3] = log_add(lo[0] + tl[2], lo[3] + tl[3], ctab);
        double m = log_add(mid[0] + tm[0], mid[1] + tm[1], ctab);
        m = log_add(m, mid[2] + tm[2], ctab);
        m = log_add(m, mid[3] + tm[3], ctab);
        out[0] = log_add(m, mid[4] + tm[4], ctab);
        out[2] = log_add(up[0] + tu[0], up[2] + tu[1], ctab);
        out[4] = log_add(up[0] + tu[2], up[4] + tu[3], ctab);
    } else {
        out[1] = log_add(log_add(lo[0] + tl[0], lo[1] + tl[1], ctab), lo[2] + tl[2], ctab);
        out[0] = log_add(log_add(mid[0] + tm[0], mid[1] + tm[1], ctab), mid[2] + tm[2], ctab);
        out[2] = log_add(log_add(up[0] + tu[0], up[2] + tu[1], ctab), up[1] + tu[2], ctab);
    }
}

/* backward, gather form.  For the cell (x,y): t2m = B.M of (x+1,y+1); tu_* = states of (x,y+1), whose upper
 * neighbour is this cell; tl_* = states of (x+1,y), whose lower neighbour is this cell.  tm/tu/tl are the term
 * rows of those three "to" cells.  Accumulation order = the order the reference's scatter visits this cell:
 * middle of diagonal d+2, then upper-of (x-y-1) and lower-of (x-y+1) on diagonal d+1. */
template <int S>
__device__ __forceinline__ void cell_backward(double *out, double t2m, const double *toU, const double *toL, const double *tm,
                                              const double *tu, const double *tl, const double *ctab) {
    if constexpr (S == 5) {
        double m = log_add(t2m + tm[0], toU[2] + tu[0], ctab);
        m = log_add(m, toU[4] + tu[2], ctab);
        m = log_add(m, toL[1] + tl[0], ctab);
        out[0] = log_add(m, toL[3] + tl[2], ctab);
        out[1] = log_add(t2m + tm[1], toL[1] + tl[1], ctab);
        out[2] = log_add(t2m + tm[2], toU[2] + tu[1], ctab);
        out[3] = log_add(t2m + tm[3], toL[3] + tl[3], ctab);
        out[4] = log_add(t2m + tm[4], toU[4] + tu[3], ctab);
    } else {
        out[0] = log_add(log_add(t2m + tm[0], toU[2] + tu[0], ctab), toL[1] + tl[0], ctab);
        out[1] = log_add(log_add(t2m + tm[1], toU[2] + tu[2], ctab), toL[1] + tl[1], ctab);
        out[2] = log_add(log_add(t2m + tm[2], toU[2] + tu[1], ctab), toL[1] + tl[2], ctab);
    }
}

/* ---------------------------------------------------------------------------------------------
 * The rolling window.  Two parity buffers of S x WCAP doubles in shared memory, indexed by
 * slot(x-y) = floor((x-y)/2) mod WCAP.  Diagonal d overwrites diagonal d-2 in place (same parity, a
 * cell and its "middle" neighbour share x-y).  Invariant kept by clear_stale(): a buffer holds the
 * cells of its latest diagonal and LOG_ZERO in every other slot, so neighbours outside the band read
 * as LOG_ZERO without any bounds test (the reference skips NULL neighbours; logAdd with LOG_ZERO is
 * the identity, so the results are the same bits).  WCAP >= the region's maxSpan (k_band) guarantees
 * that no two live cells alias.
 * ------------------------------------------------------------------------------------------- */
template <int S, int WCAP, int NT>
__device__ __forceinline__ void clear_stale(double *buf, int oldL, int oldR, int newL, int newR, int tid) {
    /* cells of [oldL, oldR] (step 2) that are not in [newL, newR] */
    const int leftEnd = min(oldR, newL - 2);
    const int nLeft = leftEnd >= oldL ? ((leftEnd - oldL) >> 1) + 1 : 0;
    const int rightStart = max(oldL, newR + 2);
    const int nRight = oldR >= rightStart ? ((oldR - rightStart) >> 1) + 1 : 0;
    for (int i = tid; i < nLeft + nRight; i += NT) {
        const int xmy = i < nLeft ? oldL + 2 * i : rightStart + 2 * (i - nLeft);
        const int sl = slot_of(xmy, WCAP - 1);
#pragma unroll
        for (int s = 0; s < S; s++) buf[s * WCAP + sl] = CPB_NEG_INF;
    }
}

/* ---------------------------------------------------------------------------------------------
 * k_forward : one CTA (WARPS warps) per region
 * ------------------------------------------------------------------------------------------- */
template <int S, int WCAP, int WARPS>
__global__ void __launch_bounds__(32 * WARPS, 32 / WARPS) k_forward(const DpArgs a, const CpbModel model) {
    extern __shared__ __align__(16) unsigned char smemRaw[];
    Tables<S> &tab = *reinterpret_cast<Tables<S> *>(smemRaw);
    double *win = reinterpret_cast<double *>(smemRaw + ((sizeof(Tables<S>) + 15) & ~size_t(15)));
    constexpr int NT = 32 * WARPS;
    constexpr int mask = WCAP - 1;
    const int tid = threadIdx.x;

    const int regionId = a.list[blockIdx.x];
    const RegionDev R = a.regions[regionId];
    fill_tables<S>(tab, model, tid, NT);
    for (int i = tid; i < 2 * S * WCAP; i += NT) win[i] = CPB_NEG_INF;
    const int N = R.lX + R.lY;
    const DiagRec *dg = a.diags + R.diagBase;
    const uint8_t *sx = a.symX + R.xBase, *sy = a.symY + R.yBase;
    double *pf = a.planesF + R.cellBase;
    double *aux = a.aux + R.auxBase;
    cta_sync<WARPS>();

    /* diagonal 0: the single cell (0,0) holds the start vector (impl/pairwiseAligner.c:776-777) */
    DiagRec prev1 = dg[0];
    if (tid == 0) {
        const double *sv = R.raggedL ? tab.rstartv : tab.startv;
        const int s0 = slot_of(0, mask);
#pragma unroll
        for (int s = 0; s < S; s++) {
            win[(0 * S + s) * WCAP + s0] = sv[s];
            if (s < a.nPlanes) pf[(int64_t) s * a.planeStride] = sv[s];
        }
    }
    int l2 = 1, r2 = -1; /* band of diagonal d-2: empty */
    DiagRec cur = dg[N >= 1 ? 1 : 0];
    cta_sync<WARPS>();

    for (int d = 1; d <= N; d++) {
        const DiagRec nxt = dg[d + 1]; /* prefetch (record N+1 is a sentinel) */
        const int par = d & 1;
        double *wOwn = win + (par * S) * WCAP;               /* holds diagonal d-2, overwritten in place with d */
        const double *wPrev = win + ((par ^ 1) * S) * WCAP;  /* diagonal d-1 */
        const bool fullToAux = a.auxF != 0 && cur.aoff != NO_AUX;
        const int curR = cur.xmyL + 2 * (cur.width - 1);
#pragma unroll 1
        for (int i = tid; i < cur.width; i += NT) {
            const int xmy = cur.xmyL + 2 * i;
            const int x = (d + xmy) >> 1, y = (d - xmy) >> 1;
            const int cX = x > 0 ? sx[x - 1] : 4, cY = y > 0 ? sy[y - 1] : 4;
            const int own = slot_of(xmy, mask);
            const int sl = par ? own : ((own - 1) & mask); /* slot of xmy-1 on diagonal d-1 */
            const int su = par ? ((own + 1) & mask) : own; /* slot of xmy+1 on diagonal d-1 */
            double lo[S], mid[S], up[S], out[S];
#pragma unroll
            for (int s = 0; s < S; s++) {
                lo[s] = wPrev[s * WCAP + sl];
                up[s] = wPrev[s * WCAP + su];
                mid[s] = wOwn[s * WCAP + own];
            }
            /* eP + tP per transition: one emission load per group, the transition comes from the constant bank */
            double tl[Shape<S>::NL], tm[Shape<S>::NM], tu[Shape<S>::NU];
            {
                const double eX = tab.eGapX[cX], eM = tab.eMatch[cX * 5 + cY], eY = tab.eGapY[cY];
#pragma unroll
                for (int k = 0; k < Shape<S>::NL; k++) tl[k] = eX + model.tLower[k];
#pragma unroll
                for (int k = 0; k < Shape<S>::NM; k++) tm[k] = eM + model.tMiddle[k];
#pragma unroll
                for (int k = 0; k < Shape<S>::NU; k++) tu[k] = eY + model.tUpper[k];
            }
            cell_forward<S>(out, lo, mid, up, tl, tm, tu, tab.ctab);
            const int64_t cell = (int64_t) cur.coff + i;
#pragma unroll
            for (int s = 0; s < S; s++) {
                wOwn[s * WCAP + own] = out[s];
                if (s < a.nPlanes) pf[(int64_t) s * a.planeStride + cell] = out[s];
            }
            if (fullToAux) {
#pragma unroll
                for (int s = 0; s < S; s++) aux[(int64_t) cur.aoff + (int64_t) s * cur.width + i] = out[s];
            }
        }
        if (l2 < cur.xmyL || r2 > curR) clear_stale<S, WCAP, NT>(wOwn, l2, r2, cur.xmyL, curR, tid);
        cta_sync<WARPS>();
        l2 = prev1.xmyL;
        r2 = prev1.xmyL + 2 * (prev1.width - 1);
        prev1 = cur;
        cur = nxt;
    }

    /* forward-only (computeForwardProbability, impl/pairwiseAligner.c:879-931): dot of the last cell with the end vector */
    if (a.forwardOut != nullptr && tid == 0) {
        double v = 0.0; /* LOG_ONE for the empty problem */
        if (N > 0) {
            const int own = slot_of(R.lX - R.lY, mask);
            const double *w = win + ((N & 1) * S) * WCAP;
            const double *ev = R.raggedR ? tab.rendv : tab.endv;
            v = w[0 * WCAP + own] + ev[0];
#pragma unroll
            for (int s = 1; s < S; s++) v = log_add(v, w[s * WCAP + own] + ev[s], tab.ctab);
        }
        a.forwardOut[regionId] = v;
    }
}

/* ---------------------------------------------------------------------------------------------
 * k_backward : one CTA per traceback block
 * ------------------------------------------------------------------------------------------- */
template <int S, int WCAP, int WARPS>
__global__ void __launch_bounds__(32 * WARPS, 32 / WARPS) k_backward(const DpArgs a, const CpbModel model) {
    extern __shared__ __align__(16) unsigned char smemRaw[];
    Tables<S> &tab = *reinterpret_cast<Tables<S> *>(smemRaw);
    double *win = reinterpret_cast<double *>(smemRaw + ((sizeof(Tables<S>) + 15) & ~size_t(15)));
    constexpr int NT = 32 * WARPS;
    constexpr int mask = WCAP - 1;
    const int tid = threadIdx.x;

    const BlockRec K = a.blocks[a.list[blockIdx.x]];
    const RegionDev R = a.regions[K.region];
    fill_tables<S>(tab, model, tid, NT);
    for (int i = tid; i < 2 * S * WCAP; i += NT) win[i] = CPB_NEG_INF;
    const DiagRec *dg = a.diags + R.diagBase;
    const uint8_t *sx = a.symX + R.xBase, *sy = a.symY + R.yBase;
    const double *pf = a.planesF + R.cellBase;
    double *pb = a.planesB + R.cellBase;
    double *aux = a.aux + R.auxBase;
    const int top = K.top, T = K.T, from = K.from;
    cta_sync<WARPS>();
    const double *endVec = (K.atEnd && R.raggedR) ? tab.rendv : tab.endv;
    const int nF = a.auxF; /* full-F planes stored in aux (0 in expectation mode) */

    int l2 = 1, r2 = -1; /* band of diagonal d+2 */
    DiagRec prev1 = dg[top];
    DiagRec cur = prev1;
    for (int d = top; d > T; d--) {
        const DiagRec nxt = dg[d - 1]; /* the next diagonal down (d-1 >= T >= 0) */
        const int par = d & 1;
        double *wOwn = win + (par * S) * WCAP;               /* holds B[d+2], overwritten in place with B[d] */
        const double *wNext = win + ((par ^ 1) * S) * WCAP;  /* B[d+1] */
        const bool owned = d <= from;
        const bool isTotal = owned && cur.aoff != NO_AUX;
        const bool feedsTotal = d - 1 > T && d - 1 <= from && nxt.aoff != NO_AUX; /* d-1 is a total diagonal: it needs F.M+B.M of d */
        const int curR = cur.xmyL + 2 * (cur.width - 1);
#pragma unroll 1
        for (int i = tid; i < cur.width; i += NT) {
            const int xmy = cur.xmyL + 2 * i;
            const int own = slot_of(xmy, mask);
            double out[S];
            if (d == top) {
#pragma unroll
                for (int s = 0; s < S; s++) out[s] = endVec[s];
            } else {
                const int x = (d + xmy) >> 1, y = (d - xmy) >> 1;
                const int cX = x < R.lX ? sx[x] : 4, cY = y < R.lY ? sy[y] : 4; /* symbols of row x+1 / column y+1 */
                const int sU = par ? own : ((own - 1) & mask); /* slot of xmy-1 on diagonal d+1: cell (x,y+1) */
                const int sL = par ? ((own + 1) & mask) : own; /* slot of xmy+1 on diagonal d+1: cell (x+1,y) */
                double toU[S], toL[S];
#pragma unroll
                for (int s = 0; s < S; s++) {
                    toU[s] = wNext[s * WCAP + sU];
                    toL[s] = wNext[s * WCAP + sL];
                }
                const double t2m = wOwn[0 * WCAP + own];
                double tl[Shape<S>::NL], tm[Shape<S>::NM], tu[Shape<S>::NU];
                {
                    const double eX = tab.eGapX[cX], eM = tab.eMatch[cX * 5 + cY], eY = tab.eGapY[cY];
#pragma unroll
                    for (int k = 0; k < Shape<S>::NL; k++) tl[k] = eX + model.tLower[k];
#pragma unroll
                    for (int k = 0; k < Shape<S>::NM; k++) tm[k] = eM + model.tMiddle[k];
#pragma unroll
                    for (int k = 0; k < Shape<S>::NU; k++) tu[k] = eY + model.tUpper[k];
                }
                cell_backward<S>(out, t2m, toU, toL, tm, tu, tl, tab.ctab);
            }
            const int64_t cell = (int64_t) cur.coff + i;
#pragma unroll
            for (int s = 0; s < S; s++) {
                wOwn[s * WCAP + own] = out[s];
                if (owned && s < a.nPlanes) pb[(int64_t) s * a.planeStride + cell] = out[s];
            }
            if (isTotal) {
                /* cell_dotProduct(F[d], B[d]) (impl/pairwiseAligner.c:402-408); the fold over cells happens in k_totals */
                double f[S];
                if (nF != 0) {
#pragma unroll
                    for (int s = 0; s < S; s++) f[s] = aux[(int64_t) cur.aoff + (int64_t) s * cur.width + i];
                } else {
#pragma unroll
                    for (int s = 0; s < S; s++) f[s] = pf[(int64_t) s * a.planeStride + cell];
                }
                double t = f[0] + out[0];
#pragma unroll
                for (int s = 1; s < S; s++) t = log_add(t, f[s] + out[s], tab.ctab);
                aux[(int64_t) cur.aoff + (int64_t) nF * cur.width + i] = t;
            }
            if (feedsTotal) {
                /* match step from F[d-2] into diagonal d dotted with B[d] (impl/pairwiseAligner.c:643-651) == F[d].M + B[d].M */
                aux[(int64_t) nxt.aoff + (int64_t) (nF + 1) * nxt.width + i] = pf[cell] + out[0];
            }
        }
        if (l2 < cur.xmyL || r2 > curR) clear_stale<S, WCAP, NT>(wOwn, l2, r2, cur.xmyL, curR, tid);
        cta_sync<WARPS>();
        l2 = prev1.xmyL;
        r2 = prev1.xmyL + 2 * (prev1.width - 1);
        if (d == top) { l2 = 1; r2 = -1; } /* nothing above the top diagonal */
        prev1 = cur;
        cur = nxt;
    }
}

/* ---------------------------------------------------------------------------------------------
 * k_totals : one warp per block, one lane per decade (10 owned diagonals share one totalProbability,
 * impl/pairwiseAligner.c:830-838).  Each fold is the reference's strictly sequential logAdd over the
 * cells of the diagonal (dpDiagonal_dotProduct, :513-523), so lanes run independent folds in parallel.
 * ------------------------------------------------------------------------------------------- */
__global__ void __launch_bounds__(128) k_totals(const DpArgs a, int nBlocks) {
    __shared__ __align__(16) double ctab[16];
    fill_coefficients(ctab, threadIdx.x);
    __syncthreads();
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= nBlocks) return;
    const BlockRec K = a.blocks[a.list[warp]];
    const RegionDev R = a.regions[K.region];
    const DiagRec *dg = a.diags + R.diagBase;
    const double *aux = a.aux + R.auxBase;
    const int nDecades = (K.from - K.T + 9) / 10;
    for (int j = lane; j < nDecades; j += 32) {
        const int dt = K.from - 10 * j;
        const DiagRec rec = dg[dt];
        const double *v = aux + rec.aoff + (int64_t) a.auxF * rec.width;
        double total = CPB_NEG_INF;
        for (int i = 0; i < rec.width; i++) total = log_add(total, v[i], ctab);
        if (dt + 1 <= K.top) {
            const int w2 = dg[dt + 1].width;
            const double *v2 = v + rec.width;
            double t2 = CPB_NEG_INF;
            for (int i = 0; i < w2; i++) t2 = log_add(t2, v2[i], ctab);
            total = log_add(total, t2, ctab);
        }
        a.totals[R.diagBase + dt] = total;
    }
}

/* ---------------------------------------------------------------------------------------------
 * k_posterior : one CTA of POST_WARPS warps per block, one warp per decade at a time.
 *   WRITE == false : evaluates exp(F+B-total) >= threshold for every owned cell (addPosteriorProb,
 *                    impl/pairwiseAligner.c:655-664), stores one ballot word per 32 cells and the
 *                    number of kept cells of the decade;
 *   (device scan of the decade counts)
 *   WRITE == true  : revisits only the set bits and writes (pInt, x, y) at the decade's offset.
 * Decades are numbered in ascending-diagonal order inside a block, so the output of a region is
 * sorted by (x+y, x).
 * ------------------------------------------------------------------------------------------- */
constexpr int POST_WARPS = 8;

struct PostArgs {
    double threshold;
    double logThresholdLo;   /* log(threshold) minus a safety margin: cells below it cannot pass p >= threshold */
    int32_t nLists;          /* 1 or 3 */
    int32_t pad_;
    int64_t nDecades;        /* decades in this chunk (stride between lists in counts/offsets) */
    int64_t maskWords;       /* mask words per list */
    int32_t *counts;         /* [nLists][nDecades] */
    const int64_t *offsets;  /* [nLists][nDecades] exclusive, global (WRITE) */
    uint32_t *masks;         /* [nLists][maskWords] */
    int32_t *out[3];
};

template <bool WRITE>
__global__ void __launch_bounds__(32 * POST_WARPS) k_posterior(const DpArgs a, const PostArgs p) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const BlockRec K = a.blocks[a.list[blockIdx.x]];
    const RegionDev R = a.regions[K.region];
    const DiagRec *dg = a.diags + R.diagBase;
    const double *tot = a.totals + R.diagBase;
    const int nDecades = (K.from - K.T + 9) / 10;
    const unsigned ltMask = (1u << lane) - 1u;
    for (int j = warp; j < nDecades; j += POST_WARPS) {
        const int dt = K.from - 10 * j;                       /* the decade's total diagonal (its highest) */
        const int64_t g = K.decadeBase + (nDecades - 1 - j);  /* ascending-diagonal numbering */
        const double total = tot[dt];
        const int dLow = max(dt - 9, K.T + 1);
        int64_t run[3] = { 0, 0, 0 };
        if (WRITE) {
            for (int l = 0; l < p.nLists; l++) run[l] = p.offsets[(int64_t) l * p.nDecades + g];
        }
        for (int d = dLow; d <= dt; d++) {
            const DiagRec rec = dg[d];
            const int64_t word0 = R.maskBase + (rec.coff >> 5) + d;
            for (int i0 = 0; i0 < rec.width; i0 += 32) {
                const int i = i0 + lane;
                const bool valid = i < rec.width;
                const int xmy = rec.xmyL + 2 * i;
                const int x = (d + xmy) >> 1, y = (d - xmy) >> 1;
                const int64_t cell = R.cellBase + (int64_t) rec.coff + i;
                const int64_t word = word0 + (i0 >> 5);
                for (int l = 0; l < p.nLists; l++) {
                    /* list 0: match (x>0 && y>0); 1: gapX (x>0); 2: gapY (y>0) -- planes 0,1,2 are M, gapX, gapY */
                    unsigned m;
                    bool keep = false;
                    int pInt = 0;
                    if (WRITE) {
                        m = p.masks[(int64_t) l * p.maskWords + word];
                        keep = (m >> lane) & 1u;
                    } else {
                        keep = valid && (l == 0 ? (x > 0 && y > 0) : (l == 1 ? x > 0 : y > 0));
                    }
                    if (keep) {
                        const double z = (a.planesF[(int64_t) l * a.planeStride + cell] + a.planesB[(int64_t) l * a.planeStride + cell]) - total;
                        keep = false;
                        if (z >= p.logThresholdLo) {
                            double pr = exp(z);
                            if (pr >= p.threshold) {
                                keep = true;
                                if (pr > 1.0) pr = 1.0;
                                pInt = (int) floor(pr * (double) CPB_PAIR_ALIGNMENT_PROB_1);
                            }
                        }
                    }
                    if (WRITE) {
                        if (keep) {
                            int32_t *o = p.out[l] + 3 * (run[l] + __popc(m & ltMask));
                            o[0] = pInt;
                            o[1] = x - 1 + R.ox;
                            o[2] = y - 1 + R.oy;
                        }
                    } else {
                        m = __ballot_sync(0xFFFFFFFFu, keep);
                        if (lane == 0) p.masks[(int64_t) l * p.maskWords + word] = m;
                    }
                    run[l] += __popc(m);
                }
            }
        }
        if (!WRITE && lane == 0) {
            for (int l = 0; l < p.nLists; l++) p.counts[(int64_t) l * p.nDecades + g] = (int32_t) run[l];
        }
    }
}

/* ---- exclusive scan of int32 counts into int64 offsets (three small kernels) ---- */
constexpr int SCAN_TILE = 2048;

__global__ void __launch_bounds__(256) k_scan_tiles(const int32_t *counts, int64_t n, int64_t *tileSums) {
    __shared__ int64_t red[256];
    const int64_t base = (int64_t) blockIdx.x * SCAN_TILE;
    int64_t s = 0;
    for (int k = threadIdx.x; k < SCAN_TILE; k += 256) {
        const int64_t i = base + k;
        if (i < n) s += counts[i];
    }
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) tileSums[blockIdx.x] = red[0];
}

/* single thread block: exclusive scan of the tile sums, starting from *startAt; writes the grand total to *totalOut */
__global__ void __launch_bounds__(1) k_scan_sums(int64_t *tileSums, int64_t nTiles, int64_t startAt, int64_t *totalOut) {
    int64_t run = startAt;
    for (int64_t t = 0; t < nTiles; t++) {
        const int64_t v = tileSums[t];
        tileSums[t] = run;
        run += v;
    }
    *totalOut = run;
}

__global__ void __launch_bounds__(256) k_scan_apply(const int32_t *counts, int64_t n, const int64_t *tileOff, int64_t *offsets) {
    /* each thread owns 8 consecutive elements of the tile */
    __shared__ int64_t part[256];
    const int64_t base = (int64_t) blockIdx.x * SCAN_TILE + (int64_t) threadIdx.x * 8;
    int32_t v[8];
    int64_t s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        v[k] = base + k < n ? counts[base + k] : 0;
        s += v[k];
    }
    part[threadIdx.x] = s;
    __syncthreads();
    /* Hillis-Steele inclusive scan over 256 partials */
    for (int o = 1; o < 256; o <<= 1) {
        int64_t t = threadIdx.x >= o ? part[threadIdx.x - o] : 0;
        __syncthreads();
        part[threadIdx.x] += t;
        __syncthreads();
    }
    int64_t run = tileOff[blockIdx.x] + part[threadIdx.x] - s;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        if (base + k < n) offsets[base + k] = run;
        run += v[k];
    }
}

/* per pair: add the kept-cell counts of its decades in this chunk; decades of a pair are contiguous */
__global__ void k_pair_counts(const int32_t *counts, int64_t nDecades, int nLists, const int64_t *pairDecade0, int nPairs, int64_t *pairCounts,
                              int64_t pairStride) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nPairs * nLists) return;
    const int l = i / nPairs, q = i % nPairs;
    int64_t s = 0;
    for (int64_t g = pairDecade0[q]; g < pairDecade0[q + 1]; g++) s += counts[(int64_t) l * nDecades + g];
    pairCounts[(int64_t) l * pairStride + q] += s;
}

/* ---------------------------------------------------------------------------------------------
 * k_expect : one warp per block; expected transition and emission counts
 * (diagonalCalculationExpectations / updateExpectations, impl/pairwiseAligner.c:418-438, :735-746)
 * partial[block][CPB_HMM_LEN(S)]
 * ------------------------------------------------------------------------------------------- */
template <int S>
__global__ void __launch_bounds__(32) k_expect(const DpArgs a, const CpbModel model, double *partials) {
    extern __shared__ __align__(16) unsigned char smemRaw[];
    Tables<S> &tab = *reinterpret_cast<Tables<S> *>(smemRaw);
    double *ePriv = reinterpret_cast<double *>(smemRaw + ((sizeof(Tables<S>) + 15) & ~size_t(15))); /* [S*16][32] lane-private */
    const int lane = threadIdx.x;
    const BlockRec K = a.blocks[a.list[blockIdx.x]];
    const RegionDev R = a.regions[K.region];
    fill_tables<S>(tab, model, lane, 32);
    for (int i = 0; i < S * 16; i++) ePriv[i * 32 + lane] = 0.0;
    __syncwarp();
    const DiagRec *dg = a.diags + R.diagBase;
    const uint8_t *sx = a.symX + R.xBase, *sy = a.symY + R.yBase;
    const double *pf = a.planesF + R.cellBase, *pb = a.planesB + R.cellBase;
    const double *tot = a.totals + R.diagBase;
    constexpr int NL = Shape<S>::NL, NM = Shape<S>::NM, NU = Shape<S>::NU;
    double accT[NL + NM + NU];
#pragma unroll
    for (int k = 0; k < NL + NM + NU; k++) accT[k] = 0.0;
    double likelihood = 0.0;

    for (int d = K.from; d > K.T; d--) {
        const int dt = K.from - 10 * ((K.from - d) / 10);
        const double total = tot[dt];
        likelihood += total; /* once per diagonal (impl/pairwiseAligner.c:743) */
        const DiagRec rec = dg[d];
        const DiagRec rec1 = dg[d - 1];
        const bool haveM = d >= 2 && !(K.T > 0 && d == K.T + 1); /* F[d-2] was freed at the block boundary (:855) */
        DiagRec rec2 = rec1;
        if (d >= 2) rec2 = dg[d - 2];
        for (int i = lane; i < rec.width; i += 32) {
            const int xmy = rec.xmyL + 2 * i;
            const int x = (d + xmy) >> 1, y = (d - xmy) >> 1;
            const int cX = x > 0 ? sx[x - 1] : 4, cY = y > 0 ? sy[y - 1] : 4;
            const int64_t cell = (int64_t) rec.coff + i;
            double b[S];
#pragma unroll
            for (int s = 0; s < S; s++) b[s] = pb[(int64_t) s * a.planeStride + cell];
            const int iL = (xmy - 1 - rec1.xmyL) >> 1, iU = iL + 1; /* xmy-1 and xmy+1 on d-1 (same parity as rec1.xmyL) */
            const bool inL = xmy - 1 >= rec1.xmyL && iL < rec1.width;
            const bool inU = xmy + 1 >= rec1.xmyL && iU < rec1.width;
            const int iM = (xmy - rec2.xmyL) >> 1;
            const bool inM = haveM && xmy >= rec2.xmyL && iM < rec2.width;
            double q[S]; /* per to-state sum for the emission expectation */
#pragma unroll
            for (int s = 0; s < S; s++) q[s] = 0.0;
            const bool emit = cX < 4 && cY < 4;
            if (inL) {
#pragma unroll
                for (int k = 0; k < NL; k++) {
                    const int f = lower_from<S>(k), t = lower_to<S>(k);
                    const double pr = exp(pf[(int64_t) f * a.planeStride + rec1.coff + iL] + b[t] + tab.tl[cX][k] - total);
                    accT[k] += pr;
                    q[t] += pr;
                }
            }
            if (inM) {
#pragma unroll
                for (int k = 0; k < NM; k++) {
                    const int f = middle_from<S>(k);
                    const double pr = exp(pf[(int64_t) f * a.planeStride + rec2.coff + iM] + b[0] + tab.tm[cX * 5 + cY][k] - total);
                    accT[NL + k] += pr;
                    q[0] += pr;
                }
            }
            if (inU) {
#pragma unroll
                for (int k = 0; k < NU; k++) {
                    const int f = upper_from<S>(k), t = upper_to<S>(k);
                    const double pr = exp(pf[(int64_t) f * a.planeStride + rec1.coff + iU] + b[t] + tab.tu[cY][k] - total);
                    accT[NL + NM + k] += pr;
                    q[t] += pr;
                }
            }
            if (emit) {
#pragma unroll
                for (int s = 0; s < S; s++) ePriv[((s * 16) + cX * 4 + cY) * 32 + lane] += q[s];
            }
        }
    }
    /* reduce over lanes in a fixed order and write this block's partial Hmm */
    double *out = partials + (int64_t) blockIdx.x * CPB_HMM_LEN(S);
    auto reduce = [&](double v) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
        return v;
    };
    double tsum[NL + NM + NU];
#pragma unroll
    for (int k = 0; k < NL + NM + NU; k++) tsum[k] = reduce(accT[k]);
    if (lane == 0) {
        for (int i = 0; i < S * S; i++) out[i] = 0.0;
#pragma unroll
        for (int k = 0; k < NL; k++) out[lower_from<S>(k) * S + lower_to<S>(k)] += tsum[k];
#pragma unroll
        for (int k = 0; k < NM; k++) out[middle_from<S>(k) * S + 0] += tsum[NL + k];
#pragma unroll
        for (int k = 0; k < NU; k++) out[upper_from<S>(k) * S + upper_to<S>(k)] += tsum[NL + NM + k];
        out[S * S + S * 16] = likelihood;
    }
    for (int i = 0; i < S * 16; i++) {
        const double v = reduce(ePriv[i * 32 + lane]);
        if (lane == 0) out[S * S + i] = v;
    }
}

/* sum block partials per pair (blocks of a pair are contiguous and in order), one thread per (pair, entry) */
__global__ void k_reduce_pairs(const double *partials, const int64_t *pairBlockOff, int nPairs, int len, double *perPair) {
    const int64_t idx = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t) nPairs * len) return;
    const int pair = (int) (idx / len), e = (int) (idx % len);
    double s = 0.0;
    for (int64_t b = pairBlockOff[pair]; b < pairBlockOff[pair + 1]; b++) s += partials[b * len + e];
    perPair[idx] += s;
}

/* total over pairs in pair order; one thread per entry (len <= 106), sequential => deterministic */
__global__ void k_reduce_total(const double *perPair, int nPairs, int len, double *total) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= len) return;
    double s = 0.0;
    for (int p = 0; p < nPairs; p++) s += perPair[(int64_t) p * len + e];
    total[e] = s;
}

/* ---------------------------------------------------------------------------------------------
 * k_encode : chars -> symbols a,c,g,t,n = 0..4 (symbol_convertCharToSymbol, impl/pairwiseAligner.c:317-334)
 * ------------------------------------------------------------------------------------------- */
__global__ void k_encode(uint8_t *s, int64_t n) {
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t c = s[i] & 0xDF; /* fold case */
    s[i] = c == 'A' ? 0 : (c == 'C' ? 1 : (c == 'G' ? 2 : (c == 'T' ? 3 : 4)));
}

/* ---------------------------------------------------------------------------------------------
 * k_band : one thread per region -- the device-side band builder and traceback scheduler.
 *   band:     band_construct / band_constructDynamic, impl/pairwiseAligner.c:94-234
 *   schedule: the traceback trigger of getPosteriorProbsWithBanding, :791-793, :810, :830, :852
 * ------------------------------------------------------------------------------------------- */
struct BandArgs {
    RegionDev *regions;
    const int32_t *anchors; /* (x,y,expansion) triples, pair coordinates */
    DiagRec *diags;
    BlockRec *blocks;
    StripRec *strips;
    int32_t nRegions;
    int32_t expansion, dynamic;
    int32_t minDiags, traceBack;
    int32_t auxF;       /* full-F planes in the aux record of a total diagonal */
    int32_t scheduleOn; /* 0: forward-only, no blocks */
};

__device__ __forceinline__ int64_t clamp_coord(int64_t z, int64_t l) { return z < 0 ? 0 : (z > l ? l : z); }

__global__ void k_band(const BandArgs b) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= b.nRegions) return;
    RegionDev R = b.regions[r];
    DiagRec *dg = b.diags + R.diagBase;
    const int32_t *an = b.anchors + 3 * R.anchorBase;
    const int64_t lX = R.lX, lY = R.lY, N = lX + lY;
    int64_t ai = 0, pxay = 0, pxmy = 0, nxay = 0, nxmy = 0, xL = 0, yL = 0, xU = 0, yU = 0;
    int64_t e = b.dynamic ? 0 : b.expansion;
    int64_t coff = 0;
    int maxW = 0, maxSpan = 1, err = 0;
    StripRec *st = b.strips + R.stripBase;
    const int nStrips = (int) (lX >> 5) + 1;
    for (int k = 0; k < nStrips; k++) {
        st[k].dFirst = 0x7FFFFFFF;
        st[k].dLast = -1;
    }
    int curLo = 0, curHi = -1; /* strips touched by the previous diagonal */
    int64_t bl1 = 0, br1 = 0, bl2 = 0, br2 = -1; /* bands of the two previous diagonals */
    for (int64_t xay = 0; xay <= N; xay++) {
        int64_t l = xL - yL, rr = xU - yU, v;
        if ((xay + l) % 2 != 0) l += 1;
        if ((xay + rr) % 2 != 0) rr += 1;
        v = (xay + l) / 2; if (v < xL) l += 2 * (xL - v);
        v = (xay - l) / 2; if (yL < v) l += 2 * (v - yL);
        v = (xay + rr) / 2; if (xU < v) rr -= 2 * (v - xU);
        v = (xay - rr) / 2; if (v < yU) rr -= 2 * (yU - v);
        if ((xay + l) % 2 != 0 || (xay + rr) % 2 != 0 || l > rr) {
            err = 1;
            rr = l; /* keep the tables well-formed; the host reports the error */
        }
        const int w = (int) ((rr - l) / 2 + 1);
        DiagRec rec;
        rec.xmyL = (int32_t) l;
        rec.width = w;
        rec.coff = (uint32_t) coff;
        rec.aoff = NO_AUX;
        dg[xay] = rec;
        coff += w;
        maxW = w > maxW ? w : maxW;
        {
            /* row strips touched by this diagonal: rows (xay + l)/2 .. (xay + rr)/2.  Only strips entering or leaving
             * the touched range are written, so the table costs O(changes), not O(cells / 32). */
            int sLo = (int) (((xay + l) / 2) >> 5), sHi = (int) (((xay + rr) / 2) >> 5);
            if (sHi >= nStrips) sHi = nStrips - 1;
            for (int k = sLo; k <= sHi; k++) {
                if (k < curLo || k > curHi) {
                    if (st[k].dFirst > (int) xay) st[k].dFirst = (int) xay;
                }
            }
            for (int k = curLo; k <= curHi; k++) {
                if (k < sLo || k > sHi) {
                    if (st[k].dLast < (int) xay - 1) st[k].dLast = (int) xay - 1;
                }
            }
            curLo = sLo;
            curHi = sHi;
        }
        {
            /* window slots the kernels need around this diagonal: same parity as xay-2, and xay-1 vs the +-1 neighbourhood */
            int span = w;
            if (xay >= 2 && br2 >= bl2) span = (int) (((rr > br2 ? rr : br2) - (l < bl2 ? l : bl2)) / 2 + 1);
            if (xay >= 1) {
                const int64_t hi = br1 > rr + 1 ? br1 : rr + 1, lo = bl1 < l - 1 ? bl1 : l - 1;
                const int sb = (int) ((hi - lo) / 2 + 1);
                span = sb > span ? sb : span;
            }
            dg[xay].aoff = (uint32_t) span; /* parked here until the schedule pass below consumes it */
            maxSpan = span > maxSpan ? span : maxSpan;
            bl2 = bl1; br2 = br1; bl1 = l; br1 = rr;
        }
        if (nxay == xay) {
            pxay = nxay;
            pxmy = nxmy;
            int64_t x = lX, y = lY;
            if (ai < R.nAnchors) {
                x = (int64_t) an[3 * ai] - R.ox + 1;
                y = (int64_t) an[3 * ai + 1] - R.oy + 1;
                if (b.dynamic) e = an[3 * ai + 2];
                ai++;
            }
            nxay = x + y;
            nxmy = x - y;
            xL = clamp_coord((pxay + (pxmy - e)) / 2, lX);
            yL = clamp_coord((nxay - (nxmy - e)) / 2, lY);
            xU = clamp_coord((nxay + (nxmy + e)) / 2, lX);
            yU = clamp_coord((pxay - (pxmy + e)) / 2, lY);
        }
    }
    DiagRec sentinel;
    sentinel.xmyL = 0;
    sentinel.width = 0;
    sentinel.coff = (uint32_t) coff;
    sentinel.aoff = NO_AUX;
    dg[N + 1] = sentinel;

    int nBlocks = 0;
    int64_t auxD = 0;
    /* the per-diagonal spans were parked in aoff; blocks take their maximum, then aoff gets its real meaning */
    if (b.scheduleOn && N > 0) {
        BlockRec *bl = b.blocks + R.blockBase;
        int64_t T = 0;
        for (int64_t d = 1; d <= N; d++) {
            const bool atEnd = d == N;
            const bool tb = d >= T + b.minDiags && dg[d].width <= 2 * b.expansion + 1;
            if (!(atEnd || tb)) continue;
            const int64_t from = d - (atEnd ? 0 : b.traceBack + 1);
            int mw = 0;
            for (int64_t k = T + 1; k <= d; k++) {
                const int sp = (int) dg[k].aoff; /* still the parked span: diagonals above T have no owner yet */
                mw = sp > mw ? sp : mw;
                if (k <= from) dg[k].aoff = NO_AUX;
            }
            if (nBlocks < R.blockCap) {
                BlockRec K;
                K.region = r;
                K.top = (int32_t) d;
                K.T = (int32_t) T;
                K.from = (int32_t) from;
                K.maxSpan = mw;
                K.atEnd = atEnd;
                K.decadeBase = 0;
                bl[nBlocks] = K;
            } else {
                err = 2;
            }
            nBlocks++;
            for (int64_t dt = from; dt > T; dt -= 10) {
                const int w = dg[dt].width;
                dg[dt].aoff = (uint32_t) auxD;
                auxD += (int64_t) (b.auxF + 1) * w + (dt + 1 <= d ? dg[dt + 1].width : 0);
            }
            T = from;
        }
        dg[0].aoff = NO_AUX;
    } else {
        for (int64_t k = 0; k <= N; k++) dg[k].aoff = NO_AUX;
    }
    R.cells = coff;
    R.auxDoubles = auxD;
    R.nBlocks = nBlocks;
    R.maxW = maxW;
    R.maxSpan = maxSpan;
    for (int k = curLo; k <= curHi; k++) st[k].dLast = (int) N; /* strips still touched by the last diagonal */
    int msr = 1;
    for (int k = 0; k < nStrips; k++) {
        const int rg = st[k].dLast - st[k].dFirst + 1;
        msr = rg > msr ? rg : msr;
    }
    R.maxStripRange = msr;
    R.err = err;
    b.regions[r] = R;
}

} /* namespace cpb */
