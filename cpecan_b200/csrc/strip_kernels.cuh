/*
 * strip_kernels.cuh -- barrier-free forward / backward wavefronts: one warp per region (forward) or per
 * traceback block (backward), no shared-memory window, no band-width limit.
 *
 * Lane l of the warp owns matrix row x = 32*s + l of the current row strip s and walks the anti-diagonals
 * d that touch the strip (k_band records [dFirst, dLast] per strip).  On diagonal d it computes the cell
 * (x, d-x).  Cells outside the band are LOG_ZERO (the reference never creates them, and logAdd with
 * LOG_ZERO is the identity, so the arithmetic that remains is the reference's, bit for bit).
 *
 * Forward.  Of the 13 (9) transitions into a cell (impl/stateMachine.c:454-479, :695-713) the "upper" group
 * comes from the lane's own previous cell and stays local.  The "lower" and "middle" groups come from row
 * x-1, i.e. lane l-1: that lane folds them itself -- with the destination cell's emissions, which it knows
 * (the next row's symbol is fixed per strip, the next column's symbol is the one it fetches anyway) -- in
 * the reference's order, and hands over only the finished sums: one message of NSH = 3 (2) doubles per
 * diagonal instead of all S states twice.  A message is {m'(d-1), g'(d)}: everything in it belongs to diagonal d+1,
 * and the long middle fold of cell d-1 is evaluated during step d, where it overlaps the short folds of cell d.
 * Backward (gather form of the reference's scatter) needs B.M of (x+1,y+1) and
 * the gap-X states of (x+1,y): the same message shape, shuffled down.
 *
 * Strips are processed in row order (forward) / reverse row order (backward); the edge lane writes its
 * messages to a per-warp ring in global memory (L2 resident), indexed by diagonal, and the first lane of the
 * next strip prefetches them one step ahead.  Out-of-band masking costs one select per received word: the
 * lane's own terms are masked through a sixth "symbol" whose emissions are LOG_ZERO.
 * Warps fetch work items from a global counter, so long and short regions balance across the chip.
 */
#pragma once
#include "kernels.cuh"

namespace cpb {

constexpr int BND_REC = 4; /* doubles per ring record: up to 3 used, 32-byte aligned */

struct StripArgs {
    const StripRec *strips;   /* per region: (lX>>5)+1 records at RegionDev.stripBase */
    double *boundary;         /* per warp slot: 2 rings of ringSize records */
    const double *negRecord;  /* one record of LOG_ZERO */
    int32_t ringSize;         /* power of two >= longest strip diagonal range + 4 */
    int32_t nItems;
    unsigned int *counter;    /* work-fetch counter (zeroed before the launch) */
    unsigned long long *progress; /* forward teams: per warp slot, (strip serial << 32) | (last published message index + 2) */
    int32_t teamSize;         /* forward teams: warps that share one region (power of two, 1 = no teams) */
    int32_t pad_;
};

/* ---- forward teams: a region's strips are dealt round-robin to the K warps of a team, and strip s consumes the messages
 * of strip s-1 while that strip is still running on the neighbouring warp.  Progress words are monotone per slot. ---- */
__device__ __forceinline__ unsigned long long strip_serial(int iter, int s) { return ((unsigned long long) (iter + 1) << 20) | (unsigned) s; }
constexpr unsigned PROGRESS_DONE = 0x7FFFFFF0u;
#ifndef CPB_TEAM_SLEEP_NS
#define CPB_TEAM_SLEEP_NS 20
#endif
#ifndef CPB_TEAM_PUBLISH_MASK
#define CPB_TEAM_PUBLISH_MASK 7 /* a strip publishes its progress every (mask + 1) diagonals */
#endif
/* Progress words and ring records both live in L2 (records are written with st.cg and read with ld.cg), the point of
 * coherence between SMs.  The producer's release store orders its earlier record stores before the progress word; the
 * consumer polls with relaxed loads and only then issues its (L2) record loads.  No L1 invalidation is needed or wanted:
 * __threadfence() would cost a CCTL.IVALL per call and wipe the L1 lines every other warp of the SM is working from. */
#ifdef CPB_TEAM_DEBUG
__device__ unsigned long long g_teamTrace[64 * 4];
__device__ unsigned long long g_stepTrace[256]; /* strip 1 and 2 of the first region: time of every step */ /* per strip (first region only): slot, start, loop start, end (ns) */
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
__device__ unsigned long long g_teamDebug[8]; /* wait cycles, waits that spun, spin iterations, wait calls, publish cycles, publishes */
#endif
__device__ __forceinline__ void team_publish(unsigned long long *slotWord, unsigned long long serial, unsigned upTo) {
#ifdef CPB_TEAM_DEBUG
    const long long t0 = clock64();
#endif
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(slotWord), "l"((serial << 32) | upTo) : "memory");
#ifdef CPB_TEAM_DEBUG
    atomicAdd(&g_teamDebug[4], (unsigned long long) (clock64() - t0));
    atomicAdd(&g_teamDebug[5], 1ull);
#endif
}
__device__ __forceinline__ unsigned long long team_peek(const unsigned long long *slotWord) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(slotWord) : "memory");
    return v;
}
__device__ __forceinline__ void team_wait(const unsigned long long *slotWord, unsigned long long serial, unsigned upTo) {
    const unsigned long long want = (serial << 32) | upTo;
#ifdef CPB_TEAM_DEBUG
    const long long t0 = clock64();
    unsigned long long spins = 0;
    while (team_peek(slotWord) < want) spins++;
    atomicAdd(&g_teamDebug[0], (unsigned long long) (clock64() - t0));
    atomicAdd(&g_teamDebug[1], spins > 0 ? 1ull : 0ull);
    atomicAdd(&g_teamDebug[2], spins);
    atomicAdd(&g_teamDebug[3], 1ull);
    return;
#endif
    while (team_peek(slotWord) < want) {
#if CPB_TEAM_SLEEP_NS > 0
        __nanosleep(CPB_TEAM_SLEEP_NS);
#endif
    }
}

/* compile-time flags for the step lambdas */
struct Yes { static constexpr bool value = true; };
struct No { static constexpr bool value = false; };

enum { FWD_REGIONS = 0, FWD_TEAMS = 1, FWD_BLOCKS = 2 };

template <int S> struct Msg;
template <> struct Msg<5> { static constexpr int N = 3; };
template <> struct Msg<3> { static constexpr int N = 2; };

/* per-CTA tables; symbol index 5 = "cell outside the band": its emissions are LOG_ZERO.  eM / eY are replicated for the
 * 16 lanes of a half warp (16 x 8 bytes = all 32 banks): 64-bit fetches of different entries never conflict. */
template <int S> struct StripTables {
    double la[LA_TABLE_DOUBLES];
    double eM[36][16];            /* [cX*6+cY][lane&15] match emission */
    double eY[6][16];             /* [cY][lane&15]      gap-Y emission */
    double tl[6][4];              /* [cX][k]            eGapX + tLower[k] (fetched once per strip) */
    double startv[5], rstartv[5], endv[5], rendv[5];
};

template <int S> __device__ __forceinline__ void fill_strip_tables(StripTables<S> &t, const CpbModel &m, int tid, int nthreads) {
    fill_logadd_rows(t.la, tid, nthreads);
    for (int i = tid; i < 36 * 16; i += nthreads) {
        const int c = i / 16, cX = c / 6, cY = c % 6;
        t.eM[c][i % 16] = (cX < 5 && cY < 5) ? m.eMatch[cX * 5 + cY] : CPB_NEG_INF;
    }
    for (int i = tid; i < 6 * 16; i += nthreads) t.eY[i / 16][i % 16] = i / 16 < 5 ? m.eGapY[i / 16] : CPB_NEG_INF;
    for (int i = tid; i < 6 * 4; i += nthreads) {
        const int c = i / 4, k = i % 4;
        t.tl[c][k] = k < Shape<S>::NL ? (c < 5 ? m.eGapX[c] : CPB_NEG_INF) + m.tLower[k] : 0.0;
    }
    if (tid < S) {
        t.startv[tid] = m.start[tid];
        t.rstartv[tid] = m.raggedStart[tid];
        t.endv[tid] = m.end[tid];
        t.rendv[tid] = m.raggedEnd[tid];
    }
}

__device__ __forceinline__ double shfl_up_f64(double v) { return __shfl_up_sync(0xFFFFFFFFu, v, 1); }
__device__ __forceinline__ double shfl_down_f64(double v) { return __shfl_down_sync(0xFFFFFFFFu, v, 1); }

template <int N> __device__ __forceinline__ void load_record(double *v, const double *rec) {
    /* 32-byte aligned record, L2 only (another lane of this warp wrote it) */
    const double2 a = __ldcg(reinterpret_cast<const double2 *>(rec));
    v[0] = a.x;
    v[1] = a.y;
    if (N > 2) v[2] = __ldcg(rec + 2);
}
template <int N> __device__ __forceinline__ void store_record(double *rec, const double *v) {
    __stcg(reinterpret_cast<double2 *>(rec), make_double2(v[0], v[1]));
#ifdef CPB_RECORD_24
    if (N > 2) __stcg(rec + 2, v[2]);
#else
    __stcg(reinterpret_cast<double2 *>(rec) + 1, N > 2 ? make_double2(v[2], v[2]) : make_double2(v[1], v[1])); /* the whole sector, see store_record_if */
#endif
}

/* The edge lane's ring store as PREDICATED instructions.  Written as `if (lane == edge) store_record(...)` the store becomes a branch
 * that ptxas cannot prove reconvergent; every shuffle after it is then guarded by a BRA.DIV whose slow-path re-entry ends the basic
 * block, and the paired step body falls apart into scheduling blocks in which the four-deep middle fold of one step can no longer
 * overlap the other step (tools/sass_sched.py: ~2000 cycles per pair for one warp against ~1000 as one block; forward -12 % on
 * the B200).  The whole 32-byte record is written (the fourth double is padding): a full sector, so L2 never has to read the
 * sector back from HBM to merge a partial write when the line is evicted. */
template <int N> __device__ __forceinline__ void store_record_if(bool edge, double *rec, const double *v) {
    if (N > 2) {
        asm volatile("{ .reg .pred p; setp.ne.s32 p, %0, 0; @p st.global.cg.v2.f64 [%1], {%2, %3}; @p st.global.cg.v2.f64 [%1+16], {%4, %4}; }" ::"r"((int) edge),
                     "l"(rec), "d"(v[0]), "d"(v[1]), "d"(v[2])
                     : "memory");
    } else {
        asm volatile("{ .reg .pred p; setp.ne.s32 p, %0, 0; @p st.global.cg.v2.f64 [%1], {%2, %3}; @p st.global.cg.v2.f64 [%1+16], {%3, %3}; }" ::"r"((int) edge),
                     "l"(rec), "d"(v[0]), "d"(v[1])
                     : "memory");
    }
}

template <int N> __device__ __forceinline__ void load_row(double *v, const double *row) {
    const double2 a = *reinterpret_cast<const double2 *>(row);
    v[0] = a.x;
    v[1] = a.y;
    if (N > 2) {
        const double2 b = *reinterpret_cast<const double2 *>(row + 2);
        v[2] = b.x;
        if (N > 3) v[3] = b.y;
    }
    if (N > 4) v[4] = row[4];
}

/* the folds a source cell performs for the cell below-right (middle group: m) and below (lower group: g), in the
 * reference's transition order; `c` = the source cell's S states */
template <int S> __device__ __forceinline__ double middle_fold(const double *c, const double *tm, const LaTable la) {
    if constexpr (S == 5) {
        double v = log_add(c[0] + tm[0], c[1] + tm[1], la);
        v = log_add(v, c[2] + tm[2], la);
        v = log_add(v, c[3] + tm[3], la);
        return log_add(v, c[4] + tm[4], la);
    } else {
        return log_add(log_add(c[0] + tm[0], c[1] + tm[1], la), c[2] + tm[2], la);
    }
}
template <int S> __device__ __forceinline__ void lower_folds(double *g, const double *c, const double *tl, const LaTable la) {
    if constexpr (S == 5) {
        g[0] = log_add(c[0] + tl[0], c[1] + tl[1], la); /* -> shortGapX */
        g[1] = log_add(c[0] + tl[2], c[3] + tl[3], la); /* -> longGapX */
    } else {
        g[0] = log_add(log_add(c[0] + tl[0], c[1] + tl[1], la), c[2] + tl[2], la); /* -> gapX */
    }
}

/* ---------------------------------------------------------------------------------------------
 * k_forward_strip<S, NP, WPC, MODE>: NP = state planes written to HBM (0 forward-only, 1 match, 3 match+gaps, S all).
 * FWD_REGIONS: every warp fetches whole regions from the work counter (many regions per launch).
 * FWD_TEAMS  : sa.teamSize consecutive warp slots share a region (few, long regions): strips are dealt round-robin and
 *              pipelined through the rings, regions are dealt round-robin to the teams.
 * FWD_BLOCKS : every warp fetches one traceback block and recomputes the forward cells of its diagonals (T, top] from the
 *              checkpoint of diagonals T-1 and T that an earlier plane-less pass over the regions left in a.ckpt (that pass is
 *              this kernel in one of the other two modes with NP = 0 and the checkpoint diagonals marked as "aux" diagonals).
 *              Long regions then offer as many work items as they have blocks, instead of one.
 * ------------------------------------------------------------------------------------------- */
template <int S, int NP, int WPC, int MODE>
__global__ void __launch_bounds__(32 * WPC, CPB_FWD_MIN_BLOCKS) k_forward_strip(const DpArgs a, const CpbModel model, const StripArgs sa) {
    constexpr bool TEAM = MODE == FWD_TEAMS, BLOCKS = MODE == FWD_BLOCKS;
    __shared__ __align__(128) StripTables<S> tab;
    fill_strip_tables<S>(tab, model, threadIdx.x, 32 * WPC);
    __syncthreads();
    constexpr int NSH = Msg<S>::N, NL = Shape<S>::NL, NM = Shape<S>::NM, NU = Shape<S>::NU;
    const LaTable la = logadd_lane_table(tab.la);
    const int l16 = threadIdx.x & 15;
    const int lane = threadIdx.x & 31;
#ifdef CPB_TEAM_SPREAD
    const int slot = TEAM ? (int) ((threadIdx.x >> 5) * gridDim.x + blockIdx.x) : (int) (blockIdx.x * WPC + (threadIdx.x >> 5)); /* consecutive slots on different SMs */
#else
    const int slot = blockIdx.x * WPC + (threadIdx.x >> 5);
#endif
    const int rm = sa.ringSize - 1;
    const size_t ringDoubles = (size_t) sa.ringSize * BND_REC; /* a slot owns two rings */
    const bool keepFull = a.auxF != 0; /* full cells of total diagonals go to the aux records (posterior modes) */
    const int K = TEAM ? sa.teamSize : 1;
    const int member = slot & (K - 1), teamBase = slot - member;
    const int nTeams = (int) ((gridDim.x * WPC) / K);

    for (int iter = 0;; iter++) {
        unsigned item = 0;
        if (TEAM) {
            item = (unsigned) (teamBase / K + iter * nTeams);
            if (teamBase / K >= nTeams) break; /* warps beyond the last whole team idle */
        } else {
            if (lane == 0) item = atomicAdd(sa.counter, 1u);
            item = __shfl_sync(0xFFFFFFFFu, item, 0);
        }
        if (item >= (unsigned) sa.nItems) break;
        if (TEAM && iter > 0) {
            /* the rings are reused: nobody starts a region before the whole team has finished the previous one */
            if (lane < K) team_wait(sa.progress + teamBase + lane, strip_serial(iter - 1, 0xFFFFF), PROGRESS_DONE);
            __syncwarp();
        }
        int regionId = a.list[item];
        int blockT = 0, blockTop = 0x7FFFFFFF; /* FWD_BLOCKS: the diagonals (T, top] of the block */
        int64_t ckBase = 0;
        if (BLOCKS) {
            const BlockRec B = a.blocks[regionId];
            regionId = B.region;
            blockT = B.T;
            blockTop = B.top;
            ckBase = B.ckBase;
        }
        const RegionDev R = a.regions[regionId];
        const int N = R.lX + R.lY;
        const DiagRec *dg = a.diags + R.diagBase;
        const uint8_t *sx = a.symX + R.xBase, *sy = a.symY + R.yBase;
        double *pf = a.planesF + R.cellBase;
        double *aux = a.aux + R.auxBase;
        const StripRec *strips = sa.strips + R.stripBase;
        const int nStrips = (R.lX >> 5) + 1;
        /* The part of a strip this work item computes.  FWD_BLOCKS: steps T+1 .. min(dLast, top); a strip that is under way on
         * diagonal T-1 or T starts from the checkpoint (fromCk) and first writes the message its step T would have left (ring
         * index T): diagonal T+1 of the next strip still hears from a row whose last band cell is on diagonal T-1. */
        auto clipped = [&](StripRec q, bool &fromCk) {
            fromCk = false;
            if (BLOCKS) {
                q.dLast = min(q.dLast, blockTop);
                if (blockT > 0 && q.dFirst <= blockT) {
                    fromCk = q.dLast >= blockT - 1;
                    q.dFirst = blockT + 1;
                    q.dLast = fromCk ? max(q.dLast, blockT) : blockT; /* not fromCk: over before the block starts, empty */
                }
            }
            return q;
        };

        /* FWD_BLOCKS: only the strips that reach diagonal T-1 and have started by `top` take part */
        int sBegin = member, sEnd = nStrips;
        if (BLOCKS && R.stripsSorted) {
            sBegin = blockT > 0 ? first_strip_with_last_at_least(strips, nStrips, blockT - 1) : 0;
            sEnd = last_strip_with_first_at_most(strips, nStrips, blockTop) + 1;
        }
        for (int s = sBegin; s < sEnd; s += K) {
            bool fromCk, prevFromCk;
            const StripRec sr = clipped(strips[s], fromCk);
            /* ring indices the previous strip writes: its diagonals and one flush record */
            int prevFirst = 1, prevLast = 0;
            if (s > 0) {
                const StripRec sp = clipped(strips[s - 1], prevFromCk);
                if (sp.dLast >= sp.dFirst || prevFromCk) {
                    prevFirst = prevFromCk ? blockT : sp.dFirst;
                    prevLast = sp.dLast + 1;
                }
            }
#ifdef CPB_TEAM_DEBUG
            if (TEAM && lane == 0 && item == 0 && s < 64) { g_teamTrace[4 * s] = slot; g_teamTrace[4 * s + 1] = gtime(); }
#endif
            const int producer = teamBase + ((s - 1) & (K - 1));
            const unsigned long long mySerial = strip_serial(iter, s), prodSerial = strip_serial(iter, s - 1);
            unsigned long long *myWord = sa.progress + slot;
            const unsigned long long *prodWord = sa.progress + producer;
            if (sr.dLast < sr.dFirst && !fromCk) {
                if (TEAM && lane == 31) team_publish(myWord, mySerial, PROGRESS_DONE);
                continue;
            }
            if (TEAM && s >= 2 * K && lane == 0) {
                /* this strip overwrites the ring of strip s-2K, whose reader is strip s-2K+1 on the next warp of the team */
                team_wait(sa.progress + teamBase + ((s + 1) & (K - 1)), strip_serial(iter, s - 2 * K + 1), PROGRESS_DONE);
            }
            __syncwarp();
            double *bOut = sa.boundary + ((size_t) slot * 2 + ((s / K) & 1)) * ringDoubles;
            const double *bIn = sa.boundary + ((size_t) producer * 2 + (((s - 1) / K) & 1)) * ringDoubles;
            unsigned published = 0; /* lane 0: how far the producer is known to have got (message index + 2) */
            /* lane 0 only: make sure message t of the previous strip has been published */
            auto await = [&](const int t) {
                if (TEAM && (unsigned) (t + 2) > published) {
                    team_wait(prodWord, prodSerial, (unsigned) (t + 2));
                    const unsigned long long w = team_peek(prodWord);
                    published = (w >> 32) == prodSerial ? (unsigned) w : PROGRESS_DONE; /* a later serial: that strip is finished */
                }
            };
            const int x = 32 * s + lane;
            const int cXn6 = (x < R.lX ? sx[x] : 4) * 6; /* symbol of row x+1, the row this lane's messages go to */
            double tlD[NL];
            load_row<NL>(tlD, tab.tl[cXn6 / 6]);
            const uint8_t *ptrY = sy - x; /* ptrY[d] = symbol of column d+1-x (symbol arrays are padded on both sides) */

            double own[S], send[NSH], bNext[NSH];
#pragma unroll
            for (int k = 0; k < S; k++) own[k] = CPB_NEG_INF;
#pragma unroll
            for (int k = 0; k < NSH; k++) send[k] = bNext[k] = CPB_NEG_INF;

            int d0 = sr.dFirst;
            if (d0 == 0) {
                /* diagonal 0: the single cell (0,0) holds the start vector (impl/pairwiseAligner.c:776-777) */
                if (lane == 0) {
                    const double *sv = R.raggedL ? tab.rstartv : tab.startv;
#pragma unroll
                    for (int k = 0; k < S; k++) {
                        own[k] = sv[k];
                        if (k < NP) pf[(int64_t) k * a.planeStride] = sv[k];
                    }
                }
                lower_folds<S>(send + 1, own, tlD, la); /* the middle fold of this cell is evaluated by step 1 */
                send[0] = CPB_NEG_INF;
                if (lane == 31) store_record<NSH>(bOut, send);
                d0 = 1;
            } else {
                if (BLOCKS && fromCk) {
                    /* the state step T would have left: own = cell (x, T-x), send = {middle fold of cell (x, T-1-x), lower folds of own} */
                    const DiagRec rB = dg[blockT], rA = dg[blockT - 1];
                    const double *ckA = a.ckpt + ckBase, *ckB = ckA + (size_t) S * rA.width; /* [state][cell] of diagonals T-1 and T */
                    const int iA = x - ((blockT - 1 + rA.xmyL) >> 1), iB = x - ((blockT + rB.xmyL) >> 1);
                    const bool inA = (unsigned) iA < (unsigned) rA.width, inB = (unsigned) iB < (unsigned) rB.width;
                    double before[S], tmD[NM];
#pragma unroll
                    for (int k = 0; k < S; k++) {
                        before[k] = inA ? ckA[(size_t) k * rA.width + iA] : CPB_NEG_INF;
                        own[k] = inB ? ckB[(size_t) k * rB.width + iB] : CPB_NEG_INF;
                    }
                    const double eM = tab.eM[cXn6 + ptrY[blockT - 1]][l16]; /* column T - x, the column of step T */
#pragma unroll
                    for (int k = 0; k < NM; k++) tmD[k] = eM + model.tMiddle[k];
                    send[0] = middle_fold<S>(before, tmD, la);
                    lower_folds<S>(send + 1, own, tlD, la);
                    if (lane == 31) store_record<NSH>(bOut + (size_t) (blockT & rm) * BND_REC, send);
                    d0 = blockT + 1;
                }
                if (lane == 0) {
                    /* message for diagonal d0 from row x-1 (the previous strip's last lane) */
                    const int t = d0 - 1;
                    const DiagRec r0 = dg[d0 <= N ? d0 : N];
                    const bool ok = t >= prevFirst && t <= prevLast && (unsigned) (x - ((d0 + r0.xmyL) >> 1)) < (unsigned) r0.width;
                    if (ok) await(t);
                    load_record<NSH>(bNext, ok ? bIn + (size_t) (t & rm) * BND_REC : sa.negRecord);
                }
            }
            /* diagonal records and column symbols are fetched two steps ahead (records N+1, N+2 are sentinels) */
            DiagRec cur = dg[d0 <= N ? d0 : N], nxt = dg[d0 <= N ? d0 + 1 : N + 1];
            int cYnow = ptrY[d0 - 1], cYnext = ptrY[d0]; /* symbols of columns d0 - x and d0 + 1 - x */

            /* One diagonal step.  Reads own = cell of d-1, leaves own = cell of d.  Without AUX the body has no branch (the plane
             * store is predicated), so the two steps of a pair form one scheduling block and the long middle fold of one
             * overlaps the short critical path of the next.  AUX: d is a "total" diagonal whose full cells are kept. */
            auto step = [&](auto auxTag, const int d, const DiagRec &cur, const DiagRec &nxt, const int cYnow) {
                constexpr bool AUX = decltype(auxTag)::value;
                const int i = x - ((d + cur.xmyL) >> 1);
                const bool inBand = (unsigned) i < (unsigned) cur.width;
                const bool useShfl = inBand && lane != 0;
                double rcv[NSH];
#pragma unroll
                for (int k = 0; k < NSH; k++) {
                    const double v = shfl_up_f64(send[k]);
                    rcv[k] = useShfl ? v : bNext[k]; /* lanes other than 0 keep LOG_ZERO in bNext: this is also the band mask */
                }
                {
                    /* lane 0 prefetches the message for diagonal d+1; LOG_ZERO if the previous strip had none or (x, d+1-x) is outside the band */
                    const bool ok = d >= prevFirst && d <= prevLast && (unsigned) (x - ((d + 1 + nxt.xmyL) >> 1)) < (unsigned) nxt.width;
                    const double *rec = ok ? bIn + (size_t) (d & rm) * BND_REC : sa.negRecord;
                    if (lane == 0) {
                        if (TEAM && ok) await(d);
                        load_record<NSH>(bNext, rec);
                    }
                }
                /* the middle fold of the previous cell (own) for (x+1, d-x): its column symbol is this step's cYnow */
                double tmD[NM], tu[NU];
                {
                    const double eM = tab.eM[cXn6 + cYnow][l16], eY = tab.eY[inBand ? cYnow : 5][l16];
#pragma unroll
                    for (int k = 0; k < NM; k++) tmD[k] = eM + model.tMiddle[k];
#pragma unroll
                    for (int k = 0; k < NU; k++) tu[k] = eY + model.tUpper[k];
                }
                const double mPrev = middle_fold<S>(own, tmD, la);
                /* the cell: middle and lower folds arrive finished; the upper group uses this lane's previous cell */
                double out[S];
                out[0] = rcv[0];
                out[1] = rcv[1];
                if constexpr (S == 5) {
                    out[3] = rcv[2];
                    out[2] = log_add(own[0] + tu[0], own[2] + tu[1], la);
                    out[4] = log_add(own[0] + tu[2], own[4] + tu[3], la);
                } else {
                    out[2] = log_add(log_add(own[0] + tu[0], own[2] + tu[1], la), own[1] + tu[2], la);
                }
                if (inBand) {
                    const int cell = (int) cur.coff + i; /* < 2^31 cells per region is enforced on the host */
#pragma unroll
                    for (int k = 0; k < NP; k++) pf[(int64_t) k * a.planeStride + cell] = out[k];
                    if (AUX) {
#pragma unroll
                        for (int k = 0; k < S; k++) aux[(size_t) cur.aoff + (size_t) k * cur.width + i] = out[k];
                    }
                }
                /* message for diagonal d+1: middle fold of cell d-1, lower folds of cell d */
                send[0] = mPrev;
                lower_folds<S>(send + 1, out, tlD, la);
                store_record_if<NSH>(lane == 31, bOut + (size_t) (d & rm) * BND_REC, send);
                if (TEAM && lane == 31 && (d & CPB_TEAM_PUBLISH_MASK) == CPB_TEAM_PUBLISH_MASK) team_publish(myWord, mySerial, (unsigned) (d + 2)); /* messages up to d are out */
#pragma unroll
                for (int k = 0; k < S; k++) own[k] = out[k];
            };

#ifdef CPB_TEAM_DEBUG
            if (TEAM && lane == 0 && item == 0 && s < 64) g_teamTrace[4 * s + 2] = gtime();
#endif
            for (int d = d0; d <= sr.dLast;) {
#ifdef CPB_TEAM_DEBUG
                if (TEAM && lane == 0 && item == 0 && (s == 1 || s == 2) && d - d0 < 128) g_stepTrace[(s - 1) * 128 + d - d0] = gtime();
#endif
                const bool auxA = keepFull && cur.aoff != NO_AUX, auxB = keepFull && nxt.aoff != NO_AUX; /* warp-uniform */
                const DiagRec nxt2 = dg[d + 2];
                const int cY2 = ptrY[d + 1];
                if (d + 1 <= sr.dLast && !auxA && !auxB) {
                    step(No(), d, cur, nxt, cYnow);
                    step(No(), d + 1, nxt, nxt2, cYnext);
                    cur = nxt2;
                    nxt = dg[d + 3];
                    cYnow = cY2;
                    cYnext = ptrY[d + 2];
                    d += 2;
                } else {
                    if (auxA) step(Yes(), d, cur, nxt, cYnow);
                    else step(No(), d, cur, nxt, cYnow);
                    cur = nxt;
                    nxt = nxt2;
                    cYnow = cYnext;
                    cYnext = cY2;
                    d++;
                }
            }
            {
                /* the middle fold of the last cell belongs to diagonal dLast + 2; cYnow is by now the symbol of column dLast + 1 - x */
                double tmD[NM], last[NSH];
                const double eM = tab.eM[cXn6 + cYnow][l16];
#pragma unroll
                for (int k = 0; k < NM; k++) tmD[k] = eM + model.tMiddle[k];
                last[0] = middle_fold<S>(own, tmD, la);
#pragma unroll
                for (int k = 1; k < NSH; k++) last[k] = CPB_NEG_INF;
                if (lane == 31) {
                    store_record<NSH>(bOut + (size_t) ((sr.dLast + 1) & rm) * BND_REC, last);
                    if (TEAM) team_publish(myWord, mySerial, PROGRESS_DONE);
                }
#ifdef CPB_TEAM_DEBUG
                if (TEAM && lane == 0 && item == 0 && s < 64) g_teamTrace[4 * s + 3] = gtime();
#endif
            }
            if (NP == 0 && a.forwardOut != nullptr && s == nStrips - 1 && sr.dLast == N && N > 0 && lane == (R.lX & 31)) {
                /* computeForwardProbability: the last cell dotted with the end vector (impl/pairwiseAligner.c:910-916) */
                const double *ev = R.raggedR ? tab.rendv : tab.endv;
                double v = own[0] + ev[0];
#pragma unroll
                for (int k = 1; k < S; k++) v = log_add(v, own[k] + ev[k], la);
                a.forwardOut[regionId] = CPB_IS_LOG_ZERO(v) ? CPB_TRUE_NEG_INF : v;
            }
            __syncwarp();
        }
        if (NP == 0 && a.forwardOut != nullptr && N == 0 && lane == 0 && member == 0) a.forwardOut[regionId] = 0.0; /* LOG_ONE for the empty problem */
        if (TEAM && lane == 31) team_publish(sa.progress + slot, strip_serial(iter, 0xFFFFF), PROGRESS_DONE); /* this warp is done with the region */
    }
}

/* ---------------------------------------------------------------------------------------------
 * k_backward_strip<S, NP, ZSUM, WPC, MINB> : one warp per traceback block, strips in descending row order.
 * MINB: resident CTAs per SM it is compiled for (the engine picks 6 -- 80 registers -- for ordinary bands, where the 24 warps per SM
 * are worth 8 %, and 3 -- 168 registers, nothing spilled -- for very wide ones, where the 80-register build is 28 % slower).
 * ZSUM: the planes written are F + B per state (what the posterior scan needs); otherwise raw B (expectations).
 * ------------------------------------------------------------------------------------------- */
template <int S> struct BwdShare; /* states of row x+1 that row x needs: M (for the middle step) and the gap-X states */
template <> struct BwdShare<5> {
    __host__ __device__ static constexpr int state(int k) { return k == 0 ? 0 : (k == 1 ? 1 : 3); }
};
template <> struct BwdShare<3> {
    __host__ __device__ static constexpr int state(int k) { return k; }
};

/* backward cell, gather form.  For the cell (x,y): t2m = B.M of (x+1,y+1); u = states of (x,y+1), whose upper
 * neighbour is this cell; l = gap-X states of (x+1,y), whose lower neighbour is this cell.  tm/tu/tl are the term
 * rows of those three "to" cells.  Accumulation order = the order the reference's scatter visits this cell:
 * middle of diagonal d+2, then upper-of (x-y-1) and lower-of (x-y+1) on diagonal d+1. */
template <int S>
__device__ __forceinline__ void cell_backward(double *out, double t2m, const double *u, const double *l, const double *tm, const double *tu,
                                              const double *tl, const LaTable la) {
    if constexpr (S == 5) {
        double m = log_add(t2m + tm[0], u[2] + tu[0], la);
        m = log_add(m, u[4] + tu[2], la);
        m = log_add(m, l[0] + tl[0], la);
        out[0] = log_add(m, l[1] + tl[2], la);
        out[1] = log_add(t2m + tm[1], l[0] + tl[1], la);
        out[2] = log_add(t2m + tm[2], u[2] + tu[1], la);
        out[3] = log_add(t2m + tm[3], l[1] + tl[3], la);
        out[4] = log_add(t2m + tm[4], u[4] + tu[3], la);
    } else {
        out[0] = log_add(log_add(t2m + tm[0], u[2] + tu[0], la), l[0] + tl[0], la);
        out[1] = log_add(log_add(t2m + tm[1], u[2] + tu[2], la), l[0] + tl[1], la);
        out[2] = log_add(log_add(t2m + tm[2], u[2] + tu[1], la), l[0] + tl[2], la);
    }
}

template <int S, int NP, bool ZSUM, int WPC, int MINB>
__global__ void __launch_bounds__(32 * WPC, MINB) k_backward_strip(const DpArgs a, const CpbModel model, const StripArgs sa) {
    __shared__ __align__(128) StripTables<S> tab;
    fill_strip_tables<S>(tab, model, threadIdx.x, 32 * WPC);
    __syncthreads();
    constexpr int NSH = Msg<S>::N, NL = Shape<S>::NL, NM = Shape<S>::NM, NU = Shape<S>::NU;
    const LaTable la = logadd_lane_table(tab.la);
    const int l16 = threadIdx.x & 15;
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
        int prevHi = 0, prevLo = 1; /* ring indices the previously processed (higher) strip wrote */

        int sHigh = nStrips - 1, sLow = 0;
        if (R.stripsSorted) { /* strips that have started by `top` and reach diagonal T+1 */
            sHigh = last_strip_with_first_at_most(strips, nStrips, top);
            sLow = first_strip_with_last_at_least(strips, nStrips, T + 1);
        }
        for (int s = sHigh; s >= sLow; s--) {
            const StripRec sr = strips[s];
            const int dHi = min(sr.dLast, top), dLo = max(sr.dFirst, T + 1);
            if (dHi < dLo) { /* no row of this strip is inside the band on the block's diagonals */
                prevHi = 0;
                prevLo = 1;
                continue;
            }
            const int x = 32 * s + lane;
            const int cXn6 = (x < R.lX ? sx[x] : 4) * 6; /* symbol of row x+1 */
            double tl[NL];
            load_row<NL>(tl, tab.tl[cXn6 / 6]);
            const uint8_t *ptrY = sy - x; /* ptrY[d] = symbol of column d-x+1, the column right of this lane's cell on diagonal d */
            double *bOut = (s & 1) ? ring1 : ring0;
            const double *bIn = (s & 1) ? ring0 : ring1;

            double own[S], send[NSH], bNext[NSH];
#pragma unroll
            for (int k = 0; k < S; k++) own[k] = CPB_NEG_INF;
#pragma unroll
            for (int k = 0; k < NSH; k++) send[k] = bNext[k] = CPB_NEG_INF;

            /* everything a finished cell has to leave in HBM; fm = the F values of the cell (planes 0..NP-1), fetched a step ahead.
             * PLAIN: an owned diagonal that is neither a total diagonal nor the one above one -- only the plane store remains. */
            auto emit = [&](auto plainTag, const int d, const DiagRec &cur, const DiagRec &nxt, const bool inBand, const int i, const double(&out)[S],
                            const double *fm) {
                constexpr bool PLAIN = decltype(plainTag)::value;
                if (!inBand) return;
                const int cell = (int) cur.coff + i;
                if (PLAIN) {
#pragma unroll
                    for (int k = 0; k < NP; k++) pb[(int64_t) k * a.planeStride + cell] = ZSUM ? fm[k] + out[k] : out[k];
                    return;
                }
                const bool owned = d <= from;
                const bool feeds = d - 1 > T && d - 1 <= from && nxt.aoff != NO_AUX; /* d-1 is a total diagonal */
                double f0 = fm[0];
                if (!ZSUM && feeds) f0 = pf[cell];
                if (owned) {
#pragma unroll
                    for (int k = 0; k < NP; k++) pb[(int64_t) k * a.planeStride + cell] = ZSUM ? fm[k] + out[k] : out[k];
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
                        for (int k = 1; k < S; k++) t = log_add(t, f[k] + out[k], la);
                        aux[(size_t) cur.aoff + (size_t) nF * cur.width + i] = t;
                    }
                }
                /* diagonal d-1 is a total diagonal: its second term is the fold of F[d].M + B[d].M (:643-651) */
                if (feeds) aux[(size_t) nxt.aoff + (size_t) (nF + 1) * nxt.width + i] = f0 + out[0];
            };
            /* F of this lane's cell on diagonal `rec` (clamped into the diagonal, so the address is always valid) */
            constexpr int NFM = ZSUM && NP > 0 ? NP : 1;
            auto fetch_f = [&](const int dd, const DiagRec &rec, double *fm) {
                if (ZSUM && NP > 0) {
                    const int ii = min(max(x - ((dd + rec.xmyL) >> 1), 0), rec.width - 1);
#pragma unroll
                    for (int k = 0; k < NFM; k++) fm[k] = pf[(int64_t) k * a.planeStride + (int) rec.coff + ii];
                }
            };

            int d = dHi;
            DiagRec cur = dg[d], nxt = dg[d - 1]; /* d - 1 >= T >= 0 */
            double fNext[NFM];
#pragma unroll
            for (int k = 0; k < NFM; k++) fNext[k] = 0.0;
            if (d == top) {
                /* the block's top diagonal holds the end vector (impl/pairwiseAligner.c:798-799) */
                const int i = x - ((d + cur.xmyL) >> 1);
                const bool inBand = (unsigned) i < (unsigned) cur.width;
                double out[S], fm[NFM];
#pragma unroll
                for (int k = 0; k < NFM; k++) fm[k] = 0.0;
                fetch_f(d, cur, fm);
#pragma unroll
                for (int k = 0; k < S; k++) out[k] = inBand ? endVec[k] : CPB_NEG_INF;
                emit(No(), d, cur, nxt, inBand, i, out, fm);
                send[0] = CPB_NEG_INF;
#pragma unroll
                for (int k = 1; k < NSH; k++) send[k] = out[BwdShare<S>::state(k)];
                if (lane == 0) store_record<NSH>(bOut + (size_t) (d & rm) * BND_REC, send);
                if (lane == 31) {
                    const bool ok = d >= prevLo && d <= prevHi && (unsigned) (x - ((d - 1 + nxt.xmyL) >> 1)) < (unsigned) nxt.width;
                    load_record<NSH>(bNext, ok ? bIn + (size_t) (d & rm) * BND_REC : sa.negRecord);
                }
#pragma unroll
                for (int k = 0; k < S; k++) own[k] = out[k];
                cur = nxt;
                d--;
                nxt = dg[d > 0 ? d - 1 : 0];
            } else if (lane == 31) {
                /* message for diagonal dHi from row x+1 (the strip processed before this one) */
                const int t = d + 1;
                const bool ok = t >= prevLo && t <= prevHi && (unsigned) (x - ((d + cur.xmyL) >> 1)) < (unsigned) cur.width;
                load_record<NSH>(bNext, ok ? bIn + (size_t) (t & rm) * BND_REC : sa.negRecord);
            }
            /* column symbols, diagonal records and F values are fetched ahead of the step that uses them */
            int cYnow = ptrY[d], cYnext = ptrY[d - 1];
            if (d >= dLo) fetch_f(d, cur, fNext);

            /* One diagonal step.  Reads own = cell of d+1, leaves own = cell of d; fetches the F values of the next step's cell.
             * PLAIN steps have no branch, so the two steps of a pair form one scheduling block. */
            auto step = [&](auto plainTag, const int d, const DiagRec &cur, const DiagRec &nxt, const int cYnow) {
                const int i = x - ((d + cur.xmyL) >> 1);
                const bool inBand = (unsigned) i < (unsigned) cur.width;
                const bool useShfl = inBand && lane != 31;
                double fm[NFM];
#pragma unroll
                for (int k = 0; k < NFM; k++) fm[k] = fNext[k];
                fetch_f(d - 1, nxt, fNext);
                double rcv[NSH];
#pragma unroll
                for (int k = 0; k < NSH; k++) {
                    const double v = shfl_down_f64(send[k]);
                    rcv[k] = useShfl ? v : bNext[k];
                }
                {
                    const bool ok = d >= prevLo && d <= prevHi && (unsigned) (x - ((d - 1 + nxt.xmyL) >> 1)) < (unsigned) nxt.width;
                    const double *rec = ok ? bIn + (size_t) (d & rm) * BND_REC : sa.negRecord;
                    if (lane == 31) load_record<NSH>(bNext, rec);
                }
                const int cYeff = inBand ? cYnow : 5;
                double tm[NM], tu[NU], out[S];
                {
                    const double eM = tab.eM[cXn6 + cYeff][l16], eY = tab.eY[cYeff][l16];
#pragma unroll
                    for (int k = 0; k < NM; k++) tm[k] = eM + model.tMiddle[k];
#pragma unroll
                    for (int k = 0; k < NU; k++) tu[k] = eY + model.tUpper[k];
                }
                cell_backward<S>(out, rcv[0], own, rcv + 1, tm, tu, tl, la);
                emit(plainTag, d, cur, nxt, inBand, i, out, fm);
                /* message for diagonal d-1: M of the cell two diagonals up, gap-X states of this one */
                send[0] = own[0];
#pragma unroll
                for (int k = 1; k < NSH; k++) send[k] = out[BwdShare<S>::state(k)];
#ifdef CPB_BWD_PREDICATED_STORE
                store_record_if<NSH>(lane == 0, bOut + (size_t) (d & rm) * BND_REC, send);
#else
                /* (not the predicated form of the forward kernel: compiled for 6 CTAs per SM -- 80 registers -- it spills twice as much,
                 * and the 24 resident warps are worth more here than one scheduling block per step pair: 66.5 against 71-75 ms) */
                if (lane == 0) store_record<NSH>(bOut + (size_t) (d & rm) * BND_REC, send);
#endif
#pragma unroll
                for (int k = 0; k < S; k++) own[k] = out[k];
            };
            /* a diagonal is plain if it is owned, not a total diagonal, and the diagonal below it is not one either */
            auto plain = [&](const int dd, const DiagRec &rec, const DiagRec &below) {
                return dd <= from && rec.aoff == NO_AUX && !(dd - 1 > T && below.aoff != NO_AUX);
            };

            while (d >= dLo) {
                const DiagRec nxt2 = dg[d >= 2 ? d - 2 : 0];
                const int cY2 = ptrY[d - 2];
                const bool plainA = plain(d, cur, nxt); /* warp-uniform */
                if (d - 1 >= dLo && plainA && plain(d - 1, nxt, nxt2)) {
                    step(Yes(), d, cur, nxt, cYnow);
                    step(Yes(), d - 1, nxt, nxt2, cYnext);
                    cur = nxt2;
                    nxt = dg[d >= 3 ? d - 3 : 0];
                    cYnow = cY2;
                    cYnext = ptrY[d - 3];
                    d -= 2;
                } else {
                    if (plainA) step(Yes(), d, cur, nxt, cYnow);
                    else step(No(), d, cur, nxt, cYnow);
                    cur = nxt;
                    nxt = nxt2;
                    cYnow = cYnext;
                    cYnext = cY2;
                    d--;
                }
            }
            if (lane == 0) {
                /* M of the last diagonal is the middle neighbour of diagonal dLo - 2 */
                double last[NSH];
                last[0] = own[0];
#pragma unroll
                for (int k = 1; k < NSH; k++) last[k] = CPB_NEG_INF;
                store_record<NSH>(bOut + (size_t) ((dLo - 1) & rm) * BND_REC, last);
            }
            prevHi = dHi;
            prevLo = dLo - 1;
            __syncwarp();
        }
    }
}

} /* namespace cpb */
