/*
 * model.c -- host side (plain C) of the pair-HMM model and alignment parameters.
 *
 * Builds the flat CpbModel image the kernels consume.  Behaviour follows cPecan's
 * impl/stateMachine.c (defaults :484-501 / :718-726, emission defaults :269-292, Hmm loaders
 * :529-620 / :747-789, N handling :351-366, start/end vectors :401-448 / :648-687) and
 * impl/pairwiseAligner.c (parameter defaults :1334-1348, split points :1206-1257).
 */
#include <math.h>
#include <string.h>

#include "cpecan_b200.h"
#include "internal.h"

enum { ST_M = 0, ST_GX = 1, ST_GY = 2, ST_LX = 3, ST_LY = 4 };

void cpb_params_default(CpbParams *p) {
    memset(p, 0, sizeof(*p));
    p->threshold = 0.01;
    p->minDiagsBetweenTraceBack = 1000;
    p->traceBackDiagonals = 40;
    p->diagonalExpansion = 20;
    p->constraintDiagonalTrim = 14;
    p->anchorMatrixBiggerThanThis = 500 * 500;
    p->repeatMaskMatrixBiggerThanThis = 500 * 500;
    p->splitMatrixBiggerThanThis = (int64_t) 3000 * 3000;
    p->alignAmbiguityCharacters = 0;
    p->gapGamma = 0.5f;
    p->dynamicAnchorExpansion = 0;
}

/* One gap "lane" (X or Y) of a state machine in log space. */
typedef struct {
    double toMatchShort, toMatchLong; /* gap -> match */
    double openShort, extendShort;
    double openLong, extendLong;
    double switchTo; /* three-state only: other gap -> this gap */
} GapLane;

static const double LOG_N_GAP = -1.386294361;   /* the reference's literal for log(0.25) */
static const double LOG_N_MATCH = -2.772588722; /* the reference's literal for log(0.0625) */

static void fill_emissions(CpbModel *m, const double match4x4[16], const double gapX[4], const double gapY[4]) {
    for (int x = 0; x < 5; x++) {
        for (int y = 0; y < 5; y++) {
            m->eMatch[x * 5 + y] = (x == 4 || y == 4) ? LOG_N_MATCH : match4x4[x * 4 + y];
        }
        m->eGapX[x] = x == 4 ? LOG_N_GAP : gapX[x];
        m->eGapY[x] = x == 4 ? LOG_N_GAP : gapY[x];
    }
}

static void default_emission_tables(double match[16], double gap[4]) {
    const double same = -2.1149196655034745, transversion = -4.5691014376830479, transition = -3.9833860032220842;
    for (int x = 0; x < 4; x++) {
        for (int y = 0; y < 4; y++) {
            /* a<->g and c<->t are transitions: symbols two apart */
            match[x * 4 + y] = x == y ? same : (((x ^ y) == 2) ? transition : transversion);
        }
        gap[x] = -1.6094379124341003;
    }
}

static void assemble(CpbModel *m, int type, double matchContinue, const GapLane *gx, const GapLane *gy) {
    const double ninf = -INFINITY;
    m->type = type;
    if (type == CPB_FIVE_STATE || type == CPB_FIVE_STATE_ASYMMETRIC) {
        m->stateNumber = 5;
        const double tl[4] = { gx->openShort, gx->extendShort, gx->openLong, gx->extendLong };
        const double tm[5] = { matchContinue, gx->toMatchShort, gy->toMatchShort, gx->toMatchLong, gy->toMatchLong };
        const double tu[4] = { gy->openShort, gy->extendShort, gy->openLong, gy->extendLong };
        memcpy(m->tLower, tl, sizeof(tl));
        memcpy(m->tMiddle, tm, sizeof(tm));
        memcpy(m->tUpper, tu, sizeof(tu));
        const double start[5] = { 0.0, ninf, ninf, ninf, ninf };
        const double rstart[5] = { ninf, ninf, ninf, 0.0, 0.0 };
        const double rend[5] = { gx->openLong, gx->openLong, gy->openLong, gx->extendLong, gy->extendLong };
        memcpy(m->start, start, sizeof(start));
        memcpy(m->raggedStart, rstart, sizeof(rstart));
        memcpy(m->end, tm, sizeof(tm)); /* ending is "as if going to a match" */
        memcpy(m->raggedEnd, rend, sizeof(rend));
    } else {
        m->stateNumber = 3;
        const double tl[4] = { gx->openShort, gx->extendShort, gx->switchTo, 0.0 };
        const double tm[5] = { matchContinue, gx->toMatchShort, gy->toMatchShort, 0.0, 0.0 };
        const double tu[4] = { gy->openShort, gy->extendShort, gy->switchTo, 0.0 };
        memcpy(m->tLower, tl, sizeof(tl));
        memcpy(m->tMiddle, tm, sizeof(tm));
        memcpy(m->tUpper, tu, sizeof(tu));
        const double start[5] = { 0.0, ninf, ninf, ninf, ninf };
        const double rstart[5] = { ninf, 0.0, 0.0, ninf, ninf };
        const double end[5] = { matchContinue, gx->toMatchShort, gy->toMatchShort, ninf, ninf };
        const double rend[5] = { (gx->openShort + gy->openShort) / 2.0, gx->extendShort, gy->extendShort, ninf, ninf };
        memcpy(m->start, start, sizeof(start));
        memcpy(m->raggedStart, rstart, sizeof(rstart));
        memcpy(m->end, end, sizeof(end));
        memcpy(m->raggedEnd, rend, sizeof(rend));
    }
}

static int valid_type(int type) { return type >= CPB_FIVE_STATE && type <= CPB_THREE_STATE_ASYMMETRIC; }

int cpb_model_default(int type, CpbModel *m) {
    if (!valid_type(type) || m == NULL) {
        cpb_set_error("cpb_model_default: unrecognised state machine type %d", type);
        return CPB_ERR_ARGUMENT;
    }
    memset(m, 0, sizeof(*m));
    GapLane lane;
    lane.toMatchShort = -1.272871422049609;
    lane.toMatchLong = -5.673280173170473;
    lane.extendShort = -0.3388262689231553;
    lane.openLong = -6.30810595366929;
    lane.extendLong = -0.003442492794189331;
    lane.switchTo = -4.910694825551255;
    /* the three-state default gap open differs from the five-state one in the reference */
    lane.openShort = (type == CPB_THREE_STATE || type == CPB_THREE_STATE_ASYMMETRIC) ? -4.21256642 : -4.34381910900448;
    assemble(m, type, -0.030064059121770816, &lane, &lane);
    double match[16], gap[4];
    default_emission_tables(match, gap);
    fill_emissions(m, match, gap, gap);
    return CPB_OK;
}

/* Marginal of the emission matrices of some gap states onto x (rows) and/or y (columns), normalised, in log space. */
static void gap_marginal(const double *emissions, const int *rowStates, int nRow, const int *colStates, int nCol, double out[4]) {
    double acc[4] = { 0.0, 0.0, 0.0, 0.0 };
    for (int k = 0; k < nRow; k++) {
        const double *e = emissions + rowStates[k] * 16;
        for (int x = 0; x < 4; x++) {
            for (int y = 0; y < 4; y++) acc[x] += e[x * 4 + y];
        }
    }
    for (int k = 0; k < nCol; k++) {
        const double *e = emissions + colStates[k] * 16;
        for (int x = 0; x < 4; x++) {
            for (int y = 0; y < 4; y++) acc[y] += e[x * 4 + y];
        }
    }
    double total = 0.0;
    for (int i = 0; i < 4; i++) total += acc[i];
    for (int i = 0; i < 4; i++) out[i] = log(acc[i] / total);
}

static void match_table(const double *emissions, int symmetric, double out[16]) {
    for (int x = 0; x < 4; x++) {
        for (int y = 0; y < 4; y++) {
            if (symmetric && x != y) {
                int lo = x < y ? x : y, hi = x < y ? y : x;
                out[x * 4 + y] = log((emissions[lo * 4 + hi] + emissions[hi * 4 + lo]) / 2.0);
            } else {
                out[x * 4 + y] = log(emissions[x * 4 + y]);
            }
        }
    }
}

/* "long" must be the lane with the larger extend probability; EM can flip them (stateMachine.c:544-550) */
static void order_short_long(GapLane *g, double *switchShort, double *switchLong) {
    if (g->extendShort > g->extendLong) {
        double t;
        t = g->extendShort; g->extendShort = g->extendLong; g->extendLong = t;
        t = g->toMatchShort; g->toMatchShort = g->toMatchLong; g->toMatchLong = t;
        t = g->openShort; g->openShort = g->openLong; g->openLong = t;
        t = *switchShort; *switchShort = *switchLong; *switchLong = t;
    }
}

int cpb_model_from_hmm(int type, const double *T, const double *E, CpbModel *m) {
    if (!valid_type(type) || m == NULL || T == NULL || E == NULL) {
        cpb_set_error("cpb_model_from_hmm: bad argument (type %d)", type);
        return CPB_ERR_ARGUMENT;
    }
    memset(m, 0, sizeof(*m));
    const int five = type == CPB_FIVE_STATE || type == CPB_FIVE_STATE_ASYMMETRIC;
    const int symmetric = type == CPB_FIVE_STATE || type == CPB_THREE_STATE;
    const int S = five ? 5 : 3;
#define P(f, t) T[(f) * S + (t)]
#define AVG(a, b) (((a) + (b)) / 2.0) /* the three-state loader divides by 2.0, the five-state by 2: same double */
    GapLane gx, gy;
    memset(&gx, 0, sizeof(gx));
    memset(&gy, 0, sizeof(gy));
    double matchContinue = log(P(ST_M, ST_M));
    double match[16], gapX[4], gapY[4];
    match_table(E, symmetric, match);
    if (five) {
        double swShortX, swLongX, swShortY, swLongY;
        if (symmetric) {
            gx.toMatchShort = log(AVG(P(ST_GX, ST_M), P(ST_GY, ST_M)));
            gx.toMatchLong = log(AVG(P(ST_LX, ST_M), P(ST_LY, ST_M)));
            gx.openShort = log(AVG(P(ST_M, ST_GX), P(ST_M, ST_GY)));
            gx.extendShort = log(AVG(P(ST_GX, ST_GX), P(ST_GY, ST_GY)));
            swShortX = log(AVG(P(ST_GX, ST_GY), P(ST_GY, ST_GX)));
            gx.openLong = log(AVG(P(ST_M, ST_LX), P(ST_M, ST_LY)));
            gx.extendLong = log(AVG(P(ST_LX, ST_LX), P(ST_LY, ST_LY)));
            swLongX = log(AVG(P(ST_LX, ST_LY), P(ST_LY, ST_LX)));
            order_short_long(&gx, &swShortX, &swLongX);
            gy = gx;
            const int xs[2] = { ST_GX, ST_LX }, ys[2] = { ST_GY, ST_LY };
            gap_marginal(E, xs, 2, ys, 2, gapX);
            memcpy(gapY, gapX, sizeof(gapX));
        } else {
            gx.toMatchShort = log(P(ST_GX, ST_M));
            gx.toMatchLong = log(P(ST_LX, ST_M));
            gx.openShort = log(P(ST_M, ST_GX));
            gx.extendShort = log(P(ST_GX, ST_GX));
            swShortX = log(P(ST_GY, ST_GX));
            gx.openLong = log(P(ST_M, ST_LX));
            gx.extendLong = log(P(ST_LX, ST_LX));
            swLongX = log(P(ST_LY, ST_LX));
            order_short_long(&gx, &swShortX, &swLongX);
            gy.toMatchShort = log(P(ST_GY, ST_M));
            gy.toMatchLong = log(P(ST_LY, ST_M));
            gy.openShort = log(P(ST_M, ST_GY));
            gy.extendShort = log(P(ST_GY, ST_GY));
            swShortY = log(P(ST_GX, ST_GY));
            gy.openLong = log(P(ST_M, ST_LY));
            gy.extendLong = log(P(ST_LY, ST_LY));
            swLongY = log(P(ST_LX, ST_LY));
            order_short_long(&gy, &swShortY, &swLongY);
            const int xs[2] = { ST_GX, ST_LX }, ys[2] = { ST_GY, ST_LY };
            gap_marginal(E, xs, 2, NULL, 0, gapX);
            gap_marginal(E, NULL, 0, ys, 2, gapY);
        }
    } else {
        const int xs[1] = { ST_GX }, ys[1] = { ST_GY };
        if (symmetric) {
            gx.toMatchShort = log(AVG(P(ST_GX, ST_M), P(ST_GY, ST_M)));
            gx.openShort = log(AVG(P(ST_M, ST_GX), P(ST_M, ST_GY)));
            gx.extendShort = log(AVG(P(ST_GX, ST_GX), P(ST_GY, ST_GY)));
            gx.switchTo = log(AVG(P(ST_GY, ST_GX), P(ST_GX, ST_GY)));
            gy = gx;
            gap_marginal(E, xs, 1, ys, 1, gapX);
            memcpy(gapY, gapX, sizeof(gapX));
        } else {
            gx.toMatchShort = log(P(ST_GX, ST_M));
            gx.openShort = log(P(ST_M, ST_GX));
            gx.extendShort = log(P(ST_GX, ST_GX));
            gx.switchTo = log(P(ST_GY, ST_GX));
            gy.toMatchShort = log(P(ST_GY, ST_M));
            gy.openShort = log(P(ST_M, ST_GY));
            gy.extendShort = log(P(ST_GY, ST_GY));
            gy.switchTo = log(P(ST_GX, ST_GY));
            gap_marginal(E, xs, 1, NULL, 0, gapX);
            gap_marginal(E, NULL, 0, ys, 1, gapY);
        }
    }
#undef P
#undef AVG
    assemble(m, type, matchContinue, &gx, &gy);
    fill_emissions(m, match, gapX, gapY);
    return CPB_OK;
}

/* ---- split points (getSplitPoints, impl/pairwiseAligner.c:1206-1257) ---- */
typedef struct {
    int64_t *out4, cap, n;
} RegionSink;

static void sink_region(RegionSink *s, int64_t x1, int64_t y1, int64_t x2, int64_t y2) {
    if (s->n < s->cap) {
        int64_t *q = s->out4 + 4 * s->n;
        q[0] = x1; q[1] = y1; q[2] = x2; q[3] = y2;
    }
    s->n++;
}

int64_t cpb_split_points(const int64_t *anchors, int64_t nAnchors, int64_t lX, int64_t lY, int64_t splitBiggerThan, int raggedLeft,
                         int raggedRight, int64_t *out4, int64_t cap) {
    RegionSink sink = { out4, cap, 0 };
    const int64_t half = (int64_t) sqrt((double) splitBiggerThan);
    int64_t openX = 0, openY = 0; /* start of the region being grown */
    int64_t prevX = 0, prevY = 0; /* one past the previous anchor */
    int lastGapSplit = 0;
    for (int64_t i = 0; i <= nAnchors; i++) {
        const int64_t nextX = i < nAnchors ? anchors[3 * i] : lX, nextY = i < nAnchors ? anchors[3 * i + 1] : lY;
        const int64_t gapX = nextX - prevX, gapY = nextY - prevY;
        lastGapSplit = 0;
        if (gapX * gapY > splitBiggerThan) {
            const int64_t hX = gapX / 2 > half ? half : gapX / 2, hY = gapY / 2 > half ? half : gapY / 2;
            /* with a ragged left end the region before the first anchor is dropped */
            if (!(raggedLeft && i == 0)) sink_region(&sink, openX, openY, prevX + hX, prevY + hY);
            openX = nextX - hX;
            openY = nextY - hY;
            lastGapSplit = 1;
        }
        prevX = nextX + 1;
        prevY = nextY + 1;
    }
    /* with a ragged right end a trailing split drops the final region */
    if (!lastGapSplit || !raggedRight) sink_region(&sink, openX, openY, lX, lY);
    return sink.n;
}
