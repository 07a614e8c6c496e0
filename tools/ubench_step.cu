// Micro-benchmark of the forward wavefront step with ONE or TWO matrix rows per lane, built from the kernels' own helpers
// (log_add, middle_fold, lower_folds, the strip tables): how many band cells per second a B200 computes when the band logic is
// taken away, as a function of rows per lane and of the register budget (CTAs per SM).  Two rows per lane = 64-row strips with
// lane l on rows 2l and 2l+1: row 2l+1 hears from row 2l inside the lane, one shuffle set and one column symbol per lane and step,
// and two independent dependency chains in one instruction stream.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=false -Iinclude -Icpecan_b200/csrc -o build/ubench_step tools/ubench_step.cu
#include "strip_kernels.cuh"
#include <cstdio>
#include <vector>
using namespace cpb;

template <int ROWS, int MINB>
__global__ void __launch_bounds__(128, MINB) k_step(const CpbModel model, const uint8_t *symX, const uint8_t *symY, double *planes, double *rings,
                                                    int nSteps, int ringMask) {
    __shared__ __align__(128) StripTables<5> tab;
    fill_strip_tables<5>(tab, model, threadIdx.x, 128);
    __syncthreads();
    constexpr int S = 5, NSH = 3;
    const LaTable la = logadd_lane_table(tab.la);
    const int lane = threadIdx.x & 31, l16 = threadIdx.x & 15;
    const int slot = blockIdx.x * 4 + (threadIdx.x >> 5);
    double *ring = rings + (size_t) slot * 2 * (ringMask + 1) * BND_REC;
    const double *ringIn = ring + (size_t) (ringMask + 1) * BND_REC;
    double *pf = planes + (size_t) slot * nSteps * 32 * ROWS;
    const uint8_t *ptrY = symY + (slot & 1023) * 64 + 2048 - ROWS * lane;
    int cXn6[ROWS];
    double tlD[ROWS][4], own[ROWS][S], send[ROWS][NSH], bNext[NSH];
#pragma unroll
    for (int r = 0; r < ROWS; r++) {
        cXn6[r] = symX[(slot * 64 + ROWS * lane + r) & 65535] * 6;
        load_row<4>(tlD[r], tab.tl[cXn6[r] / 6]);
#pragma unroll
        for (int k = 0; k < S; k++) own[r][k] = -1.0 - 0.37 * k - 0.01 * lane;
#pragma unroll
        for (int k = 0; k < NSH; k++) send[r][k] = -2.0 - 0.21 * k;
    }
#pragma unroll
    for (int k = 0; k < NSH; k++) bNext[k] = -3.0;
    int cYprev = ptrY[-1];
#pragma unroll 2
    for (int d = 0; d < nSteps; d++) {
        const int cY = ptrY[d];
        double rcv[NSH];
#pragma unroll
        for (int k = 0; k < NSH; k++) {
            const double v = shfl_up_f64(send[ROWS - 1][k]);
            rcv[k] = lane != 0 ? v : bNext[k];
        }
        if (lane == 0) load_record<NSH>(bNext, ringIn + (size_t) (d & ringMask) * BND_REC);
#pragma unroll
        for (int r = 0; r < ROWS; r++) {
            const int c = r == 0 ? cY : cYprev; /* row 2l+1 is one column behind row 2l on the same diagonal */
            double tmD[5], tu[4];
            const double eM = tab.eM[cXn6[r] + c][l16], eY = tab.eY[c][l16];
#pragma unroll
            for (int k = 0; k < 5; k++) tmD[k] = eM + model.tMiddle[k];
#pragma unroll
            for (int k = 0; k < 4; k++) tu[k] = eY + model.tUpper[k];
            const double mPrev = middle_fold<S>(own[r], tmD, la);
            double out[S];
            out[0] = rcv[0];
            out[1] = rcv[1];
            out[3] = rcv[2];
            out[2] = log_add(own[r][0] + tu[0], own[r][2] + tu[1], la);
            out[4] = log_add(own[r][0] + tu[2], own[r][4] + tu[3], la);
            pf[((size_t) d * ROWS + r) * 32 + lane] = out[0];
#pragma unroll
            for (int k = 0; k < NSH; k++) rcv[k] = send[r][k]; /* what the next row of this lane hears on this diagonal */
            send[r][0] = mPrev;
            lower_folds<S>(send[r] + 1, out, tlD[r], la);
#pragma unroll
            for (int k = 0; k < S; k++) own[r][k] = out[k];
        }
        store_record_if<NSH>(lane == 31, ring + (size_t) (d & ringMask) * BND_REC, send[ROWS - 1]);
        cYprev = cY;
    }
}

template <int ROWS, int MINB> static void run(const char *name, const CpbModel &m, const uint8_t *sx, const uint8_t *sy, double *planes, double *rings, int nSteps, int ringMask) {
    const int grid = 148 * MINB;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int it = 0; it < 4; it++) {
        cudaEventRecord(e0);
        k_step<ROWS, MINB><<<grid, 128>>>(m, sx, sy, planes, rings, nSteps, ringMask);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (it > 0 && ms < best) best = ms;
    }
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, (const void *) k_step<ROWS, MINB>);
    const double cells = (double) grid * 4 * 32 * ROWS * nSteps;
    printf("%-28s regs %3d spill %3zu  warps/SM %2d  chains/SM %2d  %7.2f ms  %6.1f Gcell/s  (%s)\n", name, fa.numRegs, (size_t) fa.localSizeBytes, MINB * 4,
           MINB * 4 * ROWS, best, cells / best * 1e-6, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    CpbModel m = {};
    m.type = CPB_FIVE_STATE;
    m.stateNumber = 5;
    const double tl[4] = {-3.9, -0.35, -9.2, -0.0003}, tm[5] = {-0.03, -1.2, -1.2, -8.1, -8.1}, tu[4] = {-3.9, -0.35, -9.2, -0.0003};
    for (int k = 0; k < 4; k++) m.tLower[k] = tl[k], m.tUpper[k] = tu[k];
    for (int k = 0; k < 5; k++) m.tMiddle[k] = tm[k], m.eGapX[k] = m.eGapY[k] = -1.386 - 0.01 * k;
    for (int i = 0; i < 25; i++) m.eMatch[i] = i / 5 == i % 5 ? -2.1 - 0.01 * (i % 5) : -4.3 - 0.02 * (i % 7);
    const int nSteps = 4096, ringMask = 1023;
    uint8_t *sx, *sy;
    double *planes, *rings;
    std::vector<uint8_t> h(1 << 17);
    unsigned s = 12345;
    for (auto &c : h) { s = s * 1664525u + 1013904223u; c = (s >> 24) & 3; }
    cudaMalloc(&sx, h.size());
    cudaMalloc(&sy, h.size());
    cudaMemcpy(sx, h.data(), h.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(sy, h.data(), h.size(), cudaMemcpyHostToDevice);
    const size_t maxWarps = 148 * 6 * 4;
    cudaMalloc(&planes, maxWarps * nSteps * 64 * sizeof(double));
    cudaMalloc(&rings, maxWarps * 2 * (ringMask + 1) * BND_REC * sizeof(double));
    cudaMemset(rings, 0xC0, maxWarps * 2 * (ringMask + 1) * BND_REC * sizeof(double));
    run<1, 4>("1 row/lane, 4 CTAs/SM", m, sx, sy, planes, rings, nSteps, ringMask);
    run<1, 3>("1 row/lane, 3 CTAs/SM", m, sx, sy, planes, rings, nSteps, ringMask);
    run<1, 5>("1 row/lane, 5 CTAs/SM", m, sx, sy, planes, rings, nSteps, ringMask);
    run<1, 6>("1 row/lane, 6 CTAs/SM", m, sx, sy, planes, rings, nSteps, ringMask);
    run<2, 2>("2 rows/lane, 2 CTAs/SM", m, sx, sy, planes, rings, nSteps, ringMask);
    run<2, 3>("2 rows/lane, 3 CTAs/SM", m, sx, sy, planes, rings, nSteps, ringMask);
    run<2, 4>("2 rows/lane, 4 CTAs/SM", m, sx, sy, planes, rings, nSteps, ringMask);
    run<2, 5>("2 rows/lane, 5 CTAs/SM", m, sx, sy, planes, rings, nSteps, ringMask);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
