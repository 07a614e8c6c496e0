// Micro-benchmarks for the instruction mix of the pair-HMM recurrence on sm_100a: dependent-issue latency and
// per-SM throughput of FP64 add/mul, 64-bit selects, shared-memory 128-bit loads and shuffles.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o ubench tools/ubench.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>

#define ITER 4096

template <int MODE> __global__ void lat(double *out, long long *cycles, double a, double b, int sel) {
    __shared__ double tab[512];
    for (int i = threadIdx.x; i < 512; i += blockDim.x) tab[i] = a + i * 1e-9;
    __syncthreads();
    double x = a + threadIdx.x * 1e-6, y = b;
    int idx = threadIdx.x & 15;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < ITER; i++) {
        if (MODE == 0) x = __dadd_rn(x, y);                       // DADD chain
        if (MODE == 1) x = __dmul_rn(x, y);                       // DMUL chain
        if (MODE == 2) x = __dadd_rn(__dmul_rn(x, y), y);         // DMUL+DADD chain
        if (MODE == 3) { x = (__double2hiint(x) < sel) ? y : x; y = __dadd_rn(y, 1e-300); }  // select fed by DADD
        if (MODE == 4) { idx = (int) tab[idx * 2 & 511] & 15; }   // LDS.64 chain (address dependent)
        if (MODE == 5) { x = __shfl_up_sync(0xffffffffu, x, 1); } // 64-bit shuffle chain
        if (MODE == 6) { idx = (idx * 3 + sel) & 1023; }          // IMAD+LOP chain
        if (MODE == 7) { const double2 v = *reinterpret_cast<const double2 *>(tab + ((idx * 2) & 510)); x = __dadd_rn(x, v.x); idx = (__double2loint(x) & 15); } // LDS.128 + DADD + index
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = x + y + idx;
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[MODE] = t1 - t0;
}

// throughput: many independent chains per thread, many warps
template <int MODE> __global__ void thr(double *out, double a, double b) {
    __shared__ double tab[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) tab[i] = a + i * 1e-9;
    __syncthreads();
    double x[8];
    for (int k = 0; k < 8; k++) x[k] = a + threadIdx.x * 1e-6 + k;
    const double *row = tab + 2 * (threadIdx.x & 7);
#pragma unroll 4
    for (int i = 0; i < ITER; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if (MODE == 0) x[k] = __dadd_rn(x[k], b);
            if (MODE == 1) x[k] = __dmul_rn(x[k], b);
            if (MODE == 2) { const double2 v = *reinterpret_cast<const double2 *>(row + 16 * ((i + k) & 63)); x[k] = __dadd_rn(x[k], v.x) + v.y; }
            if (MODE == 3) x[k] = __shfl_up_sync(0xffffffffu, x[k], 1);
        }
    }
    double s = 0;
    for (int k = 0; k < 8; k++) s += x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}


// pipe overlap: does an SMSP issue FP64 (DADD), ALU (SEL) and FMA-pipe (IMAD) instructions concurrently, or do the
// 16-lane pipes serialise at the dispatch port?  Each mode runs 8 independent chains per op class per iteration.
template <int MODE> __global__ void mix(double *out, double a, double b, int sel) {
    double x[8];
    int u[8], v[8];
    for (int k = 0; k < 8; k++) { x[k] = a + threadIdx.x * 1e-6 + k; u[k] = threadIdx.x + k; v[k] = sel + k + 3 * threadIdx.x; }
#pragma unroll 4
    for (int i = 0; i < ITER; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if (MODE & 1) x[k] = __dadd_rn(x[k], b);
            if (MODE & 2) asm volatile("{ .reg .pred p; setp.lt.s32 p, %1, %2; selp.b32 %0, %0, %3, p; }" : "+r"(u[k]) : "r"(sel), "r"(0), "r"(v[k]));
            if (MODE & 4) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(v[k]) : "r"(sel), "r"(3));
        }
    }
    double s = 0;
    for (int k = 0; k < 8; k++) s += x[k] + u[k] + v[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// LDS.128 with a per-lane row that changes every iteration (what logAdd does):
// MODE 0: compact table, 5 rows of 16 bytes in 5 different bank groups (any mix of rows: one address per bank, broadcast);
// MODE 1: lane-replicated rows (row * 128 + 16 * (lane & 7)): conflict-free, but every quarter warp reads its own 128 bytes.
template <int MODE> __global__ void ldsrows(double *out, double a) {
    __shared__ __align__(128) double tab[26 * 16];
    for (int i = threadIdx.x; i < 26 * 16; i += blockDim.x) tab[i] = a + i * 1e-9;
    __syncthreads();
    unsigned base = (unsigned) __cvta_generic_to_shared(tab);
    unsigned h = threadIdx.x * 2654435761u + blockIdx.x;
    double x[4] = {a, a, a, a};
#pragma unroll 4
    for (int i = 0; i < ITER; i++) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
            h = h * 1664525u + 1013904223u;
            const unsigned r = (h >> 24);
            const unsigned addr = MODE == 0 ? base + 16u * (r % 5u) : base + 128u * (r % 26u) + 16u * (threadIdx.x & 7);
            double u, v;
            asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(u), "=d"(v) : "r"(addr));
            x[k] = __dadd_rn(x[k], u);
            x[k] = __dadd_rn(x[k], v);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x[0] + x[1] + x[2] + x[3];
}

int main() {
    double *out;
    long long *cyc, h[16] = {0};
    cudaMalloc(&out, 1 << 24);
    cudaMalloc(&cyc, sizeof(h));
    cudaMemset(cyc, 0, sizeof(h));
    lat<0><<<1, 32>>>(out, cyc, 1.0, 1e-9, 5);
    lat<1><<<1, 32>>>(out, cyc, 1.0, 1.0000001, 5);
    lat<2><<<1, 32>>>(out, cyc, 1.0, 0.999, 5);
    lat<3><<<1, 32>>>(out, cyc, 1.0, 2.0, 5);
    lat<4><<<1, 32>>>(out, cyc, 1.0, 2.0, 5);
    lat<5><<<1, 32>>>(out, cyc, 1.0, 2.0, 5);
    lat<6><<<1, 32>>>(out, cyc, 1.0, 2.0, 5);
    lat<7><<<1, 32>>>(out, cyc, 1.0, 2.0, 5);
    cudaDeviceSynchronize();
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    const char *names[] = { "DADD", "DMUL", "DMUL+DADD", "sel<-DADD", "LDS.64 chain", "SHFL64", "IMAD+LOP", "LDS128+DADD+idx" };
    for (int m = 0; m < 8; m++) printf("latency %-16s %.2f cycles/iter\n", names[m], (double) h[m] / ITER);

    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const char *tn[] = { "DADD", "DMUL", "LDS.128 (conflict-free rows)", "SHFL64" };
    for (int warps = 4; warps <= 16; warps *= 2) {
        for (int m = 0; m < 4; m++) {
            const int blocks = 148 * 4, threads = 32 * warps / 4 * 1; // warps per SM = 4 blocks x (warps/4) warps
            float ms = 0;
            for (int rep = 0; rep < 2; rep++) {
                cudaEventRecord(e0);
                if (m == 0) thr<0><<<blocks, threads>>>(out, 1.0, 1e-9);
                if (m == 1) thr<1><<<blocks, threads>>>(out, 1.0, 1.0000001);
                if (m == 2) thr<2><<<blocks, threads>>>(out, 1.0, 1e-9);
                if (m == 3) thr<3><<<blocks, threads>>>(out, 1.0, 1e-9);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                cudaEventElapsedTime(&ms, e0, e1);
            }
            const double ops = (double) blocks * threads * ITER * 8;
            printf("throughput %-30s warps/SM %2d: %.1f lane-ops/clk/SM (at 1.965 GHz)\n", tn[m], warps, ops / (ms * 1e-3) / 148 / 1.965e9);
        }
    }

    const char *mn[] = { "", "DADD", "SEL", "DADD+SEL", "IMAD", "DADD+IMAD", "SEL+IMAD", "DADD+SEL+IMAD" };
    for (int warps = 4; warps <= 16; warps *= 2) {
        for (int m = 1; m < 8; m++) {
            const int blocks = 148 * 4, threads = 32 * warps / 4;
            float ms = 0;
            for (int rep = 0; rep < 2; rep++) {
                cudaEventRecord(e0);
                if (m == 1) mix<1><<<blocks, threads>>>(out, 1.0, 1e-9, 5);
                if (m == 2) mix<2><<<blocks, threads>>>(out, 1.0, 1e-9, 5);
                if (m == 3) mix<3><<<blocks, threads>>>(out, 1.0, 1e-9, 5);
                if (m == 4) mix<4><<<blocks, threads>>>(out, 1.0, 1e-9, 5);
                if (m == 5) mix<5><<<blocks, threads>>>(out, 1.0, 1e-9, 5);
                if (m == 6) mix<6><<<blocks, threads>>>(out, 1.0, 1e-9, 5);
                if (m == 7) mix<7><<<blocks, threads>>>(out, 1.0, 1e-9, 5);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                cudaEventElapsedTime(&ms, e0, e1);
            }
            // cycles one SMSP spends per (8 ops of each class), all its warps taken together
            const double cyc = ms * 1e-3 * 1.965e9 / ITER / (warps / 4.0);
            printf("mix %-14s warps/SM %2d: %.1f SMSP-cycles per 8 warp-instructions of each class\n", mn[m], warps, cyc);
        }
    }

    for (int warps = 4; warps <= 16; warps *= 2) {
        for (int m = 0; m < 2; m++) {
            const int blocks = 148 * 4, threads = 32 * warps / 4;
            float ms = 0;
            for (int rep = 0; rep < 2; rep++) {
                cudaEventRecord(e0);
                if (m == 0) ldsrows<0><<<blocks, threads>>>(out, 1.0);
                if (m == 1) ldsrows<1><<<blocks, threads>>>(out, 1.0);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                cudaEventElapsedTime(&ms, e0, e1);
            }
            const double ops = (double) blocks * threads * ITER * 4;
            printf("LDS.128 + 2 DADD, random rows, %-16s warps/SM %2d: %.2f SM-cycles per warp-level fetch\n", m == 0 ? "compact (5 rows)" : "lane-replicated", warps,
                   ms * 1e-3 * 1.965e9 * 148 / (ops / 32));
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
