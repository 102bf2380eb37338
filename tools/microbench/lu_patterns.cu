// Micro-benchmark of the two broadcast patterns of the staircase LU on B200 (sm_100a):
//   A  row per lane: the pivot lane stores its 34-double row (17 STS.128 from ONE lane), every lane reads it back
//      (17 uniform LDS.128) and does 33 DFMA
//   B  2D distribution, 8 row-lanes x 4 column groups, 3 row slots per lane: four lanes store 8-double segments
//      (4 STS.128 warp-wide), 8 lanes store 3 multipliers, every lane reads its segment (4 LDS.128) + 3 multipliers
//      and does 3 x 9 DFMA
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lu_patterns lu_patterns.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(double* out, int iters, int warps_per_block) {
    extern __shared__ __align__(16) double sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* buf = sm + warp * 64;
    double a[36];
#pragma unroll
    for (int i = 0; i < 36; ++i) a[i] = 1.0 + 1e-9 * (threadIdx.x + i);
    int piv = 0;
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
            if (lane == piv) {
#pragma unroll
                for (int i = 0; i < 34; i += 2) *reinterpret_cast<double2*>(buf + i) = make_double2(a[i], a[i + 1]);
            }
            __syncwarp();
            const double f = a[0] * buf[33];
#pragma unroll
            for (int i = 1; i < 34; i += 2) {
                const double2 v = *reinterpret_cast<const double2*>(buf + i - 1);
                a[i] = fma(-f, v.y, a[i]);
                if (i + 1 < 34) a[i + 1] = fma(-f, v.x, a[i + 1]);
            }
            __syncwarp();
        } else {
            const int grp = lane >> 3, ri = lane & 7;
            if (ri == (piv & 7)) {  // four lanes, one per column group
#pragma unroll
                for (int i = 0; i < 8; i += 2) *reinterpret_cast<double2*>(buf + grp * 8 + i) = make_double2(a[i], a[i + 1]);
            }
            if (grp == (piv >> 3 & 3)) {  // eight lanes: multipliers of their three rows
                *reinterpret_cast<double2*>(buf + 32 + ri * 4) = make_double2(a[8] * 1.0001, a[9] * 1.0001);
                buf[32 + ri * 4 + 2] = a[10] * 1.0001;
            }
            __syncwarp();
            const double2 m01 = *reinterpret_cast<const double2*>(buf + 32 + ri * 4);
            const double m2 = buf[32 + ri * 4 + 2];
#pragma unroll
            for (int i = 0; i < 8; i += 2) {
                const double2 v = *reinterpret_cast<const double2*>(buf + grp * 8 + i);
                a[i] = fma(-m01.x, v.x, a[i]);
                a[i + 1] = fma(-m01.x, v.y, a[i + 1]);
                a[12 + i] = fma(-m01.y, v.x, a[12 + i]);
                a[13 + i] = fma(-m01.y, v.y, a[13 + i]);
                a[24 + i] = fma(-m2, v.x, a[24 + i]);
                a[25 + i] = fma(-m2, v.y, a[25 + i]);
            }
            a[8] = fma(-m01.x, a[9], a[8]);
            a[20] = fma(-m01.y, a[9], a[20]);
            a[32] = fma(-m2, a[9], a[32]);
            __syncwarp();
        }
        piv = (piv * 5 + 3) & 31;
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 36; ++i) s += a[i];
    if (s == 12345.678) out[0] = s;
}

template <int MODE>
void run(const char* name, int blocks_per_sm, int wpb) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double* d; cudaMalloc(&d, 8);
    const int iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const size_t smem = wpb * 64 * 8;
    k<MODE><<<sms * blocks_per_sm, wpb * 32, smem>>>(d, 100, wpb);
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0); k<MODE><<<sms * blocks_per_sm, wpb * 32, smem>>>(d, iters, wpb); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    const double cyc = best * 1e-3 * clk * 1e3;
    printf("%s  %2d warps/SM: %.1f cycles per pivot step per warp, %.1f per step per SM-round (all warps)\n", name, blocks_per_sm * wpb,
           cyc / iters, cyc / iters);
    cudaFree(d);
}

int main() {
    for (int bps : {2, 4, 8}) {
        run<0>("A row-per-lane ", bps, 4);
        run<1>("B 2D 8x4       ", bps, 4);
    }
    return 0;
}
