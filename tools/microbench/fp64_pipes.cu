// Micro-benchmarks that size the FP64 design choices on B200 (sm_100a):
//   DFMA peak, DMMA (mma.sync.m8n8k4.f64) peak, both interleaved, SHFL and LDS.128-broadcast throughput.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_pipes fp64_pipes.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int MODE>
__global__ void __launch_bounds__(256) k(double* out, int iters) {
    double a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = 1.0 + 1e-9 * (threadIdx.x + i);
    const double b = 1.0000001, c = 1e-12;
    double c0[8], c1[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c0[i] = 0.0; c1[i] = 0.0; }
    __shared__ double sm[1024];
    if (threadIdx.x < 1024 / 4) for (int i = 0; i < 4; ++i) sm[threadIdx.x * 4 + i] = threadIdx.x;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0 || MODE == 2) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = fma(a[i], b, c);
        }
        if (MODE == 1 || MODE == 2) {
#pragma unroll
            for (int i = 0; i < 8; ++i) dmma(c0[i], c1[i], a[i & 1] , b);
        }
        if (MODE == 3) {  // 64-bit shuffles
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = __shfl_xor_sync(0xffffffffu, a[i], 1 + (i & 3));
        }
        if (MODE == 4) {  // LDS.128 broadcast: 4 groups of 8 lanes, each group one 16-byte address
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const double2 v = *reinterpret_cast<const double2*>(&sm[((lane >> 3) * 66 + i * 2 + (it & 31) * 2) & 1022]);
                a[2 * i] += v.x;
                a[2 * i + 1] += v.y;
            }
        }
        if (MODE == 5) {  // DFMA fed by LDS.128 broadcast, 4 FMA per LDS.128
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const double2 v = *reinterpret_cast<const double2*>(&sm[((lane >> 3) * 66 + i * 2 + (it & 31) * 2) & 1022]);
                a[4 * i] = fma(a[4 * i], v.x, c);
                a[4 * i + 1] = fma(a[4 * i + 1], v.y, c);
                a[4 * i + 2] = fma(a[4 * i + 2], v.x, c);
                a[4 * i + 3] = fma(a[4 * i + 3], v.y, c);
            }
        }
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c0[i] + c1[i];
    if (s == 12345.678) out[0] = s;
}

template <int MODE>
double run(int blocks, int threads, int iters) {
    double* d; cudaMalloc(&d, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, threads>>>(d, 100);
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0); k<MODE><<<blocks, threads>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    cudaFree(d);
    return best * 1e-3;
}

int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("SMs %d clock %d kHz\n", sms, clk);
    const int iters = 20000;
    for (int wpb : {4, 8}) {
        const int threads = wpb * 32, blocks = sms * (wpb == 4 ? 4 : 2) * 2;
        const double nthreads = (double)blocks * threads;
        double t0 = run<0>(blocks, threads, iters);
        printf("[%d thr/blk] DFMA: %.2f TFLOP/s\n", threads, 2.0 * 16 * iters * nthreads / t0 / 1e12);
        double t1 = run<1>(blocks, threads, iters);
        printf("[%d thr/blk] DMMA m8n8k4: %.2f TFLOP/s (%.3f warp-instr/clk/SM)\n", threads, 2.0 * 256 * 8 * iters * (nthreads / 32) / t1 / 1e12,
               8.0 * iters * (nthreads / 32) / t1 / sms / (clk * 1e3));
        double t2 = run<2>(blocks, threads, iters);
        printf("[%d thr/blk] DFMA+DMMA interleaved: %.2f TFLOP/s total (dfma-only time %.3f ms, dmma-only %.3f ms, both %.3f ms)\n", threads,
               (2.0 * 16 * iters * nthreads + 2.0 * 256 * 8 * iters * (nthreads / 32)) / t2 / 1e12, t0 * 1e3, t1 * 1e3, t2 * 1e3);
        double t3 = run<3>(blocks, threads, iters);
        printf("[%d thr/blk] SHFL.64: %.3f 64-bit warp-shuffles/clk/SM\n", threads, 16.0 * iters * (nthreads / 32) / t3 / sms / (clk * 1e3));
        double t4 = run<4>(blocks, threads, iters);
        printf("[%d thr/blk] LDS.128 bcast: %.3f warp-instr/clk/SM\n", threads, 8.0 * iters * (nthreads / 32) / t4 / sms / (clk * 1e3));
        double t5 = run<5>(blocks, threads, iters);
        printf("[%d thr/blk] DFMA fed by LDS.128 (4:1): %.2f TFLOP/s\n", threads, 2.0 * 16 * iters * nthreads / t5 / 1e12);
    }
    return 0;
}
