// Dedicated two-stream kernel: launch wrapper around disco_twostream_body.h (algorithm and reference citations there).
// One thread per wavelength; the [L][L] chapman table of a pseudo-spherical geometry is brought into shared memory by
// the TMA engine (cp.async.bulk + mbarrier), the optical depths of the layers above live next to it as [layer][thread].
#include <cuda_runtime.h>

#include "disco_kernels.cuh"
#include "disco_twostream_body.h"

namespace disco {

namespace {
// one-dimensional bulk copy global -> shared through the TMA engine, completion on an mbarrier
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, unsigned bytes, unsigned long long* bar) {
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    const unsigned mb = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(gmem_src), "r"(bytes), "r"(mb)
                 : "memory");
}
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    const unsigned mb = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mb), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned phase) {
    const unsigned mb = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(mb),
        "r"(phase)
        : "memory");
}

}  // namespace

// NLOS lines of sight per thread (grid.y walks batches of NLOS).  Dynamic shared memory: chapman [L][L] | od [L][blockDim]
// (pseudo-spherical only) after one 16-byte mbarrier slot.
template <int NLOS>
__global__ void __launch_bounds__(128, 3) k_twostream(ChunkView V) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(smem_raw);
    double* sh_chap = reinterpret_cast<double*>(smem_raw + 16);
    const int n = V.T.L;
    double* sh_od = sh_chap + (size_t)n * n;
    if (V.plane_parallel == 0) {
        if (((size_t)n * n * sizeof(double)) % 16 == 0) {   // bulk copies move multiples of 16 bytes
            if (threadIdx.x == 0) {
                mbar_init(bar, 1);
                tma_load_1d(sh_chap, V.chapman, (unsigned)(sizeof(double) * n * n), bar);
            }
            __syncthreads();
            mbar_wait(bar, 0);
        } else {
            for (int i = threadIdx.x; i < n * n; i += blockDim.x) sh_chap[i] = V.chapman[i];
            __syncthreads();
        }
    }
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= V.nw) return;
    ts::twostream_body<NLOS>(V, w, blockIdx.y * NLOS, sh_chap, sh_od + threadIdx.x, (int)blockDim.x);
}

size_t twostream_smem_bytes(int L, bool plane_parallel, int threads) {
    if (plane_parallel) return 16;
    return 16 + sizeof(double) * ((size_t)L * L + (size_t)L * threads);
}

// Threads per block: the pseudo-spherical kernel keeps the L x L chapman table plus one optical-depth column per thread
// in shared memory; tall grids run with narrower blocks (128 -> 64 -> 32 threads), 0: does not fit (L > ~150)
static int twostream_threads(int L, bool plane_parallel) {
    for (int t : {128, 64, 32})
        if (twostream_smem_bytes(L, plane_parallel, t) <= 220 * 1024) return t;
    return 0;
}
bool twostream_supported(int L, bool plane_parallel) { return twostream_threads(L, plane_parallel) > 0; }

void launch_twostream(const ChunkView& V, cudaStream_t s) {
    const int threads = twostream_threads(V.T.L, V.plane_parallel != 0);
    const size_t smem = twostream_smem_bytes(V.T.L, V.plane_parallel != 0, threads);
    const int nlos = V.T.nlos;
    const unsigned gx = (unsigned)((V.nw + threads - 1) / threads);
    if (nlos <= 1) {
        static DeviceOnce once;
        if (once.first()) cudaFuncSetAttribute(k_twostream<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        k_twostream<1><<<dim3(gx, (unsigned)nlos), threads, smem, s>>>(V);
    } else {
        static DeviceOnce once;
        if (once.first()) cudaFuncSetAttribute(k_twostream<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        k_twostream<2><<<dim3(gx, (unsigned)((nlos + 1) / 2)), threads, smem, s>>>(V);
    }
}

}  // namespace disco
