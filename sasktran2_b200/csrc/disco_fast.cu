// Register-resident fast path of the layer solve (N = 2, 4, 8 streams per hemisphere): instantiations + launchers.
#include <cstdlib>

#include "disco_fast_eig.cuh"
#include "disco_fast_post.cuh"
#include "disco_fast_wf.cuh"

namespace disco {

bool fast_path_supported(int N) { return N == 2 || N == 4 || N == 8; }

// blocks per azimuth order of a persistent layer kernel: 8 blocks per SM over all orders (a few waves - short
// tails, long enough to amortise the table set-up); SK_B200_WF_BLOCKS overrides
static int wf_fast_blocks_per_order(int M) {
    static const int total = [] {
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const char* e = std::getenv("SK_B200_WF_BLOCKS");
        const int per_sm = e ? std::atoi(e) : 8;
        return sms * (per_sm > 0 ? per_sm : 8);
    }();
    const int b = (total + M - 1) / M;
    return b > 0 ? b : 1;
}

template <int N>
static void launch_eig_n(const ChunkView& V, cudaStream_t s) {
    const long long nq = (long long)V.nw * V.T.L;
    const dim3 grid_t((unsigned)((nq + 127) / 128), (unsigned)V.M);
    k_eig_setup<N><<<grid_t, 128, 0, s>>>(V);
    // SK_B200_JACOBI=unrolled: the sweep as straight-line code (differential testing of the rolled rounds)
    static const bool unrolled = [] { const char* e = std::getenv("SK_B200_JACOBI"); return e && e[0] == 'u'; }();
    if (unrolled)
        k_eig_jacobi<N, false><<<grid_t, 128, 0, s>>>(V);
    else
        k_eig_jacobi<N, true><<<grid_t, 128, 0, s>>>(V);
}
template <int N>
static void launch_post_n(const ChunkView& V, cudaStream_t s) {
    const long long nq = (long long)V.nw * V.T.L;
    if (V.T.nlos > 0) {   // the spherical path solves the layers without plane-parallel lines of sight
        const long long n = (long long)V.nw * V.T.nlos * (V.T.L + 1);
        k_los_atten<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(V);
    }
    const size_t smem = (size_t)post_smem_doubles<N>(V.T.nlos) * sizeof(double);
    static DeviceOnce attr_set;
    if (attr_set.first()) {
        cudaFuncSetAttribute(k_layer_post<N, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cudaFuncSetAttribute(k_layer_post<N, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    }
    const long long nblk_p = (nq + PostCfg<N>::PPB - 1) / PostCfg<N>::PPB;
    const long long cap_p = 2LL * wf_fast_blocks_per_order((int)V.M);  // persistent blocks, three resident per SM
    const dim3 grid_p((unsigned)(nblk_p < cap_p ? nblk_p : cap_p), (unsigned)V.M);
    if (V.emission)
        k_layer_post<N, true><<<grid_p, 128, smem, s>>>(V);
    else
        k_layer_post<N, false><<<grid_p, 128, smem, s>>>(V);
}

void launch_layer_eig_fast(const ChunkView& V, cudaStream_t s) {
    switch (V.T.N) {
        case 2: launch_eig_n<2>(V, s); break;
        case 4: launch_eig_n<4>(V, s); break;
        case 8: launch_eig_n<8>(V, s); break;
        default: break;
    }
}
void launch_layer_post_fast(const ChunkView& V, cudaStream_t s) {
    switch (V.T.N) {
        case 2: launch_post_n<2>(V, s); break;
        case 4: launch_post_n<4>(V, s); break;
        case 8: launch_post_n<8>(V, s); break;
        default: break;
    }
}
// launches 4 kernels
void launch_layer_solve_fast(const ChunkView& V, cudaStream_t s) {
    launch_layer_eig_fast(V, s);
    launch_layer_post_fast(V, s);
}

// Shared memory of k_wf_layer_fast<N, G> for nlos lines of sight processed tlos at a time (bytes; WfCfg's own formulas)
static size_t wf_fast_smem_bytes(int N, int G, int nlos, int tlos) {
    const int nstr = 2 * N, NL = G + 4, NH = G + 1, ppb = (32 / N) * 4;
    const int exch = nstr * N + 2 * N * N + 2 * N, red = tlos * (NL + 1) * N;
    const int raw = tlos * 2 * NH * N + (red > exch ? red : exch);
    const int per_problem = raw + ((4 - raw % 16) + 16) % 16;  // WfCfg::per_problem
    return sizeof(double) * (size_t)(2 * nstr * N + nlos * nstr + nstr + N + ppb * per_problem);
}
// Lines of sight per tile: the largest tile with which two blocks stay resident per SM (228 KB, 1 KB reserved per
// block), then balanced over the tiles; 0 = the order tables alone do not fit (thousands of LOS: generic kernel).
int wf_layer_fast_tile(int N, int G, int nlos) {
    if (nlos < 1 || (N != 2 && N != 4 && N != 8) || G > 2) return 0;
    const size_t budget = 113 * 1024;
    int tmax = 0;
    for (int t = 1; t <= nlos && wf_fast_smem_bytes(N, G, nlos, t) <= budget; ++t) tmax = t;
    if (tmax == 0) return 0;
    // test switch: cap the tile (SK_B200_WF_TILE=3 tiles even ten lines of sight)
    static const int cap = [] { const char* e = std::getenv("SK_B200_WF_TILE"); return e ? std::atoi(e) : 0; }();
    if (cap < 0) return 0;   // SK_B200_WF_TILE=-1: the generic thread-per-problem kernel (differential tests, timing)
    if (cap > 0 && cap < tmax) tmax = cap;
    const int ntiles = (nlos + tmax - 1) / tmax;
    return (nlos + ntiles - 1) / ntiles;
}

template <int N, int G>
static void launch_wf_fast_ng(const ChunkView& V, cudaStream_t s) {
    using Cf = WfCfg<N, G>;
    const long long nq = (long long)V.nw * V.T.L;
    const int tlos = wf_layer_fast_tile(N, G, V.T.nlos);
    const size_t smem = (size_t)(Cf::table_doubles(V.T.nlos) + Cf::PPB * Cf::per_problem(tlos)) * sizeof(double);
    static DeviceOnce attr_set;
    if (attr_set.first()) {
        cudaFuncSetAttribute(k_wf_layer_fast<N, G, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        cudaFuncSetAttribute(k_wf_layer_fast<N, G, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    }
    const dim3 grid((unsigned)((nq + Cf::PPB - 1) / Cf::PPB), (unsigned)V.M);
    if (tlos >= V.T.nlos)
        k_wf_layer_fast<N, G, false><<<grid, 128, smem, s>>>(V, tlos);
    else
        k_wf_layer_fast<N, G, true><<<grid, 128, smem, s>>>(V, tlos);
}
template <int N>
static void launch_wf_fast_n(const ChunkView& V, cudaStream_t s) {
    switch (V.ngroups) {
        case 0: launch_wf_fast_ng<N, 0>(V, s); break;
        case 1: launch_wf_fast_ng<N, 1>(V, s); break;
        case 2: launch_wf_fast_ng<N, 2>(V, s); break;
        default: break;
    }
}
void launch_wf_layer_fast(const ChunkView& V, cudaStream_t s) {
    switch (V.T.N) {
        case 2: launch_wf_fast_n<2>(V, s); break;
        case 4: launch_wf_fast_n<4>(V, s); break;
        case 8: launch_wf_fast_n<8>(V, s); break;
        default: break;
    }
}

}  // namespace disco
