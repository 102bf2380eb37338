// CUDA kernels of the batched discrete-ordinates radiance solve (sm_100a, fp64).
//
//  K1  k_layer_optics   thread per (wavelength, layer): grid -> layer optical properties
//                       (OpticalLayerArray ctor, cpp/lib/sktran_disco/sktran_do_layerarray.cpp:332-477)
//  K1b k_beam           thread per wavelength: layer thickness scan, pseudo-spherical beam
//                       (OpticalLayerArray::configureTransmission, :891-979)
//  K2  k_layer_solve    thread per (wavelength, azimuth order, layer): homogeneous + particular solution and
//                       the line-of-sight source multipliers (disco_core.h)
//  K3  k_bvp            lane group per (wavelength, azimuth order): staircase LU with partial pivoting of the
//                       layer-boundary system + back substitution (replaces LAPACK dgbsv,
//                       cpp/lib/sktran_disco/sktran_do_rte.cpp:1621-1723, 1898-2294)
//  K4  k_radiance       thread per (wavelength, LOS): azimuth sum of w.x + v
//                       (source_term/do_source_planeparallel.cpp:69-158)
#include "disco_kernels.cuh"

namespace disco {

#define FULL_MASK 0xffffffffu

// -------------------------------------------------------------------------------------------------
// K1: layer optics
// -------------------------------------------------------------------------------------------------
__global__ void k_layer_optics(ChunkView V) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)V.nw * V.T.L) return;
    optics_body(V, idx);
}

// -------------------------------------------------------------------------------------------------
// K1b: thickness scan + solar beam
// -------------------------------------------------------------------------------------------------
__global__ void k_beam(ChunkView V) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= V.nw) return;
    beam_body(V, w);
}

// -------------------------------------------------------------------------------------------------
// K2: per (w, m, layer) solve, thread per problem
// -------------------------------------------------------------------------------------------------
template <int N>
__global__ void __launch_bounds__(128) k_layer_solve(ChunkView V) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)V.nw * V.M * V.T.L) return;
    layer_problem_body<N>(V, idx);
}

// -------------------------------------------------------------------------------------------------
// K3: boundary value problem per (w, m).
//
// Unknowns x = [L_0 M_0 | L_1 M_1 | ...] (2N per layer).  Rows: N TOA rows, 2N continuity rows per interface,
// N ground rows.  Eliminating the 2N unknowns of layer p only ever involves the N rows left over from the
// layers above plus the 2N rows of interface p+1, i.e. a 3N x (4N + 1) panel.  One lane owns one panel
// row in registers; the pivot row is broadcast through shared memory.  The candidate rows of every column
// are exactly the rows LAPACK's banded partial pivoting (kl = ku = 3N-1) would search, so the factorisation
// is the reference's dgbsv in a different storage scheme.
// -------------------------------------------------------------------------------------------------
template <int N>
struct BvpCfg {
    static constexpr int NC = 2 * N;
    static constexpr int ROWS = 3 * N;
    static constexpr int GL = ROWS <= 4 ? 4 : (ROWS <= 8 ? 8 : (ROWS <= 16 ? 16 : 32));
    static constexpr int R = (ROWS + GL - 1) / GL;
    static constexpr int ROWLEN = 4 * N + 1;
    static constexpr int GROUPS_PER_WARP = 32 / GL;
    static constexpr int WARPS_PER_BLOCK = (N >= 16) ? 2 : 4;
    static constexpr int GROUPS_PER_BLOCK = GROUPS_PER_WARP * WARPS_PER_BLOCK;
    static constexpr int BUF = ROWLEN + 1;  // padded
    static constexpr int SMEM_DOUBLES_PER_GROUP = 2 * BUF + NC * ROWLEN + NC;
};

template <int N>
__global__ void __launch_bounds__(BvpCfg<N>::WARPS_PER_BLOCK * 32) k_bvp(ChunkView V) {
    using C = BvpCfg<N>;
    constexpr int NC = C::NC, GL = C::GL, R = C::R, ROWLEN = C::ROWLEN;
    extern __shared__ double smem[];
    const int L = V.T.L, M = V.M;
    const int lane_w = threadIdx.x & 31;
    const int gidx_in_block = threadIdx.x / GL;
    const int lane = threadIdx.x % GL;
    const unsigned gbase = (unsigned)((lane_w / GL) * GL);
    const unsigned gmask = (GL == 32) ? FULL_MASK : (((1u << GL) - 1u) << gbase);
    long long prob = (long long)blockIdx.x * C::GROUPS_PER_BLOCK + gidx_in_block;
    const long long nprob = (long long)V.nw * M;
    const bool valid = prob < nprob;
    if (!valid) prob = nprob - 1;
    const int w = (int)(prob / M);
    const int ms = (int)(prob % M);
    const int m = V.m_list[ms];

    double* gs = smem + (size_t)gidx_in_block * C::SMEM_DOUBLES_PER_GROUP;
    double* buf = gs;                      // [2][BUF]
    double* facs = gs + 2 * C::BUF;        // [NC][ROWLEN]
    double* xs = facs + NC * ROWLEN;       // [NC]

    const size_t lay0 = ((size_t)w * M + ms) * L;  // first layer record of this (w, m)
    const double* Wp = V.Wp + lay0 * N * N;
    const double* Wm = V.Wm + lay0 * N * N;
    const double* kth = V.kth + lay0 * 2 * N;
    const double* G = V.G + lay0 * 4 * N;
    double* fac = V.fac + lay0 * NC * ROWLEN;
    double* xout = V.xsol + lay0 * NC;

    double a[R][ROWLEN];
    bool act[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        act[r] = false;
#pragma unroll
        for (int c = 0; c < ROWLEN; ++c) a[r][c] = 0.0;
    }
    // TOA rows: W+_0 L + W-_0 Theta_0 M = -G+top_0   (sktran_do_rte.cpp:1898-1942, 2131-2172)
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int sid = lane * R + r;
        if (sid < N) {
            act[r] = true;
#pragma unroll
            for (int j = 0; j < N; ++j) {
                a[r][j] = Wp[sid * N + j];
                a[r][N + j] = Wm[sid * N + j] * kth[N + j];
            }
            a[r][4 * N] = -G[sid];
        }
    }
    const unsigned lt_mask = (lane == 0) ? 0u : (((1u << lane) - 1u) << gbase);
    bool singular = false;

    for (int p = 0; p < L; ++p) {
        // ---- bring in the rows of interface p+1 (or the ground rows) into free slots
        {
            unsigned freeb[R];
            bool wasfree[R];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                wasfree[r] = !act[r];
                freeb[r] = __ballot_sync(FULL_MASK, wasfree[r]) & gmask;
            }
            const bool last = (p == L - 1);
            const int needed = last ? N : NC;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if (wasfree[r]) {
                    // rank of this free slot in slot-id order (slot id = lane * R + r)
                    int rank = 0;
#pragma unroll
                    for (int r2 = 0; r2 < R; ++r2) rank += __popc(freeb[r2] & lt_mask);
#pragma unroll
                    for (int r2 = 0; r2 < R; ++r2)
                        if (r2 < r && wasfree[r2]) rank += 1;
                    if (rank < needed) {
                        act[r] = true;
                        const double* Wpu = Wp + (size_t)p * N * N;
                        const double* Wmu = Wm + (size_t)p * N * N;
                        const double* thu = kth + (size_t)p * 2 * N + N;
                        const double* Gu = G + (size_t)p * 4 * N;
                        if (!last) {
                            // continuity between layer p (upper) and p+1 (lower), sktran_do_rte.cpp:1945-2072
                            const double* Wpl = Wpu + N * N;
                            const double* Wml = Wmu + N * N;
                            const double* thl = thu + 2 * N;
                            const double* Gl = Gu + 4 * N;
                            const bool first = rank < N;  // rows i: W- family; rows i+N: W+ family
                            const int i = first ? rank : rank - N;
                            const double* A1 = first ? Wmu : Wpu;  // multiplies L_upper (with theta)
                            const double* A2 = first ? Wpu : Wmu;  // multiplies M_upper
                            const double* B1 = first ? Wml : Wpl;  // multiplies L_lower
                            const double* B2 = first ? Wpl : Wml;  // multiplies M_lower (with theta)
#pragma unroll
                            for (int j = 0; j < N; ++j) {
                                a[r][j] = A1[i * N + j] * thu[j];
                                a[r][N + j] = A2[i * N + j];
                                a[r][2 * N + j] = -B1[i * N + j];
                                a[r][3 * N + j] = -(B2[i * N + j] * thl[j]);
                            }
                            // rhs: -G-bot_u + G-top_l (first) / -G+bot_u + G+top_l (second), :2199-2204
                            a[r][4 * N] = first ? (-Gu[3 * N + i] + Gl[N + i]) : (-Gu[2 * N + i] + Gl[i]);
                        } else {
                            // ground rows, sktran_do_rte.cpp:2075-2128, 2270-2294, sktran_do_rte.h:116-345
                            const int i = rank;
                            const bool refl = (m == 0);
                            const double alb2 = refl ? 2.0 * V.albedo[w] : 0.0;
                            const double* surf = V.surf + (size_t)w * (2 * N + 1);
#pragma unroll
                            for (int j = 0; j < N; ++j) {
                                double vm = Wmu[i * N + j], vp = Wpu[i * N + j];
                                if (refl) {
                                    vm -= alb2 * surf[j];      // - (1+d_m0) rho sum_q w mu W+_qj
                                    vp -= alb2 * surf[N + j];  // - (1+d_m0) rho sum_q w mu W-_qj
                                }
                                a[r][j] = vm * thu[j];
                                a[r][N + j] = vp;
                                a[r][2 * N + j] = 0.0;
                                a[r][3 * N + j] = 0.0;
                            }
                            double rhs = -Gu[3 * N + i];
                            if (refl) {
                                rhs += alb2 * surf[2 * N];
                                rhs += V.T.csz * V.albedo[w] / kPi * V.lay_trans[(size_t)w * (L + 1) + L];
                            }
                            a[r][4 * N] = rhs;
                        }
                    }
                }
            }
        }
        // ---- eliminate the 2N unknowns of layer p
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            double best = -1.0;
            int bsid = 0x7fffffff;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if (act[r]) {
                    const double v = fabs(a[r][c]);
                    if (v > best) {
                        best = v;
                        bsid = lane * R + r;
                    }
                }
            }
#pragma unroll
            for (int off = GL / 2; off > 0; off >>= 1) {
                const double ov = __shfl_xor_sync(FULL_MASK, best, off);
                const int oi = __shfl_xor_sync(FULL_MASK, bsid, off);
                if (ov > best || (ov == best && oi < bsid)) {
                    best = ov;
                    bsid = oi;
                }
            }
            if (!(best > 0.0)) singular = true;
            double* bc = buf + (c & 1) * C::BUF;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if (act[r] && bsid == lane * R + r) {
                    act[r] = false;
#pragma unroll
                    for (int cc = 0; cc < ROWLEN; ++cc) {
                        const double v = (cc >= c) ? a[r][cc] : 0.0;
                        if (cc >= c) bc[cc] = v;
                        facs[c * ROWLEN + cc] = v;
                    }
                }
            }
            __syncwarp();
            const double pinv = 1.0 / bc[c];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if (act[r]) {
                    const double f = a[r][c] * pinv;
#pragma unroll
                    for (int cc = c + 1; cc < ROWLEN; ++cc) a[r][cc] -= f * bc[cc];
                    a[r][c] = 0.0;
                }
            }
        }
        __syncwarp();
        // ---- flush the pivot rows of this layer (coalesced) and slide the panel window
        if (valid) {
            double* dst = fac + (size_t)p * NC * ROWLEN;
            for (int e = lane; e < NC * ROWLEN; e += GL) dst[e] = facs[e];
        }
        __syncwarp();
        if (p < L - 1) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
#pragma unroll
                for (int j = 0; j < NC; ++j) {
                    a[r][j] = a[r][NC + j];
                    a[r][NC + j] = 0.0;
                }
            }
        }
    }
    if (singular && valid) atomicOr(V.status, 4u);

    // ---- back substitution, bottom layer first; lane c owns pivot row c
    for (int p = L - 1; p >= 0; --p) {
        if (p < L - 1) {
            const double* src = fac + (size_t)p * NC * ROWLEN;
            for (int e = lane; e < NC * ROWLEN; e += GL) facs[e] = src[e];
            __syncwarp();
        }
        double acc = 0.0, myx = 0.0;
        if (lane < NC) {
            acc = facs[lane * ROWLEN + 4 * N];
            if (p < L - 1) {
#pragma unroll
                for (int j = 0; j < NC; ++j) acc -= facs[lane * ROWLEN + NC + j] * xs[j];
            }
        }
#pragma unroll
        for (int cc = NC - 1; cc >= 0; --cc) {
            double xv = 0.0;
            if (lane == cc) xv = acc / facs[cc * ROWLEN + cc];
            xv = __shfl_sync(FULL_MASK, xv, (int)gbase + cc);
            if (lane < cc) acc -= facs[lane * ROWLEN + cc] * xv;
            if (lane == cc) myx = xv;
        }
        __syncwarp();
        if (lane < NC) {
            xs[lane] = myx;
            if (valid) xout[(size_t)p * NC + lane] = myx;
        }
        __syncwarp();
    }
}

// -------------------------------------------------------------------------------------------------
// K4: radiance[w, los] = sum_m cos(m phi) * sum_p ( wvec . x + v )
// -------------------------------------------------------------------------------------------------
__global__ void k_radiance(ChunkView V) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)V.nw * V.T.nlos) return;
    radiance_body(V, idx);
}

// -------------------------------------------------------------------------------------------------
// launchers
// -------------------------------------------------------------------------------------------------
void launch_layer_optics(const ChunkView& V, cudaStream_t s) {
    const long long n = (long long)V.nw * V.T.L;
    k_layer_optics<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(V);
}
void launch_beam(const ChunkView& V, cudaStream_t s) { k_beam<<<(V.nw + 63) / 64, 64, 0, s>>>(V); }

template <int N>
static void launch_layer_solve_n(const ChunkView& V, cudaStream_t s) {
    const long long n = (long long)V.nw * V.M * V.T.L;
    k_layer_solve<N><<<(unsigned)((n + 127) / 128), 128, 0, s>>>(V);
}
template <int N>
static void launch_bvp_n(const ChunkView& V, cudaStream_t s) {
    using C = BvpCfg<N>;
    const long long nprob = (long long)V.nw * V.M;
    const size_t smem = (size_t)C::GROUPS_PER_BLOCK * C::SMEM_DOUBLES_PER_GROUP * sizeof(double);
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(k_bvp<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_set = true;
    }
    k_bvp<N><<<(unsigned)((nprob + C::GROUPS_PER_BLOCK - 1) / C::GROUPS_PER_BLOCK), C::WARPS_PER_BLOCK * 32, smem, s>>>(V);
}

#define DISCO_DISPATCH_N(fn, V, s)                 \
    switch (V.T.N) {                               \
        case 1: fn<1>(V, s); break;                \
        case 2: fn<2>(V, s); break;                \
        case 4: fn<4>(V, s); break;                \
        case 8: fn<8>(V, s); break;                \
        case 16: fn<16>(V, s); break;              \
        default: break;                            \
    }

// FP64 roofline denominator: DFMA micro-benchmark (16 independent accumulators per thread, no memory traffic)
__global__ void __launch_bounds__(256) k_dfma_peak(double* out, int iters) {
    double a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = 1.0 + 1e-9 * (threadIdx.x + i);
    const double b = 1.0000001, c = 1e-12;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = fma(a[i], b, c);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    if (s == 12345.678) out[0] = s;  // never true; keeps the chain alive
}

double measure_fp64_tflops() {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    double* d = nullptr;
    cudaMalloc(&d, sizeof(double));
    const int blocks = sms * 8, threads = 256, iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k_dfma_peak<<<blocks, threads>>>(d, 2000);  // warm-up
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        k_dfma_peak<<<blocks, threads>>>(d, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    const double flops = 2.0 * 16.0 * (double)iters * (double)blocks * threads;
    return flops / (best * 1e-3) / 1e12;
}

bool nstr_supported(int nstr) { return nstr == 2 || nstr == 4 || nstr == 8 || nstr == 16 || nstr == 32; }
void launch_layer_solve(const ChunkView& V, cudaStream_t s) { DISCO_DISPATCH_N(launch_layer_solve_n, V, s) }
void launch_bvp(const ChunkView& V, cudaStream_t s) { DISCO_DISPATCH_N(launch_bvp_n, V, s) }
void launch_radiance(const ChunkView& V, cudaStream_t s) {
    const long long n = (long long)V.nw * V.T.nlos;
    k_radiance<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(V);
}

}  // namespace disco
