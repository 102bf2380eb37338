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
#include <cstdlib>

#include "disco_kernels.cuh"
#include "disco_wf_body.h"

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
// One warp per wavelength (beam_body in disco_bodies.h is the serial statement used by the host emulation): lane 0
// runs the thickness scan in the reference's summation order, then the lanes share the O(L^2) chapman products - the
// one-thread-per-wavelength version kept 3 % of the warps of 10 blocks busy for 2 % of a step.
__global__ void __launch_bounds__(128) k_beam(ChunkView V, int scan_od) {
    const int w = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (w >= V.nw) return;
    const int L = V.T.L;
    double* od = V.lay_od + (size_t)w * L;
    double* cum = V.lay_cumod + (size_t)w * (L + 1);
    double* sec = V.lay_secant + (size_t)w * L;
    double* tr = V.lay_trans + (size_t)w * (L + 1);
    if (lane == 0 && scan_od) {   // once per chunk: a second SZA reuses the thicknesses of the first
        double ceiling = 0.0, floor_d = 0.0;
        cum[0] = 0.0;
        for (int p = 0; p < L; ++p) {
            floor_d += od[p];
            od[p] = floor_d - ceiling;  // M_OPTICAL_THICKNESS (sktran_do_opticallayer.cpp:21)
            ceiling = floor_d;
            cum[p + 1] = floor_d;
        }
    }
    __syncwarp();
    const double f0 = V.solar[w];
    if (lane == 0) tr[0] = f0;
    // slant optical depth to the floor of layer p; the one to its ceiling (p - 1) is recomputed by the same lane so
    // that no cross-lane exchange is needed
    for (int p = lane; p < L; p += 32) {
        const double* ch = V.chapman + (size_t)p * L;
        double slant = 0.0;
        for (int q = 0; q <= p; ++q) slant += ch[q] * od[q];
        double prev = 0.0;
        if (p > 0) {
            const double* chp = V.chapman + (size_t)(p - 1) * L;
            for (int q = 0; q < p; ++q) prev += chp[q] * od[q];
        }
        sec[p] = (slant - prev) / od[p];
        tr[p + 1] = exp(-slant) * f0;
    }
}

// Input validation pre-pass (Sasktran2::validate_input_atmosphere, cpp/lib/engine/engine.cpp:481-540;
// cpp/include/sasktran2/validation/validation.h:12-64): extinction finite and >= 0, single-scatter albedo finite and
// in [0, 1].  One thread per (grid point, wavelength) of the chunk; offenders raise status bits 8 .. 128.
__global__ void k_validate_inputs(ChunkView V) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)V.nw * V.T.nloc) return;
    const double k = V.ext[idx], a = V.ssa[idx];
    unsigned bits = 0u;
    if (!isfinite(k)) bits |= 8u;
    else if (k < 0.0) bits |= 16u;
    if (!isfinite(a)) bits |= 32u;
    else if (a < 0.0) bits |= 64u;
    else if (a > 1.0) bits |= 128u;
    if (bits) atomicOr(V.status, bits);
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
// K4: radiance[w, los] = sum_m cos(m phi) * sum_p ( wvec . x + v )
// -------------------------------------------------------------------------------------------------
// One warp per (wavelength, LOS): the lanes stream the contiguous [L][2N] slabs of wvec and x (coalesced) and
// the partial sums are combined with a butterfly at the end.  (radiance_body in disco_bodies.h is the serial
// statement of the same sum, used by the host emulation.)
__global__ void __launch_bounds__(128) k_radiance(ChunkView V) {
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= (long long)V.nw * V.T.nlos) return;
    const int L = V.T.L, M = V.M, nlos = V.T.nlos, nstr = V.T.nstr;
    const int w = (int)(warp / nlos), los = (int)(warp % nlos);
    const int n = L * nstr;  // 2N unknowns per layer
    double total = 0.0;
    for (int ms = 0; ms < M; ++ms) {
        const size_t o = (((size_t)w * M + ms) * nlos + los) * L;
        const double* __restrict__ wv = V.wvec + o * nstr;
        const double* __restrict__ vs = V.vsrc + o * V.vsrc_w;
        const double* __restrict__ x = V.xsol + ((size_t)w * M + ms) * n;
        double a0 = 0.0, a1 = 0.0;
        int e = lane;
        for (; e + 32 < n; e += 64) {
            a0 = fma(wv[e], x[e], a0);
            a1 = fma(wv[e + 32], x[e + 32], a1);
        }
        if (e < n) a0 = fma(wv[e], x[e], a0);
        for (int p = lane; p < L * V.vsrc_w; p += 32) a1 += vs[p];
        total = fma(a0 + a1, V.T.los_cosmphi[(size_t)los * nstr + V.m_list[ms]], total);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) total += __shfl_xor_sync(FULL_MASK, total, off);
    if (lane == 0) V.radiance[warp] = total;
}

// -------------------------------------------------------------------------------------------------
// launchers
// -------------------------------------------------------------------------------------------------
// Kernel-based surface (disco_brdf.h): per (wavelength, azimuth order) the coupling rows
//   R[i][q] = (1 + delta_m0) rho_m(mu_i, mu_q) w_q mu_q = sum_k args_k(w) Rss[k][m][i][q]   (streams)
//   Rl[los][q] likewise for the lines of sight
// applied to the bottom layer's solution: SP = R W+, SM = R W-, SG = R G+bottom for the ground rows of the BVP
// (sktran_do_rte.h:116-345), and the ground-leaving radiance toward every line of sight added to the bottom layer's
// wvec / vsrc (OpticalLayerArray::computeReflectedIntensities, sktran_do_layerarray.cpp:5-288).
// One block per (wavelength, order); thread t < N: stream row t, N <= t < N + nlos: line of sight t - N.
// Snow BRDF (Kokhanovsky): Fourier coefficients of r0 exp(-alpha K0 K0 / r0) / pi for every (stream / sun / LOS) pair of
// one wavelength by the reference's azimuth quadrature (sktran_do_surface.h:49-91) over the host tables of r0 and
// K0 K0 / r0; cos(m phi) by the Chebyshev recurrence.  Thread per (wavelength, pair); output pw[w][m][pair] in the
// convention of the kernel tables ((1 + delta_m0) w_q mu_q folded into the stream-incidence pairs).
__global__ void __launch_bounds__(128) k_brdf_expand_snow(ChunkView V, BrdfView B) {
    const int npairs = B.npairs, M = V.M;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (long long)V.nw * npairs) return;
    const int pair = (int)(tid % npairs), w = (int)(tid / npairs);
    const double alpha = sqrt(4.0 * kPi * B.args[(size_t)B.nargs * w]);
    const double* __restrict__ r0 = B.snow_r0 + (size_t)pair * B.nsamples;
    const double* __restrict__ g = B.snow_g + (size_t)pair * B.nsamples;
    double acc[32];
#pragma unroll
    for (int m = 0; m < 32; ++m) acc[m] = 0.0;
    for (int s = 0; s < B.nsamples; ++s) {
        const double f = B.snow_w[s] * r0[s] * exp(-alpha * g[s]) * (1.0 / kPi);
        const double c1 = B.snow_cos[s];
        double cm1 = 1.0, cm = c1;   // cos(0 phi), cos(1 phi)
        acc[0] += f;
#pragma unroll
        for (int m = 1; m < 32; ++m) {
            if (m < V.T.nstr) acc[m] = fma(f, cm, acc[m]);
            const double nx = fma(2.0 * c1, cm, -cm1);
            cm1 = cm;
            cm = nx;
        }
    }
    const int NN = V.T.N * V.T.N;
    const bool incidence = pair < NN || (pair >= NN + V.T.N && pair < npairs - V.T.nlos);   // stream-incidence pairs
    for (int ms = 0; ms < M; ++ms) {
        const int m = V.m_list[ms];
        // compute_expansion: result * 0.5 pi (2 - delta_m0); the weights already carry both mirror images of phi
        double v = acc[m] * 0.5 * kPi * (m == 0 ? 1.0 : 2.0);
        if (incidence) v *= (m == 0 ? 2.0 : 1.0) * B.snow_scale[pair];
        B.pw_out[((size_t)w * M + ms) * npairs + pair] = v;
    }
}

__global__ void k_surface_general(ChunkView V, BrdfView B) {
    surface_general_body(V, B, blockIdx.x / V.M, blockIdx.x % V.M, threadIdx.x);
}
void launch_brdf_expand_snow(const ChunkView& V, const BrdfView& B, cudaStream_t s) {
    const long long n = (long long)V.nw * B.npairs;
    if (n > 0) k_brdf_expand_snow<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(V, B);
}
void launch_surface_general(const ChunkView& V, const BrdfView& B, cudaStream_t s) {
    const int threads = ((V.T.N + V.T.nlos + 31) / 32) * 32;
    k_surface_general<<<(unsigned)((long long)V.nw * V.M), threads, 0, s>>>(V, B);
}

void launch_layer_optics(const ChunkView& V, cudaStream_t s) {
    const long long n = (long long)V.nw * V.T.L;
    k_layer_optics<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(V);
}
void launch_beam(const ChunkView& V, cudaStream_t s, bool scan_od) {
    k_beam<<<(unsigned)(((long long)V.nw * 32 + 127) / 128), 128, 0, s>>>(V, scan_od ? 1 : 0);
}
void launch_validate_inputs(const ChunkView& V, cudaStream_t s) {
    const long long n = (long long)V.nw * V.T.nloc;
    if (n > 0) k_validate_inputs<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(V);
}

template <int N>
static void launch_layer_solve_n(const ChunkView& V, cudaStream_t s) {
    const long long n = (long long)V.nw * V.M * V.T.L;
    k_layer_solve<N><<<(unsigned)((n + 127) / 128), 128, 0, s>>>(V);
}
// per-N entry points live in disco_bvp_inst.cu (compiled once per N)
#define DISCO_DECL_BVP(N) void launch_bvp_n##N(const ChunkView&, cudaStream_t); void launch_bvp_adjoint_n##N(const ChunkView&, cudaStream_t); void launch_bvp_multi_n##N(const ChunkView&, cudaStream_t);
DISCO_DECL_BVP(1) DISCO_DECL_BVP(2) DISCO_DECL_BVP(4) DISCO_DECL_BVP(8) DISCO_DECL_BVP(16)
static int adj_rhs_for(int nlos) { return nlos <= 4 ? 4 : 10; }
#define DISCO_DECL_WF(N) void launch_wf_layer_n##N(const ChunkView&, cudaStream_t);
DISCO_DECL_WF(1) DISCO_DECL_WF(2) DISCO_DECL_WF(4) DISCO_DECL_WF(8) DISCO_DECL_WF(16)

__global__ void k_wf_chain(ChunkView V, int G) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)V.nw * V.T.nlos) return;
    wf_chain_body(V, idx, G);
}
// Warp per (wavelength, LOS) version of wf_chain_body (disco_wf_body.h states the arithmetic serially): the lanes
// own layers, so the azimuth sums stream wf_loc coalesced, the chapman contraction reads rows of the [L][L] matrix
// coalesced, and the layer -> grid scatter goes through shared-memory atomics (at most two layers per grid point).
// dynamic shared memory per warp: gT[L+1] | dtau[L+1] | tail[L+1] | native[nloc (2+G) + 1]
__global__ void __launch_bounds__(128) k_wf_chain_warp(ChunkView V, int G) {
    extern __shared__ double sm_chain[];
    const int L = V.T.L, M = V.M, nlos = V.T.nlos, nstr = V.T.nstr, nloc = V.T.nloc;
    const int NL = G + 4, iTau = G, iOm = G + 1, iT = G + 2, iS = G + 3;
    const int nnative = nloc * (2 + G) + 1;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    double* gT = sm_chain + (size_t)wib * (3 * (L + 1) + nnative);
    double* dtau = gT + (L + 1);
    double* tail = dtau + (L + 1);
    double* native = tail + (L + 1);
    const long long idx = (long long)blockIdx.x * (blockDim.x >> 5) + wib;
    if (idx >= (long long)V.nw * nlos) return;
    const int w = (int)(idx / nlos), los = (int)(idx % nlos);
    const double mul = V.T.los_mu[los];
    const double* od = V.lay_od + (size_t)w * L;
    const double* sec = V.lay_secant + (size_t)w * L;
    const double* tr = V.lay_trans + (size_t)w * (L + 1);
    const double* gnd = V.wf_gnd + (size_t)idx * 3;
    for (int i = lane; i < nnative; i += 32) native[i] = 0.0;
    for (int p = lane; p <= L; p += 32) gT[p] = 0.0;
    __syncwarp();
    // azimuth sums of the local lanes; omega and scattering lanes go straight to the native derivatives
    for (int p = lane; p < L; p += 32) {
        double acc[8];  // NL <= 6
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[c] = 0.0;
        double src = 0.0;
        for (int ms = 0; ms < M; ++ms) {
            const double cf = V.T.los_cosmphi[(size_t)los * nstr + V.m_list[ms]];
            const size_t o = (((size_t)w * M + ms) * nlos + los) * L + p;
            const double* lc = V.wf_loc + o * NL;
#pragma unroll
            for (int c = 0; c < 8; ++c)
                if (c < NL) acc[c] = fma(cf, lc[c], acc[c]);
            src = fma(cf, V.wf_src[o], src);
        }
        const double loc_tau = acc[iTau], loc_om = acc[iOm], g_t = acc[iT], g_s = acc[iS];
        dtau[p] = loc_tau - g_s * sec[p] / od[p];
        tail[p] = src;
        // gT[p] and gT[p+1] receive from layers p-1, p and p, p+1: accumulate with shared-memory atomics
        atomicAdd(&gT[p], -tr[p] * g_t - g_s / od[p]);
        atomicAdd(&gT[p + 1], g_s / od[p]);
        const double totext = V.lay_totext[(size_t)w * L + p], scatext = V.lay_scatext[(size_t)w * L + p];
        const double ssal = V.lay_ssa[(size_t)w * L + p];
        for (int c = 0; c < 2; ++c) {
            const int q = V.interp_idx[p * 2 + c];
            if (q < 0) continue;
            const double wq = V.interp_w[p * 2 + c];
            const double kq = V.ext[(size_t)nloc * w + q], omq = V.ssa[(size_t)nloc * w + q];
            atomicAdd(&native[nloc + q], wq * loc_om * (kq / totext));
            atomicAdd(&native[q], wq * loc_om * ((omq - ssal) / totext));
            for (int g = 0; g < G; ++g) atomicAdd(&native[2 * nloc + g * nloc + q], wq * acc[g] * (omq * kq / scatext));
        }
    }
    __syncwarp();
    if (lane == 0) {
        gT[L] += -tr[L] * gnd[1];
        // LOS attenuation: dI/dtau_q -= (1/mu) (sum_{p>q} src_p + ground)
        double below = gnd[2];
        for (int q = L - 1; q >= 0; --q) {
            dtau[q] -= below / mul;
            below += tail[q];
        }
    }
    __syncwarp();
    // slant optical depths: T_{p+1} = sum_{q<=p} chapman[p][q] tau_q  ->  dtau_q += sum_{p>=q} gT[p+1] chapman[p][q]
    for (int q = lane; q < L; q += 32) {
        double s = 0.0;
        for (int pp = q; pp < L; ++pp) s = fma(gT[pp + 1], V.chapman[(size_t)pp * L + q], s);
        dtau[q] += s;
    }
    __syncwarp();
    for (int p = lane; p < L; p += 32) {
        const double dh = V.layer_dh[p];
        for (int c = 0; c < 2; ++c) {
            const int q = V.interp_idx[p * 2 + c];
            if (q < 0) continue;
            atomicAdd(&native[q], V.interp_w[p * 2 + c] * dh * dtau[p]);
        }
    }
    if (lane == 0) native[nloc * (2 + G)] = gnd[0];
    __syncwarp();
    double* out = V.wf_native + (size_t)idx * nnative;
    for (int i = lane; i < nnative; i += 32) out[i] = native[i];
}
__global__ void k_wf_map(ChunkView V, MappingView Mp, int w0, int nw_total, int G) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)V.nw * V.T.nlos * Mp.nout) return;
    wf_map_body(V, Mp, w0, nw_total, idx, G);
}
// surface (albedo) weighting function and log-radiance scaling
__global__ void k_wf_surface(ChunkView V, const double* d_brdf, double* out, int w0, int G) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)V.nw * V.T.nlos) return;
    const int nnative = V.T.nloc * (2 + G) + 1;
    const int w = (int)(idx / V.T.nlos);
    out[(size_t)w0 * V.T.nlos + idx] = V.wf_native[(size_t)idx * nnative + nnative - 1] * d_brdf[w0 + w];
}
// weighting functions w.r.t. the weights of a linear kernel BRDF: out = sum_k d_brdf[k][w] dI/d(weight k)
__global__ void k_wf_surface_args(ChunkView V, const double* d_brdf, size_t arg_stride, double* out, int w0) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)V.nw * V.T.nlos) return;
    const int w = (int)(idx / V.T.nlos);
    double acc = 0.0;
    for (int k = 0; k < V.brdf_nk; ++k) acc += V.wf_gndk[(size_t)idx * V.brdf_nk + k] * d_brdf[(size_t)k * arg_stride + w0 + w];
    out[(size_t)w0 * V.T.nlos + idx] = acc;
}
__global__ void k_wf_log_scale(ChunkView V, double* out, int nout, int w0, int nw_total) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)V.nw * V.T.nlos * nout) return;
    const int o = (int)(idx % nout);
    const long long wl = idx / nout;
    const int los = (int)(wl % V.T.nlos), w = (int)(wl / V.T.nlos);
    out[((size_t)o * nw_total + (w0 + w)) * V.T.nlos + los] /= V.radiance[(size_t)w * V.T.nlos + los];
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
#define DISCO_DISPATCH_FN(prefix, V, s)            \
    switch (V.T.N) {                               \
        case 1: prefix##1(V, s); break;            \
        case 2: prefix##2(V, s); break;            \
        case 4: prefix##4(V, s); break;            \
        case 8: prefix##8(V, s); break;            \
        case 16: prefix##16(V, s); break;          \
        default: break;                            \
    }
void launch_bvp(const ChunkView& V, cudaStream_t s) { DISCO_DISPATCH_FN(launch_bvp_n, V, s) }
void launch_bvp_adjoint(const ChunkView& V, cudaStream_t s) { DISCO_DISPATCH_FN(launch_bvp_adjoint_n, V, s) }
bool bvp_multi_supported(int N, int nsza) { return 3 * N <= 32 && nsza >= 2 && nsza <= 4; }
void launch_bvp_multi(const ChunkView& V, cudaStream_t s) { DISCO_DISPATCH_FN(launch_bvp_multi_n, V, s) }
void launch_wf_layer(const ChunkView& V, cudaStream_t s) { DISCO_DISPATCH_FN(launch_wf_layer_n, V, s) }
void launch_wf_chain(const ChunkView& V, cudaStream_t s) {
    const long long n = (long long)V.nw * V.T.nlos;
    const size_t per_warp = (size_t)(3 * (V.T.L + 1) + V.T.nloc * (2 + V.ngroups) + 1) * sizeof(double);
    if (V.ngroups + 4 <= 8 && 4 * per_warp <= 200 * 1024) {
        static DeviceOnce attr_set;
        if (attr_set.first()) {
            cudaFuncSetAttribute(k_wf_chain_warp, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        }
        k_wf_chain_warp<<<(unsigned)((n + 3) / 4), 128, 4 * per_warp, s>>>(V, V.ngroups);
    } else {
        k_wf_chain<<<(unsigned)((n + 63) / 64), 64, 0, s>>>(V, V.ngroups);
    }
}
void launch_wf_map(const ChunkView& V, const MappingView& Mp, int w0, int nw_total, bool log_space, cudaStream_t s) {
    const long long n = (long long)V.nw * V.T.nlos * Mp.nout;
    k_wf_map<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(V, Mp, w0, nw_total, V.ngroups);
    if (log_space) k_wf_log_scale<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(V, Mp.out, Mp.nout, w0, nw_total);
}
__global__ void k_wf_ground_reduce(ChunkView V) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)V.nw * V.T.nlos) return;
    wf_ground_reduce_body(V, idx);
}
void launch_wf_ground_reduce(const ChunkView& V, cudaStream_t s) {
    const long long n = (long long)V.nw * V.T.nlos;
    k_wf_ground_reduce<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(V);
}
void launch_wf_surface_args(const ChunkView& V, const double* d_brdf, size_t arg_stride, double* out, int w0, cudaStream_t s) {
    const long long n = (long long)V.nw * V.T.nlos;
    k_wf_surface_args<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(V, d_brdf, arg_stride, out, w0);
}
void launch_wf_surface(const ChunkView& V, const double* d_brdf, double* out, int w0, cudaStream_t s) {
    const long long n = (long long)V.nw * V.T.nlos;
    k_wf_surface<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(V, d_brdf, out, w0, V.ngroups);
}
// The adjoint solve reuses the forward factors when the one-row-per-lane solver applies (3N <= 32) and at least a
// quarter of the lanes of a warp carry a line of sight; SK_B200_ADJOINT=refactor forces the second factorisation of
// A^T (k_bvp_adjoint_v2), SK_B200_ADJOINT=reuse the transposed solve whenever it applies.
bool adjoint_reuses_factors(int N, int nlos) {
    if (3 * N > 32 || nlos < 1) return false;
    static const int mode = [] {
        const char* e = std::getenv("SK_B200_ADJOINT");
        if (e && e[0] == 'r' && e[1] == 'e' && e[2] == 'f') return 1;
        if (e && e[0] == 'r' && e[1] == 'e' && e[2] == 'u') return 2;
        return 0;
    }();
    const char* v3 = std::getenv("SK_B200_BVP");
    if (mode == 1 || (v3 && v3[0] == '3')) return false;  // the 2D elimination keeps no multiplier record
    if (mode == 2) return true;
    const int glt = tsolve_lanes(nlos);
    return tsolve_groups_per_warp(N, glt) * glt >= 8;
}
size_t bvp_lfac_stride(int N, int L) { return (size_t)L * 2 * N * ((3 * N + 2) & ~1); }
int adjoint_groups_per_problem(int nlos) { const int r = adj_rhs_for(nlos); return (nlos + r - 1) / r; }
int adjoint_max_rhs(int nlos) { return adj_rhs_for(nlos); }
// doubles of factor storage per solve group: (L+1) blocks of 2N pivot rows; the one-row-per-lane solver (3N <= 32)
// pads the rows to an even length and appends 1/pivot (BvpCfg2::FS in disco_bvp.cuh)
size_t bvp_fac_stride(int N, int nrhs, int L) {
    const size_t rowlen = 4 * (size_t)N + nrhs;
    const size_t fs = (3 * N <= 32) ? (((rowlen + 1) & ~(size_t)1) + 2) : rowlen;
    return (size_t)(L + 1) * 2 * N * fs;
}
void launch_radiance(const ChunkView& V, cudaStream_t s) {
    const long long n = (long long)V.nw * V.T.nlos;  // warps
    k_radiance<<<(unsigned)((n + 3) / 4), 128, 0, s>>>(V);
}

}  // namespace disco
