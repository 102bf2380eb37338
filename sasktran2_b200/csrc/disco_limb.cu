// CUDA kernels of the spherical ("limb") line-of-sight path (BASELINE config 4).
//
// Per wavelength chunk and per SZA of the DO grid the plane-parallel kernels (layer optics, beam, eigen-solve, BVP) run
// unchanged with that SZA's chapman factors; then
//   k_limb_coef<N>   thread per (order, layer, wavelength): the diffuse field of the solved layer at its sampled altitude,
//                    projected on the Legendre basis -> 16 numbers c_l per (SZA, layer, order) from which the reference's
//                    source table entry at ANY outgoing zenith angle is a dot product (DOSourceDiffuseStorage::
//                    accumulate_sources, source_term/do_source_diffuse_storage.cpp:697-1098; postprocessing multipliers
//                    include/sktran_disco/sktran_do_postprocessing.h:20-193); ground source (:436-695)
//   k_limb_table     thread per (needed table point, wavelength): the table values for every azimuth order with the
//                    reference's convergence rule (:1032-1060) applied in order
//   k_limb_phase     thread per (ray, grid point, wavelength): phase function at the ray's single-scattering angle
//                    (lib/phasefunction/phasehandler.cpp:380-412, max order of include/sasktran2/atmosphere/grid_storage.h:233-246)
//   k_limb_integrate thread per (ray, wavelength): SourceIntegrator::integrate_ray (lib/sourceintegrator/sourceintegrator.cpp:519-575)
//                    over the interpolated DO source (do_source_interpolated_pp.cpp:95-210) and the exact single-scatter
//                    source (lib/solar/singlescattersource.cpp:573-640, 949-1167)
// Layouts: the source table and the ground source are wavelength-FASTEST ([...][nw]) so that the 32 lanes of a warp (32
// consecutive wavelengths of one ray / table point) read and write consecutive addresses; the Legendre projections
// are [..][w][l] (one 128-byte run per problem), the phase function [ray][w][grid point] like the atmosphere inputs;
// the geometry tables are uniform across a warp (broadcast loads).
#include "disco_limb.cuh"

#include <cstdlib>

namespace disco {

template <int N>
__global__ void __launch_bounds__(128) k_limb_coef(ChunkView V, LimbView Lv, int s) {
    constexpr int NSTR = 2 * N;
    const int L = V.T.L, M = V.M, nw = V.nw;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (long long)nw * L * M) return;
    const int w = (int)(tid % nw);
    const int p = (int)((tid / nw) % L);
    const int ms = (int)(tid / ((long long)nw * L));
    const int m = V.m_list[ms];
    const size_t q = (size_t)w * L + p;
    const size_t idx = ((size_t)w * M + ms) * L + p;
    const double od = V.lay_od[q], ssa = V.lay_ssa[q], sec = V.lay_secant[q];
    const double trans = V.lay_trans[(size_t)w * (L + 1) + p];
    const double* __restrict__ beta = V.lay_beta + q * NSTR;
    const double* __restrict__ Wp = V.Wp + idx * N * N;   // row-major (stream, solution)
    const double* __restrict__ Wm = V.Wm + idx * N * N;
    const double* __restrict__ kth = V.kth + idx * 2 * N;
    const double* __restrict__ xs = V.xsol + idx * 2 * N;
    const double* __restrict__ lp = V.T.lp_mu + (size_t)m * N * NSTR;
    const double* __restrict__ lpc = V.T.lp_csz + (size_t)m * NSTR;

    // Green's function coefficients A+- of the solar particular solution (sktran_do_rte.cpp:556-580, 903-1010)
    double Qp[N], Qm[N];
    const double f0 = (m == 0 ? 1.0 : 2.0) * (1.0 / (4.0 * kPi));
#pragma unroll
    for (int i = 0; i < N; ++i) {
        double sp = 0.0, sm = 0.0;
        for (int l = m; l < NSTR; ++l) {
            const double pp = beta[l] * lp[i * NSTR + l] * lpc[l];
            sp += pp;
            sm += ((l - m) & 1) ? -pp : pp;
        }
        const double f = f0 * V.T.wt[i] * ssa;
        Qp[i] = sp * f;
        Qm[i] = sm * f;
    }
    // weights of the homogeneous and particular solutions at the sampled optical depth x = fraction * od
    const double frac = Lv.layer_fraction[p];
    const double x = frac * od;
    const double exs = exp(-x * sec), ets = exp(-od * sec);
    double a[N], b[N];
#pragma unroll
    for (int j = 0; j < N; ++j) {
        double norm = 0.0, ap = 0.0, am = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const double wp = Wp[i * N + j], wm = Wm[i * N + j];
            norm += V.T.wt[i] * V.T.mu[i] * (wp * wp - wm * wm);
            ap += Qp[i] * wp + Qm[i] * wm;
            am += Qm[i] * wp + Qp[i] * wm;
        }
        ap /= norm;
        am /= norm;
        const double k = kth[j];
        const double exk = exp(-x * k);
        const double hp = exp(-1.0 * k * od * frac), hm = exp(-k * od * (1.0 - frac));
        // D- = t (e^{-xk} - e^{-xs}) / (s - k) without its removable singularity; D+ = t (e^{-xs} - e^{-od s} e^{-(od-x)k}) / (k + s)
        const double Dm = trans * x * psi_value(x, k, sec, exk, exs);
        const double Dp = trans * (exs - ets * exp(-(od - x) * k)) / (k + sec);
        a[j] = hp * xs[j] + ap * Dm;
        b[j] = hm * xs[N + j] + am * Dp;
    }
    // diffuse field at the streams: E_q (same hemisphere pairing as Y+), F_q (Y-)
    double E[N], F[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        double e = 0.0, f = 0.0;
#pragma unroll
        for (int j = 0; j < N; ++j) {
            const double wp = Wp[i * N + j], wm = Wm[i * N + j];
            e += wp * a[j] + wm * b[j];
            f += wm * a[j] + wp * b[j];
        }
        E[i] = e;
        F[i] = f;
    }
    // c_l = beta_l / 2 sum_q w_q P_l^m(mu_q) ((-1)^(l-m) E_q + F_q): the single-scatter albedo of scat_phase_f cancels
    // against the division by it at do_source_diffuse_storage.cpp:1030
    double* __restrict__ out = Lv.coef + ((((size_t)s * L + p) * M + ms) * nw + w) * NSTR;
    for (int l = 0; l < NSTR; ++l) {
        double c = 0.0;
        if (l >= m) {
            const double sgn = ((l - m) & 1) ? -1.0 : 1.0;
#pragma unroll
            for (int i = 0; i < N; ++i) c += V.T.wt[i] * lp[i * NSTR + l] * (sgn * E[i] + F[i]);
            c *= 0.5 * beta[l];
        }
        out[l] = c;
    }
    // upwelling Lambertian ground source, order 0 (accumulate_ground_sources, :436-695)
    if (p == L - 1 && m == 0) {
        const double* __restrict__ G = V.G + idx * 4 * N;
        double diffuse = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            double sc = G[2 * N + i];
#pragma unroll
            for (int j = 0; j < N; ++j) sc += xs[j] * Wp[i * N + j] * kth[N + j] + xs[N + j] * Wm[i * N + j];
            diffuse += 2.0 * V.T.mu[i] * V.T.wt[i] * sc * V.albedo[w];
        }
        Lv.ground[(size_t)s * nw + w] = diffuse;
    }
}


// S1 for N = 2, 4, 8: N lanes per (wavelength, layer) problem, lane j = solution j in the column phase (column j of W+-
// in registers, loaded like k_layer_post: the lanes of a problem read consecutive doubles) and stream j in the row
// phase.  Persistent blocks per azimuth order; exchanges through per-problem shared memory:
//   Q+-_i (lane i)                      -> all lanes        A+-_j = (Q+ . W+_j + Q- . W-_j) / norm_j
//   a_j W+_qj + b_j W-_qj (lane j)      -> lane q sums      E_q, F_q: the diffuse field at stream q
//   E_q, F_q (lane q)                   -> all lanes        c_l for the lane's two Legendre orders
template <int N>
struct LimbCoefCfg {
    static constexpr int NSTR = 2 * N, PPW = 32 / N, PPB = PPW * 4;
    static constexpr int TRS = N + 1;                                    // padded row of the transpose buffers: lane j reads row j
    static constexpr int PER_PROBLEM_RAW = 2 * N + 2 * N * TRS + 2 * N;  // xq | tr[2][N][N + 1] | ef
    static constexpr int PER_PROBLEM = PER_PROBLEM_RAW + ((4 - PER_PROBLEM_RAW % 16) + 16) % 16;   // stride = 4 (mod 16) doubles: 8-bank skew between problems
    static constexpr int TABLE = NSTR * N + NSTR + N;                    // tW[l][q] | lpc[l] | wmu[q]
    static constexpr int SMEM = TABLE + PPB * PER_PROBLEM;
};

template <int N>
__global__ void __launch_bounds__(128) k_limb_coef_lanes(ChunkView V, LimbView Lv, int s) {
    using Cf = LimbCoefCfg<N>;
    constexpr int NSTR = Cf::NSTR;
    __shared__ __align__(16) double smem[Cf::SMEM];
    const int L = V.T.L, M = V.M, nw = V.nw;
    double* tW = smem;               // [l][q] w_q P_l^m(mu_q)
    double* lpc = tW + NSTR * N;     // [l]
    double* wmu = lpc + NSTR;        // [q]
    const int ms = blockIdx.y;
    const int m = V.m_list[ms];
    for (int e = threadIdx.x; e < NSTR * N; e += blockDim.x) {
        const int l = e / N, q = e % N;
        tW[e] = V.T.wt[q] * V.T.lp_mu[((size_t)m * N + q) * NSTR + l];
    }
    if (threadIdx.x < NSTR) lpc[threadIdx.x] = V.T.lp_csz[(size_t)m * NSTR + threadIdx.x];
    if (threadIdx.x < N) wmu[threadIdx.x] = V.T.wt[threadIdx.x] * V.T.mu[threadIdx.x];
    __syncthreads();
    const int j = threadIdx.x % N;
    const int pib = threadIdx.x / N;
    const unsigned lane = threadIdx.x & 31;
    const unsigned gmask = (N == 32) ? 0xffffffffu : (((1u << N) - 1u) << (lane / N * N));
    double* xq = smem + Cf::TABLE + (size_t)pib * Cf::PER_PROBLEM;   // [2N]  Q+ | Q-
    double* tr = xq + 2 * N;                                         // [2][N][N]
    constexpr int TRS = Cf::TRS;
    double* ef = tr + 2 * N * TRS;                                   // [2N]  E | F
    const double f0 = (m == 0 ? 1.0 : 2.0) * (1.0 / (4.0 * kPi));
    const long long nq = (long long)nw * L;
    for (long long qblk = blockIdx.x; qblk * Cf::PPB < nq; qblk += gridDim.x) {
        long long q = qblk * Cf::PPB + pib;   // w * L + p
        const bool valid = q < nq;
        if (!valid) q = nq - 1;
        const int w = (int)(q / L), p = (int)(q % L);
        const size_t idx = ((size_t)w * M + ms) * L + p;
        const double od = V.lay_od[q], ssa = V.lay_ssa[q], sec = V.lay_secant[q];
        const double trans = V.lay_trans[(size_t)w * (L + 1) + p];
        const double* __restrict__ beta = V.lay_beta + (size_t)q * NSTR;
        // column j of W+-
        double wp[N], wm[N];
        {
            const double* __restrict__ Wp = V.Wp + idx * N * N + j;
            const double* __restrict__ Wm = V.Wm + idx * N * N + j;
#pragma unroll
            for (int i = 0; i < N; ++i) {
                wp[i] = Wp[i * N];
                wm[i] = Wm[i * N];
            }
        }
        const double kj = V.kth[idx * 2 * N + j];
        const double Lj = V.xsol[idx * 2 * N + j], Mj = V.xsol[idx * 2 * N + N + j];
        // row phase: Q+-_i for stream i = j (sktran_do_rte.cpp:556-580)
        {
            double sp = 0.0, sm = 0.0;
            for (int l = m; l < NSTR; ++l) {
                const double pp = beta[l] * tW[l * N + j] * lpc[l];   // w_i folded into tW
                sp += pp;
                sm += ((l - m) & 1) ? -pp : pp;
            }
            xq[j] = sp * f0 * ssa;
            xq[N + j] = sm * f0 * ssa;
        }
        __syncwarp();
        double norm = 0.0, ap = 0.0, am = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const double qp = xq[i], qm = xq[N + i];
            norm = fma(wmu[i], fma(wp[i], wp[i], -wm[i] * wm[i]), norm);
            ap = fma(qp, wp[i], fma(qm, wm[i], ap));
            am = fma(qm, wp[i], fma(qp, wm[i], am));
        }
        ap /= norm;
        am /= norm;
        const double frac = Lv.layer_fraction[p];
        const double x = frac * od;
        const double exs = exp(-x * sec), ets = exp(-od * sec), exk = exp(-x * kj);
        const double hp = exp(-1.0 * kj * od * frac), hm = exp(-kj * od * (1.0 - frac));
        const double Dm = trans * x * psi_value(x, kj, sec, exk, exs);
        const double Dp = trans * (exs - ets * exp(-(od - x) * kj)) / (kj + sec);
        const double a = hp * Lj + ap * Dm, b = hm * Mj + am * Dp;
        // this solution's share of the diffuse field at every stream -> transpose through shared memory
#pragma unroll
        for (int i = 0; i < N; ++i) {
            tr[i * TRS + j] = fma(wp[i], a, wm[i] * b);
            tr[N * TRS + i * TRS + j] = fma(wm[i], a, wp[i] * b);
        }
        __syncwarp();
        {
            double e = 0.0, f = 0.0;
#pragma unroll
            for (int c = 0; c < N; ++c) {
                e += tr[j * TRS + c];
                f += tr[N * TRS + j * TRS + c];
            }
            ef[j] = e;
            ef[N + j] = f;
        }
        __syncwarp();
        // c_l for l = 2j, 2j + 1
        if (valid) {
            double* __restrict__ out = Lv.coef + ((((size_t)s * L + p) * M + ms) * nw + w) * NSTR;
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int l = 2 * j + r;
                double c = 0.0;
                if (l >= m) {
                    const double sgn = ((l - m) & 1) ? -1.0 : 1.0;
#pragma unroll
                    for (int i = 0; i < N; ++i) c = fma(tW[l * N + i], fma(sgn, ef[i], ef[N + i]), c);
                    c *= 0.5 * beta[l];
                }
                out[l] = c;
            }
        }
        // order-0 Lambertian ground source from the surface sums of the bottom layer (written by the layer kernel)
        if (p == L - 1 && m == 0) {
            const double* surf = V.surf + (size_t)w * (2 * N + 1);
            double t = fma(Lj * V.kth[idx * 2 * N + N + j], surf[j], Mj * surf[N + j]);
#pragma unroll
            for (int off = N / 2; off > 0; off >>= 1) t += __shfl_xor_sync(gmask, t, off);
            if (valid && j == 0) Lv.ground[(size_t)s * nw + w] = 2.0 * V.albedo[w] * (surf[2 * N] + t);
        }
        __syncwarp();   // the exchange areas are rewritten by the next problem
    }
}

__global__ void __launch_bounds__(128) k_limb_table(ChunkView V, LimbView Lv) {
    const int L = V.T.L, M = V.M, nw = V.nw, nstr = V.T.nstr;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (long long)nw * Lv.npts) return;
    const int w = (int)(tid % nw);
    const int pt = (int)(tid / nw);
    const int a = Lv.pt_angle[pt], p = L - 1 - Lv.pt_alt[pt], s = Lv.pt_sza[pt];
    const double* __restrict__ coef = Lv.coef + ((size_t)s * L + p) * M * nw * nstr;   // [ms][w][l]
    double* __restrict__ tab = Lv.table + (size_t)pt * M * nw + w;
    double prev = 0.0, prev_prev = 0.0;
    bool stopped = false;
    for (int ms = 0; ms < M; ++ms) {
        const int m = V.m_list[ms];   // the limb path solves every order: m == ms
        double v = 0.0;
        if (!stopped) {
            const double* __restrict__ la = Lv.lp_ang + ((size_t)a * nstr + m) * nstr;
            const double* __restrict__ cf = coef + ((size_t)ms * nw + w) * nstr;
            for (int l = m; l < nstr; ++l) v += la[l] * cf[l];
            // convergence in azimuth order (:1032-1060): later orders of this point stay zero
            if (m >= 2 && (fabs(v / prev) < 1e-4 || prev < 1e-10) && (fabs(v / prev_prev) < 1e-4 || prev_prev < 1e-10)) stopped = true;
        }
        tab[(size_t)ms * nw] = v;
        prev_prev = prev;
        prev = v;
    }
}

__global__ void __launch_bounds__(128) k_limb_phase(ChunkView V, LimbView Lv) {
    // thread per (wavelength, grid point), grid point fastest: the Legendre moments are read once (consecutive threads
    // read consecutive 8 nleg-byte runs) and every ray's phase value is written coalesced into phase[ray][w][q]
    const int nw = V.nw, nloc = V.T.nloc, nleg = V.nleg;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (long long)nw * nloc) return;
    const double* __restrict__ leg = V.leg + (size_t)nleg * tid;
    // the reference sums up to the last non-zero stored moment, at most num_singlescatter_moments
    const int nl = nleg < Lv.nss ? nleg : Lv.nss;
    int max_order = 1;
    for (int l = 0; l < nleg; ++l)
        if (leg[l] != 0.0) max_order = l + 1;
    if (max_order > nl) max_order = nl;
    constexpr int KR = 16;   // moments kept in registers
    double lg[KR];
#pragma unroll
    for (int l = 0; l < KR; ++l) lg[l] = (l < max_order) ? leg[l] : 0.0;
    for (int r = 0; r < Lv.nrays; ++r) {
        const double* __restrict__ wig = Lv.wig_ss + (size_t)r * Lv.nss;
        double ph = 0.0;
#pragma unroll
        for (int l = 0; l < KR; ++l)
            if (l < Lv.nss) ph = fma(lg[l], wig[l], ph);
        for (int l = KR; l < max_order; ++l) ph = fma(leg[l], wig[l], ph);
        Lv.phase[(size_t)r * nw * nloc + tid] = ph;
    }
}

__global__ void __launch_bounds__(64) k_limb_integrate(ChunkView V, LimbView Lv) {
    const int nw = V.nw, nloc = V.T.nloc, M = V.M;
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = Lv.ray_order[blockIdx.y];   // longest rays first: the last wave of blocks is made of short ones
    if (w >= nw) return;
    const double* __restrict__ ext = V.ext + (size_t)nloc * w;
    const double* __restrict__ ssa = V.ssa + (size_t)nloc * w;
    const int s0 = Lv.seg_start[r], s1 = Lv.seg_start[r + 1];
    const int b0 = s0 + r;   // first solar boundary of the ray
    const double solar = V.solar[w];
    auto sun_trans = [&](int b) {
        if (Lv.sol_blocked[b]) return 0.0;
        double od = 0.0;
        for (int e = Lv.sol_start[b]; e < Lv.sol_start[b + 1]; ++e) od += Lv.sol_w[e] * ext[Lv.sol_idx[e]];
        return exp(-od) * solar;
    };
    double I = 0.0, total_od = 0.0;
    double t_far = 0.0;
    if (Lv.ss_exact && s1 > s0) t_far = sun_trans(b0);
    if (Lv.gnd_hit[r] && s1 > s0) {
        if (Lv.ms_do)
            for (int k = 0; k < 2; ++k) {
                const double gw = Lv.gnd_sza_w[r * 2 + k];
                if (gw != 0.0) I += gw * Lv.ground[(size_t)Lv.gnd_sza_idx[r * 2 + k] * nw + w];
            }
        const double mu_in = Lv.gnd_mu_in[r];
        if (Lv.ss_exact && mu_in > 0.0) I += t_far * (V.albedo[w] / kPi) * mu_in;
    }
    const double* __restrict__ phase = Lv.phase + ((size_t)r * nw + w) * nloc;
    for (int sg = s0; sg < s1; ++sg) {
        const int* __restrict__ idx = Lv.od_idx + (size_t)sg * kLimbStencil;
        const double* __restrict__ ow = Lv.od_w + (size_t)sg * kLimbStencil;
        double od = 0.0;
#pragma unroll
        for (int k = 0; k < kLimbStencil; ++k)
            if (ow[k] != 0.0) od += ow[k] * ext[idx[k]];
        total_od += od;
        const double att = exp(-od);
        I *= att;
        if (Lv.ms_do) {
            const int* __restrict__ sp = Lv.src_pt + (size_t)sg * kLimbSrcEntries;
            if (sp[0] >= 0) {
                const double* __restrict__ sw = Lv.src_w + (size_t)sg * kLimbSrcEntries;
                const double* __restrict__ sc = Lv.src_cos + (size_t)sg * V.T.nstr;
                double sv = 0.0;
#pragma unroll 2
                for (int e = 0; e < kLimbSrcEntries; ++e) {
                    if (sp[e] < 0) break;
                    const double* __restrict__ tab = Lv.table + (size_t)sp[e] * M * nw + w;
                    double acc = 0.0;
                    for (int ms = 0; ms < M; ++ms) acc += sc[V.m_list[ms]] * tab[(size_t)ms * nw];
                    sv += sw[e] * acc;
                }
                const double omega = Lv.mid_w[sg * 2] * ssa[Lv.mid_idx[sg * 2]] + Lv.mid_w[sg * 2 + 1] * ssa[Lv.mid_idx[sg * 2 + 1]];
                I += omega * (1.0 - att) * sv;
            }
        }
        if (Lv.ss_exact) {
            const double t_near = sun_trans(b0 + (sg - s0) + 1);
            const double* __restrict__ we = Lv.ent_w + (size_t)sg * kLimbStencil;
            const double* __restrict__ wx = Lv.exit_w + (size_t)sg * kLimbStencil;
            auto endpoint = [&](const double* __restrict__ wt, double strans) {
                double a = 0.0, k = 0.0, ph = 0.0, ph_single = 0.0;
                int nz = 0;
#pragma unroll
                for (int c = 0; c < kLimbStencil; ++c) {
                    if (wt[c] == 0.0) continue;
                    const int q = idx[c];
                    const double pq = phase[q];
                    a += ssa[q] * wt[c];
                    k += ext[q] * wt[c];
                    ph += pq * wt[c];
                    ph_single = pq;
                    ++nz;
                }
                // a single contributing node enters unweighted (phasehandler.cpp:683-685)
                return k * a * strans / (kPi * 4) * (nz == 1 ? ph_single : ph);
            };
            const int lower = Lv.seg_lower[sg];
            const double start = endpoint(lower == 2 ? wx : we, t_near);
            const double end = endpoint(lower == 1 ? we : wx, t_far);
            const double sf = fabs(od) < 1e-12 ? 1.0 : -expm1(-od) / od;
            I += sf * (start * Lv.seg_qfrac[sg * 2] + end * Lv.seg_qfrac[sg * 2 + 1]) * Lv.seg_len[sg];
            t_far = t_near;
        }
    }
    Lv.radiance[(size_t)w * Lv.nrays + r] = I;
    if (Lv.los_od) Lv.los_od[(size_t)w * Lv.nrays + r] = total_od;
}

template <int N>
static void launch_limb_coef_lanes(const ChunkView& V, const LimbView& Lv, int s, cudaStream_t st) {
    using Cf = LimbCoefCfg<N>;
    const long long nq = (long long)V.nw * V.T.L;
    const long long nblk = (nq + Cf::PPB - 1) / Cf::PPB;
    static const int cap = [] {
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        return sms * 16;   // persistent blocks over all orders: a few waves
    }();
    long long per_order = (cap + V.M - 1) / V.M;
    if (per_order < 1) per_order = 1;
    const dim3 grid((unsigned)(nblk < per_order ? nblk : per_order), (unsigned)V.M);
    k_limb_coef_lanes<N><<<grid, 128, 0, st>>>(V, Lv, s);
}

void launch_limb_coef(const ChunkView& V, const LimbView& Lv, int s, cudaStream_t st) {
    const long long n = (long long)V.nw * V.T.L * V.M;
    const unsigned grid = (unsigned)((n + 127) / 128);
    static const bool generic = [] {
        const char* g = std::getenv("SK_B200_GENERIC");
        return g && g[0] == '1';
    }();
    if (!generic) {
        switch (V.T.N) {
            case 2: launch_limb_coef_lanes<2>(V, Lv, s, st); return;
            case 4: launch_limb_coef_lanes<4>(V, Lv, s, st); return;
            case 8: launch_limb_coef_lanes<8>(V, Lv, s, st); return;
            default: break;
        }
    }
    switch (V.T.N) {
        case 1: k_limb_coef<1><<<grid, 128, 0, st>>>(V, Lv, s); break;
        case 2: k_limb_coef<2><<<grid, 128, 0, st>>>(V, Lv, s); break;
        case 4: k_limb_coef<4><<<grid, 128, 0, st>>>(V, Lv, s); break;
        case 8: k_limb_coef<8><<<grid, 128, 0, st>>>(V, Lv, s); break;
        case 16: k_limb_coef<16><<<grid, 128, 0, st>>>(V, Lv, s); break;
        default: break;
    }
}
void launch_limb_table(const ChunkView& V, const LimbView& Lv, cudaStream_t st) {
    const long long n = (long long)V.nw * Lv.npts;
    if (n > 0) k_limb_table<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(V, Lv);
}
void launch_limb_phase(const ChunkView& V, const LimbView& Lv, cudaStream_t st) {
    const long long n = (long long)V.nw * V.T.nloc;
    if (n > 0 && Lv.nrays > 0) k_limb_phase<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(V, Lv);
}
void launch_limb_integrate(const ChunkView& V, const LimbView& Lv, cudaStream_t st) {
    if (V.nw <= 0 || Lv.nrays <= 0) return;
    const dim3 grid((unsigned)((V.nw + 63) / 64), (unsigned)Lv.nrays);
    k_limb_integrate<<<grid, 64, 0, st>>>(V, Lv);
}

}  // namespace disco
