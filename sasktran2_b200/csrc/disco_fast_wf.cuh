// Layer-local weighting-function kernel, N lanes per (wavelength, order, layer) problem; lane j owns solution j:
// column j of W+-, its derivatives and every scalar indexed by j live in registers.
//
// What is computed is disco_wf_body.h's wf_layer_body (reverse mode: dI = d(source) + z^T (db - dA x); reference:
// sktran_do_rte.cpp:198-298, 903-1332, 1793-1895; sktran_do_opticallayer.cpp:94-555).  How:
//   * only the "heavy" lanes e in {eps_g, omega} touch matrices; tau, t (beam transmittance) and s (secant) enter
//     through scalar functions of (k_j, tau, s, t) whose partials are written out by hand (everything particular
//     is linear in t);
//   * eigen-derivatives by perturbation theory in projected form: with p_l[j] = sum_a w_a P_l(mu_a) X_aj the
//     coupling matrix is P_ij = -[sum_odd c_l pm_l[i] pm_l[j] + lambda_i sum_even c_l p_l[i] p_l[j]] / n_i, i.e. a
//     contraction of a shared (all-gathered) [l][i] table with lane-private coefficients; S+ dX_j = sum_i Xm_i Q_ij
//     needs no S+;
//   * the adjoint term is regrouped per solution:  adj_j = W+_j . xi+ + W-_j . xi- with
//     xi+ = L zeta1 + A zeta2 + B zeta3 + M zeta4, xi- = A zeta1 + L zeta2 + M zeta3 + B zeta4,
//     A = theta M + A- C-, B = theta L + A+ C+, zeta = the four N-slices of z on the layer's two boundaries, so the
//     derivative is d(adj_j) = dW+_j . xi+ + dW-_j . xi- + dA (W+_j.zeta2 + W-_j.zeta1) + dB (W+_j.zeta3 + W-_j.zeta4);
//   * exchanges (projection tables, X / Xm, the per-LOS phase sums, the final sum over j) go through shared memory
//     as uniform LDS.128 broadcasts; no shuffles on the hot path.
#pragma once
#include "disco_fast_post.cuh"

namespace disco {

struct R3 {  // value and partials with respect to (k_j, tau, secant)
    double v, k, a, s;
};

// phi(x) = (1 - e^-x)/x and phi'(x) for x >= 0, given r = e^-x (only used when x > 0.1)
__device__ __forceinline__ void phi_pair(double x, double r, double& ph, double& dph) {
    if (x > 0.1) {
        const double ix = div_fast(1.0, x);
        ph = (1.0 - r) * ix;
        dph = (r - ph) * ix;
    } else {
        double s = 1.0 / 39916800.0;
        s = fma(s, -x, 1.0 / 3628800.0);
        s = fma(s, -x, 1.0 / 362880.0);
        s = fma(s, -x, 1.0 / 40320.0);
        s = fma(s, -x, 1.0 / 5040.0);
        s = fma(s, -x, 1.0 / 720.0);
        s = fma(s, -x, 1.0 / 120.0);
        s = fma(s, -x, 1.0 / 24.0);
        s = fma(s, -x, 1.0 / 6.0);
        s = fma(s, -x, 0.5);
        ph = fma(s, -x, 1.0);
        // phi'(x) = sum_n (-1)^(n+1) (n+1) x^n / (n+2)!
        double d = -10.0 / 39916800.0;
        d = fma(d, -x, -9.0 / 3628800.0);
        d = fma(d, -x, -8.0 / 362880.0);
        d = fma(d, -x, -7.0 / 40320.0);
        d = fma(d, -x, -6.0 / 5040.0);
        d = fma(d, -x, -5.0 / 720.0);
        d = fma(d, -x, -4.0 / 120.0);
        d = fma(d, -x, -3.0 / 24.0);
        d = fma(d, -x, -2.0 / 6.0);
        dph = fma(d, -x, -0.5);
    }
}

// psi(a; k1, k2) = (e1 - e2)/(a (k2 - k1)) with partials w.r.t. k1 (-> .k), a (-> .a) and k2 (-> .s)
__device__ __forceinline__ R3 psi_dual(double a, double k1, double k2, double e1, double e2) {
    R3 r;
    double ph, dph;
    if (k2 >= k1) {
        const double x = a * (k2 - k1);
        // e2 / e1 = exp(-x).  div_fast seeds from a single-precision reciprocal: exponentials of optically thick layers
        // leave the float range long before they underflow in double (exp(-k tau) < 1e-38 at k tau > 87), where the
        // quotient comes from the exponential itself
        const double ratio = (x > 0.1 && e1 > 0.0) ? (e1 > 1e-30 ? div_fast(e2, e1) : exp(-x)) : 0.0;
        phi_pair(x, ratio, ph, dph);
        r.v = e1 * ph;
        r.k = -a * e1 * (ph + dph);
        r.a = e1 * fma(dph, k2 - k1, -k1 * ph);
        r.s = a * e1 * dph;
    } else {
        const double x = a * (k1 - k2);
        const double ratio = (x > 0.1 && e2 > 0.0) ? (e2 > 1e-30 ? div_fast(e1, e2) : exp(-x)) : 0.0;
        phi_pair(x, ratio, ph, dph);
        r.v = e2 * ph;
        r.k = a * e2 * dph;
        r.a = e2 * fma(dph, k1 - k2, -k2 * ph);
        r.s = -a * e2 * (ph + dph);
    }
    return r;
}

template <int N, int G>
struct WfCfg {
    static constexpr int NSTR = 2 * N;
    static constexpr int NL = G + 4, NH = G + 1;
    static constexpr int PPW = 32 / N, WARPS = 4, PPB = PPW * WARPS;
    // per-problem shared memory (doubles)
    static constexpr int EXCH = NSTR * N + 2 * N * N + 2 * N;  // projs | Xs | Xms | lam, pad
    __host__ __device__ static constexpr int lps_doubles(int nlos) { return nlos * 2 * NH * N; }
    __host__ __device__ static constexpr int red_doubles(int nlos) { return nlos * (NL + 1) * N; }
    // Skew between the private areas of a warp's 32/N problems: the stride is padded to 4 (mod 16) doubles = 8 banks.
    // Broadcast reads (every problem its own address, LDS.64 / LDS.128) then land on disjoint banks, and the "lane j
    // writes element j" stores of the 32/N problems (N consecutive doubles each) overlap at most two deep, the minimum
    // for 256 bytes.  (The former 2-double skew left the stores four deep: 38 % of this kernel's shared wavefronts
    // were bank conflicts in profiles/ncu_r02_v1_summary.csv.)
    __host__ __device__ static constexpr int per_problem_raw(int nlos) {
        return lps_doubles(nlos) + (red_doubles(nlos) > EXCH ? red_doubles(nlos) : EXCH);
    }
    __host__ __device__ static constexpr int per_problem(int nlos) {
        return per_problem_raw(nlos) + ((4 - per_problem_raw(nlos) % 16) + 16) % 16;
    }
    // block tables: tW[NSTR][N] | tM[NSTR][N] | tL[nlos][NSTR] | lpc[NSTR] | wmu[N]
    __host__ __device__ static constexpr int table_doubles(int nlos) { return 2 * NSTR * N + nlos * NSTR + NSTR + N; }
    __host__ __device__ static constexpr int smem_doubles(int nlos) { return table_doubles(nlos) + PPB * per_problem(nlos); }
};

template <int N, int G, bool TILED>
__global__ void __launch_bounds__(128) k_wf_layer_fast(ChunkView V, int tlos_arg) {
    // TILED: lines of sight in tiles of tlos.  The per-LOS private areas (phase sums, staged adjoint slices, reduction
    // rows) are sized for one tile and reused tile after tile, so the shared-memory footprint does not grow with nlos.
    // The untiled instantiation (all LOS fit, tlos = nlos) has no tile loop: its registers are all spoken for.
    using Cf = WfCfg<N, G>;
    constexpr int NSTR = Cf::NSTR, NL = Cf::NL, NH = Cf::NH;
    constexpr int iTau = G, iOm = G + 1, iT = G + 2, iS = G + 3;
    extern __shared__ __align__(16) double smem[];
    const int L = V.T.L, M = V.M, nlos = V.T.nlos;
    const int tlos = TILED ? tlos_arg : nlos;
    double* tW = smem;                   // [l][q]  w_q P_l^m(mu_q)
    double* tM = tW + NSTR * N;          // [l][a]  P_l^m(mu_a) / mu_a
    double* tL = tM + NSTR * N;          // [los][l]
    double* lpc = tL + nlos * NSTR;      // [l]
    double* wmu = lpc + NSTR;            // [i] w_i mu_i
    const int ms = blockIdx.y;
    const int m = V.m_list[ms];
    {
        // The order's Legendre tables, precomputed in this layout at engine creation (Tables::wf_tab), come in as ONE
        // bulk copy of the TMA engine (cp.async.bulk global -> shared, completion on an mbarrier): no LSU
        // instructions, and the block's threads set up their problem while it is in flight.  The size is a multiple of
        // 16 bytes for every instantiated N (N even); source and destination are 16-byte aligned.
        __shared__ __align__(8) unsigned long long tab_bar;
        const unsigned td_bytes = (unsigned)(Cf::table_doubles(nlos) * sizeof(double));
        const unsigned bar = (unsigned)__cvta_generic_to_shared(&tab_bar);
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(td_bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             (unsigned)__cvta_generic_to_shared(smem)),
                         "l"(V.T.wf_tab + (size_t)m * Cf::table_doubles(nlos)), "r"(td_bytes), "r"(bar)
                         : "memory");
        }
        __syncthreads();   // the barrier is initialised before anyone polls it
        asm volatile(
            "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t"
            "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(bar)
            : "memory");
    }

    const int j = threadIdx.x % N;
    const int pib = threadIdx.x / N;
    const unsigned lane = threadIdx.x & 31;
    const unsigned gmask = (N == 32) ? 0xffffffffu : (((1u << N) - 1u) << (lane / N * N));
    double* pp = smem + Cf::table_doubles(nlos) + (size_t)pib * Cf::per_problem(tlos);
    double* lpsS = pp;                            // [los in tile][2 NH][q]: minus, plus, then (d minus, d plus) per group
    double* red = pp + Cf::lps_doubles(tlos);     // [los in tile][NL + 1][j]   (aliases the eigen exchange area)
    double* projs = red;                          // [lo][i]
    double* Xs = projs + NSTR * N;                // [a][i]
    double* Xms = Xs + N * N;                     // [a][i]
    double* lam = Xms + N * N;                    // [i]

    // one pass per block: a persistent variant (block strides over its problems, tables set up once) was measured
    // 20 % slower - the loop-carried state costs registers this kernel does not have
    long long q = (long long)blockIdx.x * Cf::PPB + pib;  // w * L + p
    const long long nq = (long long)V.nw * L;
    bool valid = q < nq;
    if (!valid) q = nq - 1;
    const int w = (int)(q / L), p = (int)(q % L);
    // above a kernel-based BRDF the layer on the ground is left to k_wf_layer (every order reflects there, with the
    // reflection rows of k_surface_general): its lanes run along and store nothing
    if (V.wf_bottom_only && p == L - 1) valid = false;
    const size_t idx = ((size_t)w * M + ms) * L + p;
    const double od = V.lay_od[q], ssa = V.lay_ssa[q], secant = V.lay_secant[q];
    const double trans_top = V.lay_trans[(size_t)w * (L + 1) + p];
    const double* __restrict__ beta = V.lay_beta + (size_t)q * NSTR;
    const double* __restrict__ dbeta = V.lay_dbeta + (size_t)q * G * NSTR;  // [g][l]
    const int nl = NSTR - m;
    const double f0 = (m == 0 ? 1.0 : 2.0) * (1.0 / (4.0 * kPi));

    // ---- prologue: LOS phase sums lps-+[q] (lane q) and their eps_g derivatives for the first tile -> shared memory
    const int nt0 = TILED ? (nlos < tlos ? nlos : tlos) : nlos;
    {
        double ob[NSTR], tq[NSTR], obd[G > 0 ? G : 1][NSTR];
#pragma unroll
        for (int lo = 0; lo < NSTR; ++lo) {
            const bool in = lo < nl;
            const int l = in ? m + lo : m;
            // the stream factor is folded into the moment factors once (one product less per moment and LOS below)
            tq[lo] = in ? 0.5 * tW[l * N + j] : 0.0;
            ob[lo] = in ? ssa * beta[l] * tq[lo] : 0.0;
#pragma unroll
            for (int g = 0; g < G; ++g) obd[g][lo] = in ? ssa * dbeta[g * NSTR + l] * tq[lo] : 0.0;
        }
        for (int los = 0; los < nt0; ++los) {
            const double* __restrict__ tl = tL + los * NSTR + m;
            // even and odd (l - m) accumulated separately: two independent DFMA chains per sum, and
            // lps_minus = even + odd, lps_plus = even - odd without the sign flips
            double ae = 0.0, ao = 0.0, dae[G > 0 ? G : 1], dao[G > 0 ? G : 1];
#pragma unroll
            for (int g = 0; g < G; ++g) dae[g] = dao[g] = 0.0;
#pragma unroll
            for (int c = 0; c < NSTR / 4; ++c) {
                if (4 * c < nl) {
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const int lo = 4 * c + r;
                        const double x = tl[lo];
                        if (lo & 1) {
                            ao = fma(ob[lo], x, ao);
#pragma unroll
                            for (int g = 0; g < G; ++g) dao[g] = fma(obd[g][lo], x, dao[g]);
                        } else {
                            ae = fma(ob[lo], x, ae);
#pragma unroll
                            for (int g = 0; g < G; ++g) dae[g] = fma(obd[g][lo], x, dae[g]);
                        }
                    }
                }
            }
            double* o = lpsS + (size_t)los * 2 * NH * N;
            o[j] = ae + ao;      // lps_minus
            o[N + j] = ae - ao;  // lps_plus
#pragma unroll
            for (int g = 0; g < G; ++g) {
                o[(2 + 2 * g) * N + j] = dae[g] + dao[g];
                o[(3 + 2 * g) * N + j] = dae[g] - dao[g];
            }
        }
    }

    // ---- column j of W+-, eigenvalue, BVP solution
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
    const double kj = V.kth[idx * 2 * N + j], thj = V.kth[idx * 2 * N + N + j];
    const double Lj = V.xsol[idx * 2 * N + j], Mj = V.xsol[idx * 2 * N + N + j];
    const double lamj = kj * kj;
    double norm = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) norm = fma(wmu[i], fma(wp[i], wp[i], -wm[i] * wm[i]), norm);
    const double inv_norm = div_fast(1.0, norm);

    // ---- pass 1: projections of X = W+ + W- (even l - m) and Xm = k (W+ - W-) (odd) -> exchange
    double pr[NSTR];
    {
        const double sc_even = kj * inv_norm;  // lambda_j / n_j,  n_j = k_j norm_j
        const double sc_odd = div_fast(inv_norm, kj);   // 1 / n_j
#pragma unroll
        for (int lo = 0; lo < NSTR; ++lo) {
            pr[lo] = 0.0;
            if (lo < nl) {
                const double* __restrict__ t = tW + (m + lo) * N;
                double u = 0.0, v = 0.0;
#pragma unroll
                for (int qq = 0; qq < N; ++qq) {
                    u = fma(t[qq], wp[qq], u);
                    v = fma(t[qq], wm[qq], v);
                }
                pr[lo] = (lo & 1) ? kj * (u - v) : (u + v);
            }
            projs[lo * N + j] = pr[lo] * ((lo & 1) ? sc_odd : sc_even);
        }
#pragma unroll
        for (int a = 0; a < N; ++a) {
            Xs[a * N + j] = wp[a] + wm[a];
            Xms[a * N + j] = kj * (wp[a] - wm[a]);
        }
        lam[j] = lamj;
    }
    __syncwarp();

    // ---- pass 2: eigen-derivatives for the heavy lanes e = eps_0..eps_{G-1}, omega
    double dwp[NH][N], dwm[NH][N], dk[NH];
    const double ikj = div_fast(1.0, kj);
    {
        double P[NH][N], dxm[NH][N];
#pragma unroll
        for (int e = 0; e < NH; ++e)
#pragma unroll
            for (int i = 0; i < N; ++i) P[e][i] = dxm[e][i] = 0.0;
#pragma unroll
        for (int lo = 0; lo < NSTR; ++lo) {
            if (lo < nl) {
                const int l = m + lo;
                double cp[NH];
#pragma unroll
                for (int e = 0; e < NH; ++e) cp[e] = -pr[lo] * (e < G ? ssa * dbeta[(e < G ? e : 0) * NSTR + l] : beta[l]);
                const double* __restrict__ ps = projs + lo * N;
#pragma unroll
                for (int i = 0; i < N; ++i) {
                    const double x = ps[i];
#pragma unroll
                    for (int e = 0; e < NH; ++e) P[e][i] = fma(cp[e], x, P[e][i]);
                }
                if (!(lo & 1)) {  // dS+ X: -(1/mu_a) sum_even c_l P_l(mu_a) p_l[j]
                    const double* __restrict__ tm = tM + l * N;
#pragma unroll
                    for (int a = 0; a < N; ++a) {
                        const double x = tm[a];
#pragma unroll
                        for (int e = 0; e < NH; ++e) dxm[e][a] = fma(cp[e], x, dxm[e][a]);
                    }
                }
            }
        }
        // Q_ij = P_ij / (lambda_j - lambda_i), dk_j = P_jj / (2 k_j)
        double Q[NH][N];
#pragma unroll
        for (int e = 0; e < NH; ++e) dk[e] = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const bool self = (i == j);
            const double rden = self ? 0.0 : div_fast(1.0, lamj - lam[i]);
#pragma unroll
            for (int e = 0; e < NH; ++e) {
                if (self) dk[e] = P[e][i];   // P_jj, scaled below (no division per i)
                Q[e][i] = P[e][i] * rden;
            }
        }
#pragma unroll
        for (int e = 0; e < NH; ++e) dk[e] *= 0.5 * ikj;
#pragma unroll
        for (int a = 0; a < N; ++a) {
            double dx[NH];
#pragma unroll
            for (int e = 0; e < NH; ++e) dx[e] = 0.0;
            const double* __restrict__ xr = Xs + a * N;
            const double* __restrict__ xmr = Xms + a * N;
#pragma unroll
            for (int i = 0; i < N; ++i) {
                const double x = xr[i], xm = xmr[i];
#pragma unroll
                for (int e = 0; e < NH; ++e) {
                    dx[e] = fma(x, Q[e][i], dx[e]);
                    dxm[e][a] = fma(xm, Q[e][i], dxm[e][a]);
                }
            }
            const double xmj = wp[a] - wm[a];  // Xm_aj / k_j
#pragma unroll
            for (int e = 0; e < NH; ++e) {
                const double t1 = (dxm[e][a] - xmj * dk[e]) * ikj;
                dwp[e][a] = 0.5 * (dx[e] + t1);
                dwm[e][a] = 0.5 * (dx[e] - t1);
            }
        }
    }
    __syncwarp();  // the exchange area is reused by the LOS reduction below
    // Stage the adjoint solution slices of a tile's LOS (the 4N entries of z on the layer's two boundaries) into the
    // head of each LOS's reduction row with cp.async: the HBM latency overlaps pass 3 and the scalar set-up instead of
    // stalling every LOS iteration (2 warps per scheduler cannot hide it).  Layout per LOS: zt[2N] | zb[2N].
    // zadj is [row][los] (LOS fastest): the rows of this layer's two boundaries x the tile's LOS are runs of nt
    // contiguous elements (one run when the tile is all LOS), copied element by element (lane-contiguous 8-byte chunks)
    const unsigned red0 = (unsigned)__cvta_generic_to_shared(red);
    auto stage_z = [&](int los0, int nt) {
        const bool bottom_ = (p == L - 1);
        const int row0 = (p == 0) ? 0 : N + (p - 1) * 2 * N;
        const int nrows = ((p == 0) ? N : 2 * N) + (bottom_ ? N : 2 * N);   // <= 4 N: at most four rows per lane
        const int toa_shift = (p == 0) ? N : 0;       // the TOA boundary has N rows: zb starts at 2N
        const double* zsrc = V.zadj + (((size_t)w * M + ms) * ((size_t)2 * N * L) + row0) * nlos + los0;
        // lane j copies rows j, j + N, j + 2N, j + 3N of every LOS of the tile: the row bookkeeping is done once, the
        // copies of one LOS are four cp.async with fixed offsets (the element-by-element walk over the contiguous
        // run cost ~25 instructions per 8-byte copy)
        unsigned doff[4];
        const double* src[4];
        bool ok[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int zi = j + r * N;
            ok[r] = zi < nrows;
            doff[r] = red0 + 8u * (unsigned)((zi >= N) ? zi + toa_shift : zi);
            src[r] = zsrc + (size_t)zi * nlos;
        }
        for (int t = 0; t < nt; ++t) {
            const unsigned d0 = 8u * (unsigned)(t * (NL + 1) * N);
#pragma unroll
            for (int r = 0; r < 4; ++r)
                if (ok[r])
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(doff[r] + d0), "l"(src[r] + t) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    stage_z(0, nt0);

    // ---- pass 3: Green's coefficients A+-_j with heavy-lane derivatives
    double ap, am, dap[NH], dam[NH];
    {
        double apn = 0.0, amn = 0.0, dapn[NH], damn[NH], dnorm[NH];
#pragma unroll
        for (int e = 0; e < NH; ++e) dapn[e] = damn[e] = dnorm[e] = 0.0;
#pragma unroll
        for (int lo = 0; lo < NSTR; ++lo) {
            if (lo < nl) {
                const int l = m + lo;
                const double* __restrict__ t = tW + l * N;
                double du[NH], dv[NH];
#pragma unroll
                for (int e = 0; e < NH; ++e) du[e] = dv[e] = 0.0;
#pragma unroll
                for (int qq = 0; qq < N; ++qq) {
                    const double x = t[qq];
#pragma unroll
                    for (int e = 0; e < NH; ++e) {
                        du[e] = fma(x, dwp[e][qq], du[e]);
                        dv[e] = fma(x, dwm[e][qq], dv[e]);
                    }
                }
                const double bl = beta[l], lc = lpc[l];
                const double c = ssa * bl * lc;
                // u +- v are pass 1's projections: pr[lo] = u + v (even l - m), k_j (u - v) (odd)
                const double s1 = (lo & 1) ? pr[lo] * ikj : pr[lo];   // u + s v
                const double s2 = (lo & 1) ? -s1 : s1;                // s u + v
                apn = fma(c, s1, apn);
                amn = fma(c, s2, amn);
#pragma unroll
                for (int e = 0; e < NH; ++e) {
                    const double dc = (e < G ? ssa * dbeta[(e < G ? e : 0) * NSTR + l] : bl) * lc;
                    const double d1 = (lo & 1) ? (du[e] - dv[e]) : (du[e] + dv[e]);
                    const double d2 = (lo & 1) ? (dv[e] - du[e]) : (du[e] + dv[e]);
                    dapn[e] = fma(dc, s1, fma(c, d1, dapn[e]));
                    damn[e] = fma(dc, s2, fma(c, d2, damn[e]));
                }
            }
        }
#pragma unroll
        for (int i = 0; i < N; ++i)
#pragma unroll
            for (int e = 0; e < NH; ++e)
                dnorm[e] = fma(2.0 * wmu[i], fma(wp[i], dwp[e][i], -wm[i] * dwm[e][i]), dnorm[e]);
        ap = f0 * apn * inv_norm;
        am = f0 * amn * inv_norm;
#pragma unroll
        for (int e = 0; e < NH; ++e) {
            dap[e] = (f0 * dapn[e] - ap * dnorm[e]) * inv_norm;
            dam[e] = (f0 * damn[e] - am * dnorm[e]) * inv_norm;
        }
    }

    // ---- scalars of solution j with partials (k, tau, s); everything particular is linear in t
    const double exp_sec = exp(-od * secant);
    const R3 psk = psi_dual(od, kj, secant, thj, exp_sec);
    R3 Cp, Cm;
    if (fabs(secant - kj) > kGreensEps) {
        Cp.v = trans_top * od * psk.v;
        Cp.k = trans_top * od * psk.k;
        Cp.a = trans_top * fma(od, psk.a, psk.v);
        Cp.s = trans_top * od * psk.s;
    } else {
        const double g = 1.0 - 0.5 * od * (secant - kj);
        Cp.v = trans_top * thj * od * g;
        Cp.k = trans_top * od * thj * (0.5 * od - od * g);
        Cp.a = trans_top * thj * (g + od * (-kj * g - 0.5 * (secant - kj)));
        Cp.s = -0.5 * trans_top * od * od * thj;
    }
    {
        const double est = exp_sec * thj, ispk = div_fast(1.0, secant + kj);
        Cm.v = trans_top * (1.0 - est) * ispk;
        Cm.k = (trans_top * od * est - Cm.v) * ispk;
        Cm.a = trans_top * est;
        Cm.s = Cm.k;
    }
    const double amc = am * Cm.v, apc = ap * Cp.v;
    const double Aj = fma(thj, Mj, amc), Bj = fma(thj, Lj, apc);
    // dA, dB per output lane [eps.. | tau | omega | t | s]
    double dA[NL], dB[NL];
#pragma unroll
    for (int e = 0; e < NH; ++e) {
        const int ln = e < G ? e : iOm;
        const double dth = -od * thj * dk[e];
        dA[ln] = fma(Mj, dth, fma(dam[e], Cm.v, am * Cm.k * dk[e]));
        dB[ln] = fma(Lj, dth, fma(dap[e], Cp.v, ap * Cp.k * dk[e]));
    }
    dA[iTau] = fma(Mj, -kj * thj, am * Cm.a);
    dB[iTau] = fma(Lj, -kj * thj, ap * Cp.a);
    dA[iS] = am * Cm.s;
    dB[iS] = ap * Cp.s;
    // once per problem (beam transmittances span the double range; below it - slant optical depths above 745 - every
    // particular term is exactly zero and so is its derivative chain)
    const double inv_tt = trans_top > 1e-290 ? 1.0 / trans_top : 0.0;   // (denormal beams would overflow the reciprocal)
    dA[iT] = amc * inv_tt;
    dB[iT] = apc * inv_tt;

    const bool bottom = (p == L - 1);
    const bool refl = bottom && (m == 0);
    const double albedo = V.albedo[w];
    const double t_floor = V.lay_trans[(size_t)w * (L + 1) + L];
    const double inv_spk = div_fast(1.0, secant + kj);
    // this lane's share of the single-scatter phase sum Q (l' = 2j, 2j+1) and its heavy-lane derivatives
    double cq0[NH + 1], cq1[NH + 1];  // [value, d eps_g.., d omega]
    {
        const bool i0 = 2 * j < nl, i1 = 2 * j + 1 < nl;
        const int l0 = i0 ? m + 2 * j : m, l1 = i1 ? m + 2 * j + 1 : m;
        const double b0 = i0 ? beta[l0] * lpc[l0] : 0.0, b1 = i1 ? -beta[l1] * lpc[l1] : 0.0;
        cq0[0] = ssa * b0;
        cq1[0] = ssa * b1;
#pragma unroll
        for (int g = 0; g < G; ++g) {
            cq0[1 + g] = i0 ? ssa * dbeta[g * NSTR + l0] * lpc[l0] : 0.0;
            cq1[1 + g] = i1 ? -ssa * dbeta[g * NSTR + l1] * lpc[l1] : 0.0;
        }
        cq0[NH] = b0;
        cq1[NH] = b1;
    }
    const size_t nrow = (size_t)2 * N * L;
    double gsum_v = 0.0;
    if (refl) {  // 2 sG + 2 sum_j (s+_j theta_j L_j + s-_j M_j), rare
        const double* surf = V.surf + (size_t)w * (2 * N + 1);
        double s = fma(surf[j] * thj, Lj, surf[N + j] * Mj);
#pragma unroll
        for (int off = N / 2; off > 0; off >>= 1) s += __shfl_xor_sync(gmask, s, off);
        gsum_v = 2.0 * (surf[2 * N] + s);
    }

    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    const double inv_ssa = div_fast(1.0, ssa);
    // per-LOS scalars are fetched one iteration ahead (software pipelining of the global loads)
    double n_mu, n_att, n_E, n_inv, n_atop;
    {
        const double* ll = V.los_lay + (((size_t)w * nlos + 0) * L + p) * 3;
        n_mu = V.T.los_mu[0];
        n_att = ll[0];
        n_E = ll[1];
        n_inv = ll[2];
        n_atop = V.los_att[((size_t)w * nlos + 0) * (L + 1) + p];
    }
    int los0 = 0;
    do {
    const int nt = TILED ? (nlos - los0 < tlos ? nlos - los0 : tlos) : nlos;
    if (TILED && los0 > 0) {
        // later tiles: the previous tile's rows have been reduced; restage z and recompute the phase sums (the
        // first tile's register tables are gone: the operands come back from shared memory / L1, same summation order)
        __syncwarp();
        stage_z(los0, nt);
        for (int t = 0; t < nt; ++t) {
            const double* __restrict__ tl = tL + (los0 + t) * NSTR + m;
            double ae = 0.0, ao = 0.0, dae[G > 0 ? G : 1], dao[G > 0 ? G : 1];
#pragma unroll
            for (int g = 0; g < G; ++g) dae[g] = dao[g] = 0.0;
#pragma unroll 2
            for (int lo = 0; lo < nl; ++lo) {
                const int l = m + lo;
                const double x = tl[lo], tqv = 0.5 * tW[l * N + j];
                const double obv = ssa * beta[l] * tqv;
                if (lo & 1) {
                    ao = fma(obv, x, ao);
#pragma unroll
                    for (int g = 0; g < G; ++g) dao[g] = fma(ssa * dbeta[g * NSTR + l] * tqv, x, dao[g]);
                } else {
                    ae = fma(obv, x, ae);
#pragma unroll
                    for (int g = 0; g < G; ++g) dae[g] = fma(ssa * dbeta[g * NSTR + l] * tqv, x, dae[g]);
                }
            }
            double* o = lpsS + (size_t)t * 2 * NH * N;
            o[j] = ae + ao;
            o[N + j] = ae - ao;
#pragma unroll
            for (int g = 0; g < G; ++g) {
                o[(2 + 2 * g) * N + j] = dae[g] + dao[g];
                o[(3 + 2 * g) * N + j] = dae[g] - dao[g];
            }
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
    }
    for (int tt = 0; tt < nt; ++tt) {
        const int los = los0 + tt;
        const double mu = n_mu, att = n_att, E = n_E, inv_1mus = n_inv, att_top = n_atop;
        const double imu = div_fast(1.0, mu);
        {
            const int ln = los + 1 < nlos ? los + 1 : los;
            const double* ll = V.los_lay + (((size_t)w * nlos + ln) * L + p) * 3;
            n_mu = V.T.los_mu[ln];
            n_att = ll[0];
            n_E = ll[1];
            n_inv = ll[2];
            n_atop = V.los_att[((size_t)w * nlos + ln) * (L + 1) + p];
        }
        if (V.T.los_zero && V.T.los_zero[m * nlos + los]) {
            // nothing of order m reaches this line of sight (uniform over the block): zero partials, no multipliers
            double* r = red + (size_t)tt * (NL + 1) * N;
#pragma unroll
            for (int c = 0; c <= NL; ++c) r[c * N + j] = 0.0;
            continue;
        }
        // -- source part: Y+-_j and heavy-lane derivatives from the shared phase sums
        const double* __restrict__ ls = lpsS + (size_t)tt * 2 * NH * N;
        double Yp = 0.0, Ym = 0.0, dYp[NH], dYm[NH];
        {
            // the "b" and "a" halves of every sum are separate DFMA chains (latency, not throughput, limits this
            // kernel at 8 warps per SM)
            double Ypa = 0.0, Yma = 0.0, dYpa[NH], dYma[NH], dYpg[G > 0 ? G : 1], dYmg[G > 0 ? G : 1];
            double dYpga[G > 0 ? G : 1], dYmga[G > 0 ? G : 1];
#pragma unroll
            for (int e = 0; e < NH; ++e) dYp[e] = dYm[e] = dYpa[e] = dYma[e] = 0.0;
#pragma unroll
            for (int g = 0; g < G; ++g) dYpg[g] = dYmg[g] = dYpga[g] = dYmga[g] = 0.0;
#pragma unroll
            for (int qq = 0; qq < N; ++qq) {
                const double a = ls[qq], b = ls[N + qq];  // lps_minus, lps_plus
                Yp = fma(b, wp[qq], Yp);
                Ypa = fma(a, wm[qq], Ypa);
                Ym = fma(b, wm[qq], Ym);
                Yma = fma(a, wp[qq], Yma);
#pragma unroll
                for (int e = 0; e < NH; ++e) {
                    dYp[e] = fma(b, dwp[e][qq], dYp[e]);
                    dYpa[e] = fma(a, dwm[e][qq], dYpa[e]);
                    dYm[e] = fma(b, dwm[e][qq], dYm[e]);
                    dYma[e] = fma(a, dwp[e][qq], dYma[e]);
                }
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    const double da = ls[(2 + 2 * g) * N + qq], db = ls[(3 + 2 * g) * N + qq];
                    dYpg[g] = fma(db, wp[qq], dYpg[g]);
                    dYpga[g] = fma(da, wm[qq], dYpga[g]);
                    dYmg[g] = fma(db, wm[qq], dYmg[g]);
                    dYmga[g] = fma(da, wp[qq], dYmga[g]);
                }
            }
            Yp += Ypa;
            Ym += Yma;
#pragma unroll
            for (int e = 0; e < NH; ++e) {
                dYp[e] += dYpa[e];
                dYm[e] += dYma[e];
            }
#pragma unroll
            for (int g = 0; g < G; ++g) {
                dYp[g] += dYpg[g] + dYpga[g];
                dYm[g] += dYmg[g] + dYmga[g];
            }
        }
        {
            const double iw = inv_ssa;
            dYp[G] = fma(Yp, iw, dYp[G]);
            dYm[G] = fma(Ym, iw, dYm[G]);
        }
        // -- multipliers with partials (k, tau, s)
        const double tha = thj * att;
        R3 hp, hm, Dp, Dm;
        {
            const double iden = div_fast(1.0, 1.0 + mu * kj);
            hp.v = (1.0 - tha) * iden;
            hp.k = (od * tha - hp.v * mu) * iden;
            hp.a = (kj + imu) * tha * iden;
        }
        {
            const double den = 1.0 - mu * kj;
            if (fabs(den) > 0.0001) {
                const R3 ps = psi_dual(od, kj, imu, thj, att);
                hm.v = od * imu * ps.v;
                hm.k = od * imu * ps.k;
                hm.a = imu * fma(od, ps.a, ps.v);
            } else {
                const double g = 1.0 - od * (kj - imu);
                hm.v = thj * od * imu * g;
                hm.k = od * imu * (-od * thj * g - od * thj);
                hm.a = imu * thj * g + od * imu * (-kj * thj * g - thj * (kj - imu));
            }
        }
        const double esa = exp_sec * att;
        const double E_a = trans_top * inv_1mus * (secant + imu) * esa;
        const double E_s = fma(-mu * inv_1mus, E, trans_top * inv_1mus * od * esa);
        {
            const double tes = trans_top * exp_sec;
            const double F_k = tes * hm.k;
            const double F_a = tes * fma(-secant, hm.v, hm.a);
            const double F_s = -od * tes * hm.v;
            Dp.v = (E - tes * hm.v) * inv_spk;
            Dp.k = (-F_k - Dp.v) * inv_spk;
            Dp.a = (E_a - F_a) * inv_spk;
            Dp.s = (E_s - F_s - Dp.v) * inv_spk;
        }
        {
            const double H_v = od * att * psk.v;
            const double H_k = od * att * psk.k;
            const double H_a = att * (psk.v - od * imu * psk.v + od * psk.a);
            const double H_s = od * att * psk.s;
            const double ti = trans_top * inv_1mus;
            Dm.v = ti * (mu * hp.v - H_v);
            Dm.k = ti * (mu * hp.k - H_k);
            Dm.a = ti * (mu * hp.a - H_a);
            Dm.s = -ti * H_s - mu * inv_1mus * Dm.v;
        }
        const double* __restrict__ tl = tL + los * NSTR + m;
        const double t0 = tl[2 * j], t1 = tl[2 * j + 1];
        const double Qj = V.include_ss ? f0 * fma(cq0[0], t0, cq1[0] * t1) : 0.0;
        const double alp = fma(hp.v, Lj, ap * Dm.v), alm = fma(hm.v, Mj, am * Dp.v);
        const double srcj = fma(Yp, alp, fma(Ym, alm, Qj * E));
        const double Sk = fma(Yp, fma(Lj, hp.k, ap * Dm.k), Ym * fma(Mj, hm.k, am * Dp.k));
        double out[NL];
        out[iTau] = fma(Yp, fma(Lj, hp.a, ap * Dm.a), fma(Ym, fma(Mj, hm.a, am * Dp.a), Qj * E_a));
        out[iS] = fma(Yp * ap, Dm.s, fma(Ym * am, Dp.s, Qj * E_s));
        out[iT] = fma(Yp * ap, Dm.v, fma(Ym * am, Dp.v, Qj * E)) * inv_tt;
#pragma unroll
        for (int e = 0; e < NH; ++e) {
            const int ln = e < G ? e : iOm;
            const double dQ = V.include_ss ? f0 * fma(cq0[1 + e], t0, cq1[1 + e] * t1) : 0.0;
            out[ln] = fma(dYp[e], alp, fma(dYm[e], alm, fma(Yp * Dm.v, dap[e], fma(Ym * Dp.v, dam[e], fma(dk[e], Sk, dQ * E)))));
        }
#pragma unroll
        for (int c = 0; c < NL; ++c) out[c] *= att_top;

        // -- adjoint part: zeta slices of z on the two boundaries of the layer
        const double* zt = red + (size_t)tt * (NL + 1) * N;  // staged above
        const double* zb = zt + 2 * N;
        double zg_sum = 0.0;
        if (bottom) {
#pragma unroll
            for (int i = 0; i < N; ++i) zg_sum += zb[i];
        }
        const double attg = bottom ? V.los_att[((size_t)w * nlos + los) * (L + 1) + L] : 0.0;
        double s21 = 0.0, s34 = 0.0;          // W+.zeta2 + W-.zeta1,  W+.zeta3 + W-.zeta4
        double s21b = 0.0, s34b = 0.0;        // second halves (independent DFMA chains)
        double dadj[NH], dadjb[NH];
#pragma unroll
        for (int e = 0; e < NH; ++e) dadj[e] = dadjb[e] = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            double z1, z2, z3, z4;
            if (p == 0) {
                z1 = -zt[i];
                z2 = 0.0;
            } else {
                z2 = zt[i];
                z1 = zt[N + i];
            }
            z4 = -zb[i];
            if (!bottom)
                z3 = -zb[N + i];
            else
                z3 = refl ? 2.0 * wmu[i] * albedo * (zg_sum + attg) : 0.0;
            const double xp = fma(Lj, z1, fma(Aj, z2, fma(Bj, z3, Mj * z4)));
            const double xm = fma(Aj, z1, fma(Lj, z2, fma(Mj, z3, Bj * z4)));
            s21 = fma(wp[i], z2, s21);
            s21b = fma(wm[i], z1, s21b);
            s34 = fma(wp[i], z3, s34);
            s34b = fma(wm[i], z4, s34b);
#pragma unroll
            for (int e = 0; e < NH; ++e) {
                dadj[e] = fma(dwp[e][i], xp, dadj[e]);
                dadjb[e] = fma(dwm[e][i], xm, dadjb[e]);
            }
        }
        s21 += s21b;
        s34 += s34b;
#pragma unroll
        for (int e = 0; e < NH; ++e) dadj[e] += dadjb[e];
#pragma unroll
        for (int c = 0; c < NL; ++c) out[c] = fma(dA[c], s21, fma(dB[c], s34, out[c]));
#pragma unroll
        for (int e = 0; e < NH; ++e) out[e < G ? e : iOm] += dadj[e];
        // -- partials of this solution -> reduction buffer (over the staged z of this LOS: everyone is done reading)
        __syncwarp();
        double* r = red + (size_t)tt * (NL + 1) * N;
#pragma unroll
        for (int c = 0; c < NL; ++c) r[c * N + j] = out[c];
        r[NL * N + j] = srcj * att_top;
        if (refl && j == 0 && valid) {
            const double direct = V.include_ss ? V.T.csz / kPi * t_floor : 0.0;
            double* gnd = V.wf_gnd + ((size_t)w * nlos + los) * 3;
            gnd[0] = attg * (direct + gsum_v) + zg_sum * (V.T.csz * t_floor / kPi + gsum_v);
            gnd[1] = (V.include_ss ? attg * albedo * V.T.csz / kPi : 0.0) + zg_sum * V.T.csz * albedo / kPi;
            gnd[2] = attg * albedo * (direct + gsum_v);
        }
    }
    __syncwarp();
    // ---- sum over the solutions j and store: rows (los, c) are dealt round-robin to the problem's lanes
    if (valid) {
        const int nrows = nt * (NL + 1);
        for (int row = j; row < nrows; row += N) {
            const double* r = red + (size_t)row * N;
            double s;
            if (N >= 4) {  // pairwise tree: log2(N) dependent additions instead of N
                double t[N];
#pragma unroll
                for (int i = 0; i < N; ++i) t[i] = r[i];
#pragma unroll
                for (int w2 = N / 2; w2 >= 1; w2 /= 2)
#pragma unroll
                    for (int i = 0; i < w2; ++i) t[i] += t[i + w2];
                s = t[0];
            } else {
                s = 0.0;
#pragma unroll
                for (int i = 0; i < N; ++i) s += r[i];
            }
            const int los = los0 + row / (NL + 1), c = row % (NL + 1);
            const size_t o = (((size_t)w * M + ms) * nlos + los) * L + p;
            if (c < NL)
                V.wf_loc[o * NL + c] = s;
            else
                V.wf_src[o] = s;
        }
    }
    los0 += tlos;
    } while (TILED && los0 < nlos);
}

}  // namespace disco
