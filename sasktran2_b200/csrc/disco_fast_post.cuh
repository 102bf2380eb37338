// Particular (Green's function) solution and line-of-sight source multipliers of one layer, N lanes per
// (wavelength, order, layer) problem: lane j owns solution j, i.e. column j of W+ and W- in registers.
//
// What is computed is disco_core.h's layer_solve (Green's part) + los_layer_terms + the ground terms of
// layer_problem_body (reference: sktran_do_rte.cpp:903-1332, sktran_do_opticallayer.cpp:94-555, 785-938,
// sktran_do_layerarray.cpp:5-288).  How it is arranged for the FP64 pipe:
//   * every Legendre sum is taken in "projected" form: with u_l = sum_q w_q P_l^m(mu_q) W+_qj and v_l likewise for
//     W-, the Green's numerators A+-_j and the LOS sums Y+-_j(los) = sum_l P_l^m(mu_los) Z+-_l are contractions of
//     a per-block shared-memory table (uniform LDS.128 broadcasts, >= 4 DFMA each) with lane-private registers -
//     no shuffles, no per-LOS N x N products;
//   * everything indexed by j (h+-, D+-, C+-, the rows of wvec) is lane-local; the only exchange is the pair of
//     coefficient vectors A-_j C-_j, A+_j C+_j that the row phase (lane = stream i) needs for G+- (2N doubles
//     through shared memory);
//   * exponentials that do not depend on the order m (LOS attenuation, the beam factor E) come from k_los_atten.
#pragma once
#include "disco_kernels.cuh"

namespace disco {

// branch-free division for denominators inside the float range (1 + mu k, secant + k, ...)
__device__ __forceinline__ double div_fast(double num, double den) {
    float seed;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(seed) : "f"((float)den));
    double y = (double)seed;
    double e = fma(-den, y, 1.0);
    y = fma(y, e, y);
    e = fma(-den, y, 1.0);
    y = fma(y, e, y);
    const double t = num * y;
    return fma(fma(-den, t, num), y, t);
}

// psi(a; k1, k2) = (e1 - e2) / (a (k2 - k1)), e1 = exp(-a k1), e2 = exp(-a k2) (disco_core.h), without
// transcendental calls: direct difference when x = a |k2 - k1| > 0.1, Taylor series of phi(x) otherwise.
__device__ __forceinline__ double psi_fast(double a, double k1, double k2, double e1, double e2) {
    const double dx = a * (k2 - k1);
    const double x = fabs(dx);
    if (x > 0.1) return div_fast(e1 - e2, dx);
    // phi(x) = sum_n (-x)^n / (n+1)!
    double s = 1.0 / 39916800.0;  // 1/11!
    s = fma(s, -x, 1.0 / 3628800.0);
    s = fma(s, -x, 1.0 / 362880.0);
    s = fma(s, -x, 1.0 / 40320.0);
    s = fma(s, -x, 1.0 / 5040.0);
    s = fma(s, -x, 1.0 / 720.0);
    s = fma(s, -x, 1.0 / 120.0);
    s = fma(s, -x, 1.0 / 24.0);
    s = fma(s, -x, 1.0 / 6.0);
    s = fma(s, -x, 0.5);
    s = fma(s, -x, 1.0);
    return (dx >= 0.0 ? e1 : e2) * s;
}

// K2c-pre: thread per (wavelength, LOS, layer boundary): order-independent exponentials of the LOS integration
//   los_att [nw][nlos][L+1]  exp(-cum_od(top of layer p) / mu_los), p = L: the whole column
//   los_lay [nw][nlos][L][3] exp(-od_p / mu_los) | E_p = t_p / (1 + mu s_p) (1 - e^{-od s} e^{-od/mu}) | 1 / (1 + mu s_p)
__global__ void __launch_bounds__(128) k_los_atten(ChunkView V) {
    const int L = V.T.L, nlos = V.T.nlos;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)V.nw * nlos * (L + 1)) return;
    const int p = (int)(idx % (L + 1));
    const int los = (int)((idx / (L + 1)) % nlos);
    const int w = (int)(idx / ((long long)(L + 1) * nlos));
    const double mu = V.T.los_mu[los];
    V.los_att[idx] = exp(-V.lay_cumod[(size_t)w * (L + 1) + p] / mu);
    if (p < L) {
        const double od = V.lay_od[(size_t)w * L + p], s = V.lay_secant[(size_t)w * L + p];
        const double t = V.lay_trans[(size_t)w * (L + 1) + p];
        const double att = exp(-od / mu);
        const double inv = 1.0 / (1.0 + mu * s);
        double* o = V.los_lay + (((size_t)w * nlos + los) * L + p) * 3;
        o[0] = att;
        o[1] = t * inv * (1.0 - exp(-od * s) * att);
        o[2] = inv;
    }
}

template <int N>
struct PostCfg {
    static constexpr int NSTR = 2 * N;
    static constexpr int PPW = 32 / N;           // problems per warp
    static constexpr int WARPS = 4;
    static constexpr int PPB = PPW * WARPS;      // problems per block
};

// dynamic shared memory layout (doubles): tW[NSTR][N] | tL[nlos][NSTR] | lpc[NSTR] | wmu[N] | xch[PPB][2N]
template <int N>
__host__ __device__ constexpr int post_smem_doubles(int nlos) {
    return PostCfg<N>::NSTR * N + nlos * PostCfg<N>::NSTR + PostCfg<N>::NSTR + N + PostCfg<N>::PPB * 2 * N;
}

// THERMAL: the instantiation with the thermal-emission terms (its extra live values cost the third resident block, so
// the solar-only instantiation stays as it was)
template <int N, bool THERMAL>
__global__ void __launch_bounds__(128) k_layer_post(ChunkView V) {
    using Cf = PostCfg<N>;
    constexpr int NSTR = Cf::NSTR;
    extern __shared__ __align__(16) double smem[];
    const int L = V.T.L, M = V.M, nlos = V.T.nlos;
    double* tW = smem;                      // [l][q] w_q P_l^m(mu_q)
    double* tL = tW + NSTR * N;             // [los][l] P_l^m(mu_los)
    double* lpc = tL + nlos * NSTR;         // [l] P_l^m(-mu_0)-type table of the solar direction
    double* wmu = lpc + NSTR;               // [i] w_i mu_i
    double* xch = wmu + N;                  // [problem in block][2N]  A-C- | A+C+
    const int ms = blockIdx.y;
    const int m = V.m_list[ms];
    for (int e = threadIdx.x; e < NSTR * N; e += blockDim.x) {
        const int l = e / N, q = e % N;
        tW[e] = V.T.wt[q] * V.T.lp_mu[((size_t)m * N + q) * NSTR + l];
    }
    for (int e = threadIdx.x; e < nlos * NSTR; e += blockDim.x) {
        const int los = e / NSTR, l = e % NSTR;
        tL[e] = V.T.lp_los[((size_t)los * NSTR + m) * NSTR + l];
    }
    if (threadIdx.x < NSTR) lpc[threadIdx.x] = V.T.lp_csz[(size_t)m * NSTR + threadIdx.x];
    if (threadIdx.x < N) wmu[threadIdx.x] = V.T.wt[threadIdx.x] * V.T.mu[threadIdx.x];
    __syncthreads();

    const int j = threadIdx.x % N;                  // solution index (column phase) / stream index (row phase)
    const int pib = threadIdx.x / N;                // problem in block
    const unsigned lane = threadIdx.x & 31;
    const unsigned gmask = (N == 32) ? 0xffffffffu : (((1u << N) - 1u) << (lane / N * N));
    // persistent blocks: the order's tables are set up once, the block strides over its (wavelength, layer) problems
    const long long nq = (long long)V.nw * L;
    for (long long qblk = blockIdx.x; qblk * Cf::PPB < nq; qblk += gridDim.x) {
    long long q = qblk * Cf::PPB + pib;  // w * L + p
    const bool valid = q < nq;
    if (!valid) q = nq - 1;
    const int w = (int)(q / L), p = (int)(q % L);
    const size_t idx = ((size_t)w * M + ms) * L + p;
    const double od = V.lay_od[q], ssa = V.lay_ssa[q], secant = V.lay_secant[q];
    const double trans_top = V.lay_trans[(size_t)w * (L + 1) + p];
    const double* __restrict__ beta = V.lay_beta + (size_t)q * NSTR;

    // ---- column phase: lane j owns W+-[:, j]
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
    // projections on the Legendre basis; Z+-_l (l >= m, stored from index 0) and the Green's numerators
    double Zp[NSTR], Zm[NSTR];
    double ap = 0.0, am = 0.0;
    const int nl = NSTR - m;
#pragma unroll
    for (int lo = 0; lo < NSTR; ++lo) {
        Zp[lo] = 0.0;
        Zm[lo] = 0.0;
        if (lo < nl) {  // uniform over the block
            const int l = m + lo;
            const double* __restrict__ t = tW + l * N;
            double u = 0.0, v = 0.0;
#pragma unroll
            for (int qq = 0; qq < N; ++qq) {
                u = fma(t[qq], wp[qq], u);
                v = fma(t[qq], wm[qq], v);
            }
            const double ob = ssa * beta[l];
            const double su = (lo & 1) ? -u : u, sv = (lo & 1) ? -v : v;
            const double c = ob * lpc[l];
            ap = fma(c, u + sv, ap);
            am = fma(c, su + v, am);
            Zp[lo] = 0.5 * ob * (su + v);
            Zm[lo] = 0.5 * ob * (sv + u);
        }
    }
    double norm = 0.0, spj = 0.0, smj = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        norm = fma(wmu[i], fma(wp[i], wp[i], -wm[i] * wm[i]), norm);
        spj = fma(wmu[i], wp[i], spj);
        smj = fma(wmu[i], wm[i], smj);
    }
    const double f0 = (m == 0 ? 1.0 : 2.0) * (1.0 / (4.0 * kPi));
    {
        const double rn = div_fast(f0, norm);
        ap *= rn;
        am *= rn;
    }
    const double exp_sec = exp(-od * secant);
    const double psi_ks = psi_fast(od, kj, secant, thj, exp_sec);  // also feeds D-
    double Cp, Cm;
    if (fabs(secant - kj) > kGreensEps)
        Cp = trans_top * od * psi_ks;
    else
        Cp = trans_top * thj * od * (1.0 - od / 2.0 * (secant - kj));
    if (fabs(secant + kj) > kGreensEps)
        Cm = trans_top * div_fast(1.0 - exp_sec * thj, secant + kj);
    else
        Cm = trans_top * od * (1.0 - od / 2.0 * (secant + kj));
    double amc = am * Cm, apc = ap * Cp;
    // thermal source S(x) = b0 exp(-b1 x) of the layer, order 0 only (solveParticularGreenThermal,
    // sktran_do_rte.cpp:1335-1617): isotropic, so A+ = A- = (1 - ssa) sum_i w_i (W+_ij + W-_ij) / norm_j; its
    // A C products join the solar ones in G+-
    const bool thermal = THERMAL && (m == 0);   // uniform over the block
    double ath = 0.0, b0 = 0.0, b1 = 0.0, e_b1 = 0.0;
    if (thermal) {
        b0 = V.lay_thermal[(size_t)q * 2];
        b1 = V.lay_thermal[(size_t)q * 2 + 1];
        double a = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) a = fma(V.T.wt[i], wp[i] + wm[i], a);
        ath = (1.0 - ssa) * a / norm;
        e_b1 = exp(-od * b1);
        double Cpt, Cmt;
        if (fabs(b1 - kj) > kGreensEps)
            Cpt = b0 * (thj - e_b1) / (b1 - kj);
        else
            Cpt = b0 * thj * od * (1.0 - od / 2.0 * (b1 - kj));
        if (fabs(b1 + kj) > kGreensEps)
            Cmt = b0 * (1.0 - e_b1 * thj) / (b1 + kj);
        else
            Cmt = b0 * od * (1.0 - od / 2.0 * (b1 + kj));
        amc = fma(ath, Cmt, amc);
        apc = fma(ath, Cpt, apc);
    }
    xch[pib * 2 * N + j] = amc;
    xch[pib * 2 * N + N + j] = apc;
    __syncwarp();

    // ---- row phase: lane i = j owns stream i: G+-top/bottom_i = sum_j coefficient_j W-+_ij
    double gpb_i;
    {
        const int i = j;
        const double2* __restrict__ Wp2 = reinterpret_cast<const double2*>(V.Wp + idx * N * N + i * N);
        const double2* __restrict__ Wm2 = reinterpret_cast<const double2*>(V.Wm + idx * N * N + i * N);
        const double* xa = xch + pib * 2 * N;
        double gpt = 0.0, gmt = 0.0, gpb = 0.0, gmb = 0.0;
        if (N >= 2) {
#pragma unroll
            for (int c = 0; c < N / 2; ++c) {
                const double2 a = Wp2[c], b = Wm2[c];
                const double m0 = xa[2 * c], m1 = xa[2 * c + 1], p0 = xa[N + 2 * c], p1 = xa[N + 2 * c + 1];
                gpt = fma(m0, b.x, gpt); gpt = fma(m1, b.y, gpt);
                gmt = fma(m0, a.x, gmt); gmt = fma(m1, a.y, gmt);
                gpb = fma(p0, a.x, gpb); gpb = fma(p1, a.y, gpb);
                gmb = fma(p0, b.x, gmb); gmb = fma(p1, b.y, gmb);
            }
        } else {
            const double a = V.Wp[idx], b = V.Wm[idx];
            gpt = xa[0] * b; gmt = xa[0] * a; gpb = xa[1] * a; gmb = xa[1] * b;
        }
        if (valid) {
            double* __restrict__ G = V.G + idx * 4 * N;
            G[i] = gpt;
            G[N + i] = gmt;
            G[2 * N + i] = gpb;
            G[3 * N + i] = gmb;
        }
        gpb_i = gpb;
    }
    // Lambertian surface couples only m = 0 (sktran_do_surface.h:53-60): sums over the streams of the bottom layer
    const bool ground = (p == L - 1) && (m == 0);
    double sG = 0.0;
    if (ground) {  // rare (one layer, one order): a butterfly over the problem's lanes is fine here
        sG = wmu[j] * gpb_i;
#pragma unroll
        for (int off = N / 2; off > 0; off >>= 1) sG += __shfl_xor_sync(gmask, sG, off);
        if (valid) {
            double* surf = V.surf + (size_t)w * (2 * N + 1);
            surf[j] = spj;
            surf[N + j] = smj;
            if (j == 0) surf[2 * N] = sG;
        }
    }

    // ---- LOS loop (column phase again)
    const double inv_spk = div_fast(1.0, secant + kj);
    const double c_ss0 = (2 * j < nl) ? ssa * beta[m + 2 * j] * lpc[m + 2 * j] : 0.0;           // (-1)^(l-m) = +1
    const double c_ss1 = (2 * j + 1 < nl) ? -ssa * beta[m + 2 * j + 1] * lpc[m + 2 * j + 1] : 0.0;
    const double albedo = V.albedo[w];
    const double trans_floor = V.lay_trans[(size_t)w * (L + 1) + L];
    // per-LOS scalars are fetched one iteration ahead (the loop is short on independent work to hide an L2 round trip)
    double n_mu = 1.0, n_att = 0.0, n_E = 0.0, n_inv = 0.0, n_atop = 0.0;
    if (nlos > 0) {
        const double* ll0 = V.los_lay + (((size_t)w * nlos + 0) * L + p) * 3;
        n_mu = V.T.los_mu[0];
        n_att = ll0[0];
        n_E = ll0[1];
        n_inv = ll0[2];
        n_atop = V.los_att[((size_t)w * nlos + 0) * (L + 1) + p];
    }
    for (int los = 0; los < nlos; ++los) {
        const double mu = n_mu, att = n_att, E = n_E, inv_1mus = n_inv, att_top = n_atop;
        {
            const int ln = los + 1 < nlos ? los + 1 : los;
            const double* lln = V.los_lay + (((size_t)w * nlos + ln) * L + p) * 3;
            n_mu = V.T.los_mu[ln];
            n_att = lln[0];
            n_E = lln[1];
            n_inv = lln[2];
            n_atop = V.los_att[((size_t)w * nlos + ln) * (L + 1) + p];
        }
        if (V.T.los_zero && V.T.los_zero[m * nlos + los]) {   // order m gives this LOS nothing (uniform over the block)
            if (valid) {
                const size_t o = (((size_t)w * M + ms) * nlos + los) * L + p;
                V.wvec[o * 2 * N + j] = 0.0;
                V.wvec[o * 2 * N + N + j] = 0.0;
                V.vsrc[o * N + j] = 0.0;
            }
            continue;
        }
        const double imu = div_fast(1.0, mu);
        const double* __restrict__ tl = tL + los * NSTR + m;
        double Yp = 0.0, Ym = 0.0, Yp1 = 0.0, Ym1 = 0.0;  // two DFMA chains per sum (latency, not throughput, binds)
#pragma unroll
        for (int c = 0; c < NSTR / 4; ++c) {
            if (4 * c < nl) {  // uniform; Z is zero-padded up to the next multiple of 4 (tl may run past nl: the
                               // table row is followed by valid shared memory and multiplied by zero)
#pragma unroll
                for (int r = 0; r < 4; r += 2) {
                    Yp = fma(tl[4 * c + r], Zp[4 * c + r], Yp);
                    Ym = fma(tl[4 * c + r], Zm[4 * c + r], Ym);
                    Yp1 = fma(tl[4 * c + r + 1], Zp[4 * c + r + 1], Yp1);
                    Ym1 = fma(tl[4 * c + r + 1], Zm[4 * c + r + 1], Ym1);
                }
            }
        }
        Yp += Yp1;
        Ym += Ym1;
        const double hp = div_fast(1.0 - thj * att, 1.0 + mu * kj);
        double hm;
        {
            const double den = 1.0 - mu * kj;
            if (fabs(den) > 0.0001)
                hm = od * imu * psi_fast(od, kj, imu, thj, att);
            else
                hm = thj * od * imu * (1.0 - od * (kj - imu));
        }
        const double Dp = (E - trans_top * exp_sec * hm) * inv_spk;
        const double Dm = trans_top * (mu * hp - od * att * psi_ks) * inv_1mus;
        double cpos = Yp * hp * att_top, cneg = Ym * hm * att_top;
        // this lane's share of the particular + single-scatter term (summed over j by k_radiance)
        double v = ap * Yp * Dm + am * Ym * Dp;
        if (V.include_ss) v = fma(f0 * E, fma(c_ss0, tl[2 * j], c_ss1 * tl[2 * j + 1]), v);
        if (thermal) {
            // thermal part of V and the unscattered emission E_thermal (1 - ssa) (sktran_do_opticallayer.cpp:421-478,
            // 524-531, 941-957; x = 0); the latter once per problem
            const double E_th = b0 / (1.0 + mu * b1) * (1.0 - e_b1 * att);
            const double Dp_th = (E_th - b0 * e_b1 * hm) / (b1 + kj);
            const double Dm_th = (b0 * hp - E_th) / (b1 - kj);
            v = fma(ath, fma(Yp, Dm_th, Ym * Dp_th), v);
            if (j == 0) v = fma(E_th, 1.0 - ssa, v);
        }
        v *= att_top;
        if (ground) {
            // ground-leaving radiance toward the LOS (sktran_do_layerarray.cpp:5-288), attenuated by the column
            const double attg = V.los_att[((size_t)w * nlos + los) * (L + 1) + L] * albedo;
            cpos = fma(attg * 2.0 * spj, thj, cpos);
            cneg = fma(attg * 2.0, smj, cneg);
            if (j == 0) v += attg * ((V.include_ss ? V.T.csz / kPi * trans_floor : 0.0) + 2.0 * sG);
            // surface emission: unreflected, inside the reference's direct-bounce branch (sktran_do_layerarray.cpp:225-266)
            if (j == 0 && V.semis && V.include_ss) v += V.los_att[((size_t)w * nlos + los) * (L + 1) + L] * V.semis[w];
        }
        if (valid) {
            const size_t o = (((size_t)w * M + ms) * nlos + los) * L + p;
            V.wvec[o * 2 * N + j] = cpos;
            V.wvec[o * 2 * N + N + j] = cneg;
            V.vsrc[o * N + j] = v;
        }
    }
    __syncwarp();  // the exchange slots are rewritten by the next problem
    }
}

}  // namespace disco
