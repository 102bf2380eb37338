// Weighting functions (linearisation) of the DO solve — per-thread bodies, host/device.
//
// The reference propagates layer-local derivative lanes forward through each layer's solution and either
// solves the BVP for every lane (dgbtrs with 3L+1 right-hand sides, sktran_do_rte.cpp:1730-1789) or, in
// "backprop" mode, solves the transposed BVP once per line of sight (RTESolver::backprop, :1793-1895).
// This path is the reverse mode: with  I = sum_p (wvec_p . x_p + v_p),  A x = b,  A^T z = wvec,
//     dI/dtheta = sum_p (dwvec_p . x_p + dv_p)  +  z^T (db - dA x)
// and every layer only needs its own W+-, k, Green's coefficients, its 4N entries of x and the 4N entries of z on
// its two boundaries.  Per layer the local inputs are
//     [eps_g (one per scattering group) | tau | omega | t (beam transmittance at the ceiling) | s (average secant)]
// whose derivatives are carried as small dual numbers; the cross-layer dependence (t_p, s_p and the LOS
// attenuation depend on the optical depths of the layers above) is chained afterwards in wf_chain_body.
// Eigen-derivatives use first-order perturbation theory with the left eigenvectors X^-1 = diag(1/n) Xm^T D^2
// (Xm = S+ X, D^2 = diag(w mu)) instead of the reference's bordered (N+1)x(N+1) LU per eigenvalue
// (linearizeHomogeneous, sktran_do_rte.cpp:198-298): same derivative, gauge X_j^T-free instead of X_j^T dX_j = 0.
#pragma once
#include "disco_bodies.h"

namespace disco {

template <int NL>
struct Dk {
    double v;
    double d[NL];
    DISCO_HD Dk() {}
    DISCO_HD Dk(double x) : v(x) {
        for (int i = 0; i < NL; ++i) d[i] = 0.0;
    }
};
template <int NL>
DISCO_HD Dk<NL> operator+(const Dk<NL>& a, const Dk<NL>& b) {
    Dk<NL> r;
    r.v = a.v + b.v;
    for (int i = 0; i < NL; ++i) r.d[i] = a.d[i] + b.d[i];
    return r;
}
template <int NL>
DISCO_HD Dk<NL> operator-(const Dk<NL>& a, const Dk<NL>& b) {
    Dk<NL> r;
    r.v = a.v - b.v;
    for (int i = 0; i < NL; ++i) r.d[i] = a.d[i] - b.d[i];
    return r;
}
template <int NL>
DISCO_HD Dk<NL> operator-(const Dk<NL>& a) {
    Dk<NL> r;
    r.v = -a.v;
    for (int i = 0; i < NL; ++i) r.d[i] = -a.d[i];
    return r;
}
template <int NL>
DISCO_HD Dk<NL> operator*(const Dk<NL>& a, const Dk<NL>& b) {
    Dk<NL> r;
    r.v = a.v * b.v;
    for (int i = 0; i < NL; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i];
    return r;
}
template <int NL>
DISCO_HD Dk<NL> operator*(const Dk<NL>& a, double b) {
    Dk<NL> r;
    r.v = a.v * b;
    for (int i = 0; i < NL; ++i) r.d[i] = a.d[i] * b;
    return r;
}
template <int NL>
DISCO_HD Dk<NL> operator*(double b, const Dk<NL>& a) {
    return a * b;
}
template <int NL>
DISCO_HD Dk<NL> operator/(const Dk<NL>& a, const Dk<NL>& b) {
    Dk<NL> r;
    const double inv = 1.0 / b.v;
    r.v = a.v * inv;
    for (int i = 0; i < NL; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) * inv;
    return r;
}
template <int NL>
DISCO_HD Dk<NL> operator+(const Dk<NL>& a, double b) {
    Dk<NL> r = a;
    r.v += b;
    return r;
}
template <int NL>
DISCO_HD Dk<NL> operator+(double b, const Dk<NL>& a) {
    Dk<NL> r = a;
    r.v += b;
    return r;
}
template <int NL>
DISCO_HD Dk<NL> operator-(const Dk<NL>& a, double b) {
    Dk<NL> r = a;
    r.v -= b;
    return r;
}
template <int NL>
DISCO_HD Dk<NL> operator-(double b, const Dk<NL>& a) {
    Dk<NL> r = -a;
    r.v += b;
    return r;
}
template <int NL>
DISCO_HD Dk<NL> dexp(const Dk<NL>& a) {
    Dk<NL> r;
    r.v = exp(a.v);
    for (int i = 0; i < NL; ++i) r.d[i] = r.v * a.d[i];
    return r;
}

template <int NL>
DISCO_HD Dk<NL> dphi(const Dk<NL>& a) {
    Dk<NL> r;
    r.v = phi_value(a.v);
    const double dp = phi_deriv(a.v);
    for (int i = 0; i < NL; ++i) r.d[i] = dp * a.d[i];
    return r;
}
// psi(a; k1, k2) of disco_core.h with dual arguments; e1 = exp(-a k1), e2 = exp(-a k2)
template <int NL>
DISCO_HD Dk<NL> dpsi(const Dk<NL>& a, const Dk<NL>& k1, const Dk<NL>& k2, const Dk<NL>& e1, const Dk<NL>& e2) {
    if (k2.v >= k1.v) return e1 * dphi(a * (k2 - k1));
    return e2 * dphi(a * (k1 - k2));
}

// Green's function particular solution with dual inputs (same formulas and branches as layer_solve)
template <int N, int NL>
DISCO_HD void particular_dual(const Tables& T, int m, const Dk<NL>& od, const Dk<NL>& ssa, const Dk<NL>* beta,
                              const Dk<NL>& secant, const Dk<NL>& trans_top, const Dk<NL>* k, const Dk<NL>* theta,
                              const Dk<NL>* Wp, const Dk<NL>* Wm, Dk<NL>* Ap, Dk<NL>* Am, Dk<NL>* Gpt, Dk<NL>* Gmt,
                              Dk<NL>* Gpb, Dk<NL>* Gmb) {
    constexpr int NSTR = 2 * N;
    using D = Dk<NL>;
    const double* lp = T.lp_mu + (size_t)m * N * NSTR;
    const double* lpc = T.lp_csz + (size_t)m * NSTR;
    D Qp[N], Qm[N];
    for (int i = 0; i < N; ++i) {
        D sp(0.0), sm(0.0);
        for (int l = m; l < NSTR; ++l) {
            const double pp = lp[i * NSTR + l] * lpc[l];
            sp = sp + beta[l] * pp;
            sm = sm + beta[l] * (((l - m) & 1) ? -pp : pp);
        }
        const double factor = (m == 0 ? 1.0 : 2.0) * (1.0 / (4.0 * kPi)) * T.wt[i];
        Qp[i] = sp * ssa * factor;
        Qm[i] = sm * ssa * factor;
    }
    for (int i = 0; i < N; ++i) Gpt[i] = Gmt[i] = Gpb[i] = Gmb[i] = D(0.0);
    const D exp_sec = dexp(-(od * secant));
    for (int j = 0; j < N; ++j) {
        D norm(0.0), ap(0.0), am(0.0);
        for (int i = 0; i < N; ++i) {
            const D& wp = Wp[i * N + j];
            const D& wm = Wm[i * N + j];
            norm = norm + (wp * wp - wm * wm) * (T.wt[i] * T.mu[i]);
            ap = ap + Qp[i] * wp + Qm[i] * wm;
            am = am + Qm[i] * wp + Qp[i] * wm;
        }
        ap = ap / norm;
        am = am / norm;
        Ap[j] = ap;
        Am[j] = am;
        const D& kj = k[j];
        const D& exp_k = theta[j];
        D Cp, Cm;
        if (fabs(secant.v - kj.v) > kGreensEps)
            Cp = trans_top * od * dpsi(od, kj, secant, exp_k, exp_sec);
        else
            Cp = trans_top * exp_k * od * (1.0 - od * 0.5 * (secant - kj));
        if (fabs(secant.v + kj.v) > kGreensEps)
            Cm = trans_top * (1.0 - exp_sec * exp_k) / (secant + kj);
        else
            Cm = trans_top * od * (1.0 - od * 0.5 * (secant + kj));
        const D amc = am * Cm, apc = ap * Cp;
        for (int i = 0; i < N; ++i) {
            Gpt[i] = Gpt[i] + amc * Wm[i * N + j];
            Gmt[i] = Gmt[i] + amc * Wp[i * N + j];
            Gpb[i] = Gpb[i] + apc * Wp[i * N + j];
            Gmb[i] = Gmb[i] + apc * Wm[i * N + j];
        }
    }
}

// Un-attenuated layer source toward one LOS,  src = sum_j cpos_j L_j + cneg_j M_j + v, with dual layer inputs
template <int N, int NL>
DISCO_HD Dk<NL> los_source_dual(const Tables& T, int m, int los, const Dk<NL>& od, const Dk<NL>& ssa,
                                const Dk<NL>* beta, const Dk<NL>& secant, const Dk<NL>& trans_top, bool include_ss,
                                const Dk<NL>* k, const Dk<NL>* theta, const Dk<NL>* Wp, const Dk<NL>* Wm,
                                const Dk<NL>* Ap, const Dk<NL>* Am, const double* Lc, const double* Mc) {
    constexpr int NSTR = 2 * N;
    using D = Dk<NL>;
    const double mu = T.los_mu[los];
    const double* lp = T.lp_mu + (size_t)m * N * NSTR;
    const double* lpl = T.lp_los + ((size_t)los * NSTR + m) * NSTR;
    const double* lpc = T.lp_csz + (size_t)m * NSTR;
    D lps_plus[N], lps_minus[N];
    for (int q = 0; q < N; ++q) {
        D a(0.0), b(0.0);
        for (int l = m; l < NSTR; ++l) {
            const double pp = lpl[l] * lp[q * NSTR + l];
            a = a + beta[l] * pp;
            b = b + beta[l] * (((l - m) & 1) ? -pp : pp);
        }
        lps_minus[q] = a * ssa * (0.5 * T.wt[q]);
        lps_plus[q] = b * ssa * (0.5 * T.wt[q]);
    }
    D Q(0.0);
    if (include_ss) {
        D acc(0.0);
        for (int l = m; l < NSTR; ++l) {
            const double pp = lpl[l] * lpc[l];
            acc = acc + beta[l] * (((l - m) & 1) ? -pp : pp);
        }
        Q = acc * ssa * ((m == 0 ? 1.0 : 2.0) * (1.0 / (4.0 * kPi)));
    }
    const D att = dexp(-(od * (1.0 / mu)));
    const D expfactor = dexp(-(od * secant));
    const D E = trans_top / (1.0 + secant * mu) * (1.0 - expfactor * att);
    D src(0.0);
    for (int j = 0; j < N; ++j) {
        D Yp(0.0), Ym(0.0);
        for (int q = 0; q < N; ++q) {
            const D& wp = Wp[q * N + j];
            const D& wm = Wm[q * N + j];
            Yp = Yp + lps_plus[q] * wp + lps_minus[q] * wm;
            Ym = Ym + lps_plus[q] * wm + lps_minus[q] * wp;
        }
        const D& kj = k[j];
        D hp, hm;
        {
            const D den = 1.0 + kj * mu;
            if (fabs(den.v) > 0.0001)
                hp = (1.0 - theta[j] * att) / den;
            else
                hp = od * (1.0 / mu) * (1.0 - od * (kj + 1.0 / mu));
        }
        {
            const D den = 1.0 - kj * mu;
            if (fabs(den.v) > 0.0001)
                hm = od * (1.0 / mu) * dpsi(od, kj, D(1.0 / mu), theta[j], att);
            else
                hm = theta[j] * od * (1.0 / mu) * (1.0 - od * (kj - 1.0 / mu));
        }
        const D Dp = (E - trans_top * expfactor * hm) / (secant + kj);
        const D Dm = trans_top * (hp * mu - od * att * dpsi(od, kj, secant, theta[j], expfactor)) / (1.0 + secant * mu);
        src = src + Yp * hp * Lc[j] + Ym * hm * Mc[j] + Ap[j] * Yp * Dm + Am[j] * Ym * Dp;
    }
    return src + Q * E;
}

// Kernel-based BRDF, layer on the ground (every order reflects).  Kept out of line on the device: wf_layer_body is at
// the register allocator's limit (255 registers, 19 KB of stack at N = 8) and inlining these rarely taken paths changed
// the code generated for every other layer.
#if defined(__CUDA_ARCH__)
#define DISCO_NOINLINE __noinline__
#else
#define DISCO_NOINLINE
#endif
// downwelling field at the streams on the ground: X_q = G+bottom_q + sum_j (W+_qj Theta_j L_j + W-_qj M_j)
template <int N, int NL>
DISCO_HD DISCO_NOINLINE void wf_ground_field(const Dk<NL>* Gpb, const Dk<NL>* Wp, const Dk<NL>* Wm, const Dk<NL>* theta,
                                             const double* Lc, const double* Mc, Dk<NL>* Xg) {
    for (int q = 0; q < N; ++q) {
        Dk<NL> acc = Gpb[q];
        for (int j = 0; j < N; ++j) acc = acc + Wp[q * N + j] * theta[j] * Lc[j] + Wm[q * N + j] * Mc[j];
        Xg[q] = acc;
    }
}
// One line of sight: the reflection terms of the ground rows (adjoint part) and of the ground-leaving radiance, added
// to `tot`; the ground pieces of the cross-layer chain accumulated over the orders with their azimuth factor.
//   grow: [(N + nlos)][N + 1] rows (1 + delta_m0) w_q mu_q rho_m(row, mu_q) | rho_m(row, mu_0) of this (wavelength, order)
template <int N, int NL>
DISCO_HD DISCO_NOINLINE void wf_ground_general(const ChunkView& V, int w, int ms, int m, int los, const double* grow,
                                               const Dk<NL>* Xg, const double* zg, size_t zs, double attg, double t_floor,
                                               Dk<NL>& tot) {
    for (int q = 0; q < N; ++q) {
        double c = 0.0;   // ground rows: - sum_q R(i, q) X_q on the left-hand side
        for (int i = 0; i < N; ++i) c += zg[i * zs] * grow[i * (N + 1) + q];
        tot = tot + Xg[q] * c;
    }
    const double* lrow = grow + (size_t)(N + los) * (N + 1);
    Dk<NL> refl_los(0.0);
    for (int q = 0; q < N; ++q) refl_los = refl_los + Xg[q] * lrow[q];
    tot = tot + refl_los * attg;
    const double cf = V.T.los_cosmphi[(size_t)los * V.T.nstr + m];
    const double direct = V.include_ss ? V.T.csz / kPi * t_floor * lrow[N] : 0.0;
    double zsun = 0.0;   // right-hand side of the ground rows: csz rho(i, sun) t_floor / pi
    for (int i = 0; i < N; ++i) zsun += zg[i * zs] * grow[i * (N + 1) + N];
    const int nk = V.wf_gndk ? V.brdf_nk : 0;
    double* part = V.wf_gnd_part + ((((size_t)w * V.M + ms) * V.T.nlos + los) * (size_t)(2 + V.brdf_nk));
    part[0] = cf * ((V.include_ss ? attg * lrow[N] * V.T.csz / kPi : 0.0) + zsun * V.T.csz / kPi);
    part[1] = cf * attg * (direct + refl_los.v);
    // d/d(weight of kernel k): the BRDF is linear in the weights, so the partial derivative of the ground rows and of
    // the ground-leaving term at fixed solution is the same expression with kernel k's own Fourier coefficients
    const int nstr = V.T.nstr, nlos = V.T.nlos;
    const double sun = V.T.csz / kPi * t_floor;
    for (int k = 0; k < nk; ++k) {
        const double* Rs = V.brdf_Rss + (((size_t)k * nstr + m) * N) * N;          // [i][q]
        const double* rs = V.brdf_rsun + ((size_t)k * nstr + m) * N;               // [i]
        const double* Rl = V.brdf_Rls + (((size_t)k * nstr + m) * nlos + los) * N; // [q]
        const double rl = V.brdf_rlsun[((size_t)k * nstr + m) * nlos + los];
        double acc = V.include_ss ? sun * rl : 0.0;
        for (int q = 0; q < N; ++q) acc += Rl[q] * Xg[q].v;
        acc *= attg;
        for (int i = 0; i < N; ++i) {
            double row = sun * rs[i];
            for (int q = 0; q < N; ++q) row += Rs[i * N + q] * Xg[q].v;
            acc += zg[i * zs] * row;
        }
        part[2 + k] = cf * acc;
    }
}

// Sum of the per-order ground pieces in slot order: one (wavelength, LOS)
DISCO_HD void wf_ground_reduce_body(const ChunkView& V, long long idx) {
    const int nlos = V.T.nlos, M = V.M, stride = 2 + V.brdf_nk;
    const int w = (int)(idx / nlos), los = (int)(idx % nlos);
    double g1 = 0.0, g2 = 0.0, gk[4] = {0.0, 0.0, 0.0, 0.0};
    for (int ms = 0; ms < M; ++ms) {
        const double* part = V.wf_gnd_part + (((size_t)w * M + ms) * nlos + los) * (size_t)stride;
        g1 += part[0];
        g2 += part[1];
        for (int k = 0; k < V.brdf_nk && k < 4; ++k) gk[k] += part[2 + k];
    }
    double* gnd = V.wf_gnd + (size_t)idx * 3;
    gnd[0] = 0.0;
    gnd[1] = g1;
    gnd[2] = g2;
    if (V.wf_gndk)
        for (int k = 0; k < V.brdf_nk && k < 4; ++k) V.wf_gndk[(size_t)idx * V.brdf_nk + k] = gk[k];
}

// K5 body: one (wavelength, azimuth slot, layer).  G = number of scattering groups, NL = G + 4 local lanes
// ordered [eps_0..eps_{G-1} | tau | omega | t | s].
//   wf_loc [nw][M][nlos][L][NL]   d(I_m,los)/d(local lane of layer p) (direct + adjoint parts, LOS-attenuated)
//   wf_src [nw][M][nlos][L]       attenuated source of the layer (value), for the attenuation chain
//   wf_gnd [nw][nlos][3]          (m = 0) dI/d(albedo), dI/d(t at the ground), attenuated ground term (value)
template <int N, int G>
DISCO_HD void wf_layer_body(const ChunkView& V, long long idx) {
    constexpr int NSTR = 2 * N, NL = G + 4;
    constexpr int iTau = G, iOm = G + 1, iT = G + 2, iS = G + 3;
    using D = Dk<NL>;
    const int L = V.T.L, M = V.M, nlos = V.T.nlos;
    const int p = (int)(idx % L);
    const int ms = (int)((idx / L) % M);
    const int w = (int)(idx / ((long long)L * M));
    const int m = V.m_list[ms];
    const size_t wl = (size_t)w * L + p;
    const double* mu = V.T.mu;
    const double* wt = V.T.wt;
    const double* lp = V.T.lp_mu + (size_t)m * N * NSTR;

    D od(V.lay_od[wl]), ssa(V.lay_ssa[wl]), secant(V.lay_secant[wl]), trans_top(V.lay_trans[(size_t)w * (L + 1) + p]);
    od.d[iTau] = 1.0;
    ssa.d[iOm] = 1.0;
    trans_top.d[iT] = 1.0;
    secant.d[iS] = 1.0;
    D beta[NSTR];
    for (int l = 0; l < NSTR; ++l) {
        beta[l] = D(V.lay_beta[wl * NSTR + l]);
        for (int g = 0; g < G; ++g) beta[l].d[g] = V.lay_dbeta[(wl * G + g) * NSTR + l];
    }
    // stored homogeneous solution of this problem
    const double* Wp0 = V.Wp + (size_t)idx * N * N;
    const double* Wm0 = V.Wm + (size_t)idx * N * N;
    const double* kv = V.kth + (size_t)idx * 2 * N;
    double X[N * N], Xm[N * N], lam[N], nrm[N];
    for (int j = 0; j < N; ++j) lam[j] = kv[j] * kv[j];
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) {
            X[i * N + j] = Wp0[i * N + j] + Wm0[i * N + j];
            Xm[i * N + j] = kv[j] * (Wp0[i * N + j] - Wm0[i * N + j]);
        }
    for (int j = 0; j < N; ++j) {
        double s = 0.0;
        for (int a = 0; a < N; ++a) s += Xm[a * N + j] * wt[a] * mu[a] * X[a * N + j];
        nrm[j] = s;
    }
    double Spv[N * N];  // S+ (un-symmetrised): delta/mu_a - ssa w_b even_ab / mu_a
    for (int a = 0; a < N; ++a)
        for (int b = 0; b < N; ++b) {
            double even = 0.0;
            for (int l = m; l < NSTR; l += 2) even += beta[l].v * lp[a * NSTR + l] * lp[b * NSTR + l];
            Spv[a * N + b] = (a == b ? 1.0 / mu[a] : 0.0) - ssa.v * wt[b] * even / mu[a];
        }
    D k[N], Wp[N * N], Wm[N * N];
    for (int j = 0; j < N; ++j) k[j] = D(kv[j]);
    for (int i = 0; i < N * N; ++i) {
        Wp[i] = D(Wp0[i]);
        Wm[i] = D(Wm0[i]);
    }
    // eigen lanes: scattering groups and omega
    for (int e = 0; e <= G; ++e) {
        const int lane = (e < G) ? e : iOm;
        // ce_ab = d(omega * even_ab), co_ab = d(omega * odd_ab)  (symmetric)
        double ce[N * N], co[N * N];
        for (int a = 0; a < N; ++a)
            for (int b = 0; b <= a; ++b) {
                double even = 0.0, odd = 0.0;
                for (int l = m; l < NSTR; ++l) {
                    const double coef = (e < G) ? beta[l].d[e] * ssa.v : beta[l].v;
                    const double pp = coef * lp[a * NSTR + l] * lp[b * NSTR + l];
                    if ((l - m) & 1)
                        odd += pp;
                    else
                        even += pp;
                }
                ce[a * N + b] = ce[b * N + a] = even;
                co[a * N + b] = co[b * N + a] = odd;
            }
        // D^2 dS+_ab = -w_a w_b ce_ab,  D^2 dS-_ab = -w_a w_b co_ab
        // P_ij = ( Xm_i^T (D^2 dS-) Xm_j + lam_i X_i^T (D^2 dS+) X_j ) / n_i
        double Tm1[N * N], Tp1[N * N];  // (D^2 dS-) Xm and (D^2 dS+) X
        for (int a = 0; a < N; ++a)
            for (int j = 0; j < N; ++j) {
                double s1 = 0.0, s2 = 0.0;
                for (int b = 0; b < N; ++b) {
                    s1 -= wt[a] * wt[b] * co[a * N + b] * Xm[b * N + j];
                    s2 -= wt[a] * wt[b] * ce[a * N + b] * X[b * N + j];
                }
                Tm1[a * N + j] = s1;
                Tp1[a * N + j] = s2;
            }
        double P[N * N];
        for (int i = 0; i < N; ++i)
            for (int j = 0; j < N; ++j) {
                double s1 = 0.0, s2 = 0.0;
                for (int a = 0; a < N; ++a) {
                    s1 += Xm[a * N + i] * Tm1[a * N + j];
                    s2 += X[a * N + i] * Tp1[a * N + j];
                }
                P[i * N + j] = (s1 + lam[i] * s2) / nrm[i];
            }
        double dX[N * N], dXm[N * N], dk[N];
        for (int j = 0; j < N; ++j) dk[j] = P[j * N + j] / (2.0 * kv[j]);
        for (int a = 0; a < N; ++a)
            for (int j = 0; j < N; ++j) {
                double s = 0.0;
                for (int i = 0; i < N; ++i)
                    if (i != j) s += X[a * N + i] * P[i * N + j] / (lam[j] - lam[i]);
                dX[a * N + j] = s;
            }
        // dXm = dS+ X + S+ dX,  S+_ab = delta/mu_a - ssa w_b even_ab / mu_a,  dS+_ab = -(w_b/mu_a) ce_ab
        for (int a = 0; a < N; ++a)
            for (int j = 0; j < N; ++j) {
                double s = 0.0;
                for (int b = 0; b < N; ++b)
                    s += Spv[a * N + b] * dX[b * N + j] - (wt[b] / mu[a]) * ce[a * N + b] * X[b * N + j];
                dXm[a * N + j] = s;
            }
        for (int j = 0; j < N; ++j) {
            k[j].d[lane] = dk[j];
            for (int a = 0; a < N; ++a) {
                const double t1 = dXm[a * N + j] / kv[j] - Xm[a * N + j] * dk[j] / lam[j];
                Wp[a * N + j].d[lane] = 0.5 * (dX[a * N + j] + t1);
                Wm[a * N + j].d[lane] = 0.5 * (dX[a * N + j] - t1);
            }
        }
    }
    D theta[N];
    for (int j = 0; j < N; ++j) theta[j] = dexp(-(k[j] * od));
    D Ap[N], Am[N], Gpt[N], Gmt[N], Gpb[N], Gmb[N];
    particular_dual<N, NL>(V.T, m, od, ssa, beta, secant, trans_top, k, theta, Wp, Wm, Ap, Am, Gpt, Gmt, Gpb, Gmb);

    const double* x = V.xsol + (size_t)idx * 2 * N;
    const double* Lc = x;
    const double* Mc = x + N;
    // r1 = W+ L + (W- Theta) M ; r2 = W- L + (W+ Theta) M ; r3 = (W+ Theta) L + W- M ; r4 = (W- Theta) L + W+ M
    D r1[N], r2[N], r3[N], r4[N];
    for (int i = 0; i < N; ++i) {
        D a1(0.0), a2(0.0), a3(0.0), a4(0.0);
        for (int j = 0; j < N; ++j) {
            const D wpt = Wp[i * N + j] * theta[j];
            const D wmt = Wm[i * N + j] * theta[j];
            a1 = a1 + Wp[i * N + j] * Lc[j] + wmt * Mc[j];
            a2 = a2 + Wm[i * N + j] * Lc[j] + wpt * Mc[j];
            a3 = a3 + wpt * Lc[j] + Wm[i * N + j] * Mc[j];
            a4 = a4 + wmt * Lc[j] + Wp[i * N + j] * Mc[j];
        }
        r1[i] = a1;
        r2[i] = a2;
        r3[i] = a3;
        r4[i] = a4;
    }
    const bool bottom = (p == L - 1);
    // kernel-based BRDF (every order reflects): rows[(N + nlos)][N + 1] of this (wavelength, order) hold
    // (1 + delta_m0) w_q mu_q rho_m(row, mu_q) | rho_m(row, mu_0) for the stream rows and the LOS rows (k_surface_general)
    const bool general = bottom && V.gsurf_rows != nullptr && V.gsurf != nullptr;
    const double* grow = general ? V.gsurf_rows + ((size_t)w * M + ms) * (size_t)(N + nlos) * (N + 1) : nullptr;
    D Xg[N];
    if (general) wf_ground_field<N, NL>(Gpb, Wp, Wm, theta, Lc, Mc, Xg);
    const bool refl = bottom && (m == 0) && !general;
    const double albedo = V.albedo[w];
    D gsum(0.0);  // 2 sG + 2 sum_j (s+_j Theta_j L_j + s-_j M_j): the surface-reflected stream integral
    if (refl) {
        for (int q = 0; q < N; ++q) {
            const double f = 2.0 * wt[q] * mu[q];
            D acc = Gpb[q];
            for (int j = 0; j < N; ++j) acc = acc + Wp[q * N + j] * theta[j] * Lc[j] + Wm[q * N + j] * Mc[j];
            gsum = gsum + acc * f;
        }
    }
    const double cum_top = V.lay_cumod[(size_t)w * (L + 1) + p];
    const double cum_all = V.lay_cumod[(size_t)w * (L + 1) + L];
    const double t_floor = V.lay_trans[(size_t)w * (L + 1) + L];
    const size_t nrow = (size_t)2 * N * L;
    for (int los = 0; los < nlos; ++los) {
        // zadj is [row][los] (LOS fastest): element `row` of this LOS sits at z[row * nlos]
        const double* z = V.zadj + ((size_t)w * M + ms) * nrow * nlos + los;
        const size_t zs = (size_t)nlos;
        D adj(0.0);
        if (p == 0) {
            for (int i = 0; i < N; ++i) adj = adj - (Gpt[i] + r1[i]) * z[i * zs];
        } else {
            const double* zt = z + (N + (size_t)(p - 1) * 2 * N) * zs;
            for (int i = 0; i < N; ++i) adj = adj + (Gmt[i] + r2[i]) * zt[i * zs] + (Gpt[i] + r1[i]) * zt[(N + i) * zs];
        }
        double zg_sum = 0.0;
        if (!bottom) {
            const double* zb = z + (N + (size_t)p * 2 * N) * zs;
            for (int i = 0; i < N; ++i) adj = adj - (Gmb[i] + r4[i]) * zb[i * zs] - (Gpb[i] + r3[i]) * zb[(N + i) * zs];
        } else {
            const double* zg = z + (N + (size_t)(L - 1) * 2 * N) * zs;
            for (int i = 0; i < N; ++i) {
                adj = adj - (Gmb[i] + r4[i]) * zg[i * zs];
                zg_sum += zg[i * zs];
            }
            if (refl) adj = adj + gsum * (albedo * zg_sum);
        }
        const double mul = V.T.los_mu[los];
        const double att = exp(-cum_top / mul);
        D src = los_source_dual<N, NL>(V.T, m, los, od, ssa, beta, secant, trans_top, V.include_ss != 0, k, theta, Wp,
                                       Wm, Ap, Am, Lc, Mc);
        D tot = src * att + adj;
        const size_t o = (((size_t)w * M + ms) * nlos + los) * L + p;
        double srcval = src.v * att;
        if (refl) {
            const double attg = exp(-cum_all / mul);
            const double direct = V.include_ss ? V.T.csz / kPi * t_floor : 0.0;
            tot = tot + gsum * (attg * albedo);
            double* gnd = V.wf_gnd + ((size_t)w * nlos + los) * 3;
            gnd[0] = attg * (direct + gsum.v) + zg_sum * (V.T.csz * t_floor / kPi + gsum.v);  // d/d albedo
            gnd[1] = (V.include_ss ? attg * albedo * V.T.csz / kPi : 0.0) + zg_sum * V.T.csz * albedo / kPi;  // d/d t_floor
            gnd[2] = attg * albedo * (direct + gsum.v);                                        // ground term value
        }
        if (general)
            wf_ground_general<N, NL>(V, w, ms, m, los, grow, Xg, z + (N + (size_t)(L - 1) * 2 * N) * zs, zs, exp(-cum_all / mul), t_floor, tot);
        double* out = V.wf_loc + o * NL;
        for (int i = 0; i < NL; ++i) out[i] = tot.d[i];
        V.wf_src[o] = srcval;
    }
}

// K6a body: one (wavelength, LOS): azimuth sum, cross-layer chain (beam transmittance, average secant, LOS
// attenuation), layer -> native atmosphere derivatives following the reference's group_and_triangle_fraction
// weights (sktran_do_layerarray.cpp:487-652, 660-868; do_source_planeparallel.cpp:160-179).
//   scratch [nw][nlos][3][L+1] work arrays;  native [nw][nlos][nloc*(2+G)+1]
DISCO_HD void wf_chain_body(const ChunkView& V, long long idx, int G) {
    const int L = V.T.L, M = V.M, nlos = V.T.nlos, nstr = V.T.nstr, nloc = V.T.nloc;
    const int NL = G + 4, iTau = G, iOm = G + 1, iT = G + 2, iS = G + 3;
    const int w = (int)(idx / nlos), los = (int)(idx % nlos);
    const int nnative = nloc * (2 + G) + 1;
    double* native = V.wf_native + (size_t)idx * nnative;
    for (int i = 0; i < nnative; ++i) native[i] = 0.0;
    double* gT = V.wf_scratch + (size_t)idx * 3 * (L + 1);  // dI/dT_p, p = 0..L
    double* dtau = gT + (L + 1);                             // dI/dtau_p
    double* tail = dtau + (L + 1);                           // attenuated source of layer p (azimuth-summed)
    const double mul = V.T.los_mu[los];
    const double* od = V.lay_od + (size_t)w * L;
    const double* sec = V.lay_secant + (size_t)w * L;
    const double* tr = V.lay_trans + (size_t)w * (L + 1);
    const double* gnd = V.wf_gnd + (size_t)idx * 3;
    for (int p = 0; p <= L; ++p) gT[p] = 0.0;
    // azimuth sum of the local lanes
    for (int p = 0; p < L; ++p) {
        double loc_tau = 0.0, loc_om = 0.0, g_t = 0.0, g_s = 0.0, src = 0.0;
        for (int ms = 0; ms < M; ++ms) {
            const double cf = V.T.los_cosmphi[(size_t)los * nstr + V.m_list[ms]];
            const size_t o = (((size_t)w * M + ms) * nlos + los) * L + p;
            const double* lc = V.wf_loc + o * NL;
            loc_tau += cf * lc[iTau];
            loc_om += cf * lc[iOm];
            g_t += cf * lc[iT];
            g_s += cf * lc[iS];
            src += cf * V.wf_src[o];
        }
        dtau[p] = loc_tau - g_s * sec[p] / od[p];
        tail[p] = src;
        gT[p] += -tr[p] * g_t - g_s / od[p];
        gT[p + 1] += g_s / od[p];
        // omega lane and scattering lanes go straight to the native derivatives
        const double dh = V.layer_dh[p];
        (void)dh;
        for (int c = 0; c < 2; ++c) {
            const int q = V.interp_idx[p * 2 + c];
            if (q < 0) continue;
            const double wq = V.interp_w[p * 2 + c];
            const double kq = V.ext[(size_t)nloc * w + q], omq = V.ssa[(size_t)nloc * w + q];
            const double totext = V.lay_totext[(size_t)w * L + p], scatext = V.lay_scatext[(size_t)w * L + p];
            const double ssal = V.lay_ssa[(size_t)w * L + p];
            native[nloc + q] += wq * loc_om * (kq / totext);
            native[q] += wq * loc_om * ((omq - ssal) / totext);
            for (int g = 0; g < G; ++g) {
                double de = 0.0;
                for (int ms = 0; ms < M; ++ms) {
                    const double cf = V.T.los_cosmphi[(size_t)los * nstr + V.m_list[ms]];
                    const size_t o = (((size_t)w * M + ms) * nlos + los) * L + p;
                    de += cf * V.wf_loc[o * NL + g];
                }
                native[2 * nloc + g * nloc + q] += wq * de * (omq * kq / scatext);
            }
        }
    }
    // ground: t at the floor of the bottom layer (m = 0 only, cos(0) = 1)
    gT[L] += -tr[L] * gnd[1];
    // LOS attenuation: dI/dtau_q -= (1/mu) (sum_{p>q} src_p + ground)
    double below = gnd[2];
    for (int q = L - 1; q >= 0; --q) {
        dtau[q] -= below / mul;
        below += tail[q];
    }
    // slant optical depths: T_{p+1} = sum_{q<=p} chapman[p][q] tau_q
    for (int pp = 0; pp < L; ++pp) {
        const double g = gT[pp + 1];
        if (g == 0.0) continue;
        const double* ch = V.chapman + (size_t)pp * L;
        for (int q = 0; q <= pp; ++q) dtau[q] += g * ch[q];
    }
    for (int p = 0; p < L; ++p) {
        const double dh = V.layer_dh[p];
        for (int c = 0; c < 2; ++c) {
            const int q = V.interp_idx[p * 2 + c];
            if (q < 0) continue;
            native[q] += V.interp_w[p * 2 + c] * dh * dtau[p];
        }
    }
    native[nloc * (2 + G)] = gnd[0];
}

// K6b body: one (wavelength, LOS, mapping output index): OutputC::assign_lane (cpp/lib/output/outputc.cpp:37-160)
struct MappingView {
    const double* d_ssa;        // [nloc, nw] of the staged range
    const double* d_ext;        // [nloc, nw]
    const double* scat_factor;  // [nloc, nw] or null
    int scat_index;
    const double* interp;       // [nloc, nout] column-major or null
    int nout;
    double* out;                // [nout][nw][nlos]
};

DISCO_HD void wf_map_body(const ChunkView& V, const MappingView& Mp, int w0, int nw_total, long long idx, int G) {
    const int nlos = V.T.nlos, nloc = V.T.nloc;
    const int o = (int)(idx % Mp.nout);
    const long long wl = idx / Mp.nout;
    const int los = (int)(wl % nlos), w = (int)(wl / nlos);
    const int nnative = nloc * (2 + G) + 1;
    const double* native = V.wf_native + ((size_t)w * nlos + los) * nnative;
    const size_t wg = (size_t)(w0 + w);
    auto term = [&](int q) {
        double s = Mp.d_ssa[nloc * wg + q] * native[nloc + q] + Mp.d_ext[nloc * wg + q] * native[q];
        if (Mp.scat_factor) s += Mp.scat_factor[nloc * wg + q] * native[2 * nloc + Mp.scat_index * nloc + q];
        return s;
    };
    double acc = 0.0;
    if (Mp.interp) {
        for (int q = 0; q < nloc; ++q) {
            const double c = Mp.interp[(size_t)o * nloc + q];
            if (c != 0.0) acc += c * term(q);
        }
    } else {
        acc = term(o);
    }
    Mp.out[((size_t)o * nw_total + wg) * nlos + los] = acc;
}

}  // namespace disco
