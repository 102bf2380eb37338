// Per-layer discrete-ordinates math shared by the CUDA kernels (and compiled for the host by the
// emulation test in tests/host_emul.cpp).  fp64 throughout.
//
// What it computes follows the reference's per-layer solver (file:line relative to the reference tree):
//   S+/S- build            cpp/include/sktran_disco/sktran_do_lpproduct.h:165-263
//   homogeneous solution   cpp/lib/sktran_disco/sktran_do_rte.cpp:383-553
//   Green's particular     cpp/lib/sktran_disco/sktran_do_rte.cpp:556-580, 903-1332
//   LOS source multipliers cpp/lib/sktran_disco/sktran_do_opticallayer.cpp:94-555, 785-938
// but not how: the eigenproblem of S-S+ is solved through the symmetric similarity
//   D S+- D^-1 (D = diag sqrt(w_i mu_i)),  S~- = H H^T,  C = H^T S~+ H = Z diag(k^2) Z^T   (cyclic Jacobi)
//   X = D^-1 H Z,   S+ X = D^-1 (S~+ H) Z
// which needs no Hessenberg/QR iteration, has fixed control flow and keeps small k^2 relatively accurate.
// Eigenvector scale/sign is a free gauge of the radiance (A+- carry 1/norm, L/M absorb the rest).
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define DISCO_HD __host__ __device__ __forceinline__
#else
#define DISCO_HD inline
#endif

namespace disco {

constexpr double kPi = 3.14159265358979323846;
constexpr double kGreensEps = 1e-4;   // SKTRAN_DO_GREENS_EPS (sktran_do_types.h:11)
constexpr double kSsaDither = 1e-9;   // sktran_do_specs.h:104

// Particular-solution multipliers without removable singularities.  The reference evaluates C+, h- and D-
// (sktran_do_rte.cpp:1203-1226, sktran_do_opticallayer.cpp:339-344, 897-938) as differences of exponentials over
// (secant - k) or (1 - mu k); close to those degeneracies the values lose digits and the secant-derivatives lose
// twice as many, which the weighting-function chain then divides by the layer optical depth.  With
//   phi(x) = (1 - e^-x) / x,   psi(a; k1, k2) = (e^{-a k1} - e^{-a k2}) / (a (k2 - k1)) = e^{-a min(k1,k2)} phi(a |k2 - k1|)
// the same quantities are  C+ = t a psi(a; k, s),  h- = (a / mu) psi(a; k, 1/mu),
// D- = t (mu h+ - a e^{-a/mu} psi(a; k, s)) / (1 + mu s).  The reference's own Taylor branches (|secant - k| <= 1e-4,
// |1 - mu k| <= 1e-4) are kept verbatim, so results differ from the reference's only by the rounding noise of its
// direct formulas.
DISCO_HD double phi_value(double x) { return x == 0.0 ? 1.0 : -expm1(-x) / x; }
DISCO_HD double phi_deriv(double x) {
    if (fabs(x) > 0.01) return (exp(-x) - phi_value(x)) / x;
    double term = -0.5, sum = -0.5;  // phi'(x) = sum_n (-1)^(n+1) (n+1) x^n / (n+2)!
#pragma unroll
    for (int n = 1; n < 9; ++n) {
        term *= -x * (n + 1.0) / (n * (n + 2.0));
        sum += term;
    }
    return sum;
}
// e1 = exp(-a k1), e2 = exp(-a k2)
DISCO_HD double psi_value(double a, double k1, double k2, double e1, double e2) {
    return (k2 >= k1) ? e1 * phi_value(a * (k2 - k1)) : e2 * phi_value(a * (k1 - k2));
}

// Geometry tables (device or host pointers), see disco_plan.h for layouts
struct Tables {
    int nstr, N, L, nloc, nlos;
    double csz;
    const double* mu;      // [nstr]
    const double* wt;      // [nstr]
    const double* lp_mu;   // [m][i<N][l]
    const double* lp_csz;  // [m][l]
    const double* lp_los;  // [los][m][l]
    const double* los_mu;  // [nlos]
    const double* los_cosmphi;  // [nlos][m]  cos(m * azimuth)
    // per azimuth order, the shared-memory tables of k_wf_layer_fast in its own layout (null: other paths):
    // [m][ tW[nstr][N] | tM[nstr][N] | tL[nlos][nstr] | lpc[nstr] | wmu[N] ]
    const double* wf_tab;
    // [m][nlos] 1: P_l^m(mu_los) = 0 for every l (m > 0 at exactly nadir) - order m gives this line of sight nothing,
    // the register-resident layer kernels skip its multipliers and partials; null: not tabulated
    const unsigned char* los_zero;
};

// Output of the per-layer solve, thread-local
template <int N>
struct LayerSol {
    double k[N];          // eigenvalues (separation constants)
    double theta[N];      // exp(-k tau)
    double Wp[N * N];     // row-major W+(i,j): stream i, solution j
    double Wm[N * N];
    double Ap[N], Am[N];  // Green's function coefficients
    double Ath[N];        // thermal Green's coefficient (A+ = A-), only set by thermal_particular
    double Gpt[N], Gmt[N], Gpb[N], Gmb[N];
    int status;           // 0 ok, 1 Cholesky failed, 2 non-positive eigenvalue
};

// Symmetric cyclic Jacobi on C (full storage), accumulating the same column rotations into nacc extra
// row-major N x N matrices (acc0, acc1).  On exit C is diagonal (to working accuracy).
template <int N>
DISCO_HD void jacobi_eig(double* C, double* acc0, double* acc1) {
    for (int sweep = 0; sweep < 40; ++sweep) {
        int rotations = 0;
        for (int p = 0; p < N - 1; ++p) {
            for (int q = p + 1; q < N; ++q) {
                double apq = C[p * N + q];
                double app = C[p * N + p], aqq = C[q * N + q];
                if (fabs(apq) <= 1e-17 * sqrt(fabs(app * aqq))) {
                    C[p * N + q] = 0.0;
                    C[q * N + p] = 0.0;
                    continue;
                }
                ++rotations;
                double zeta = (aqq - app) / (2.0 * apq);
                double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                double c = 1.0 / sqrt(1.0 + t * t);
                double s = t * c;
                for (int r = 0; r < N; ++r) {  // columns p, q
                    double arp = C[r * N + p], arq = C[r * N + q];
                    C[r * N + p] = c * arp - s * arq;
                    C[r * N + q] = s * arp + c * arq;
                }
                for (int r = 0; r < N; ++r) {  // rows p, q
                    double apr = C[p * N + r], aqr = C[q * N + r];
                    C[p * N + r] = c * apr - s * aqr;
                    C[q * N + r] = s * apr + c * aqr;
                }
                C[p * N + q] = 0.0;
                C[q * N + p] = 0.0;
                for (int r = 0; r < N; ++r) {
                    double v0 = acc0[r * N + p], v1 = acc0[r * N + q];
                    acc0[r * N + p] = c * v0 - s * v1;
                    acc0[r * N + q] = s * v0 + c * v1;
                    double u0 = acc1[r * N + p], u1 = acc1[r * N + q];
                    acc1[r * N + p] = c * u0 - s * u1;
                    acc1[r * N + q] = s * u0 + c * u1;
                }
            }
        }
        if (rotations == 0) break;
    }
}

// Homogeneous + Green's particular solution of one layer for azimuth order m.
//   od: layer optical thickness, ssa: (dithered) single-scatter albedo, beta[2N]: layer Legendre moments,
//   secant / trans_top: pseudo-spherical beam average secant and transmittance at the layer ceiling.
template <int N>
DISCO_HD void layer_solve(const Tables& T, int m, double od, double ssa, const double* beta, double secant,
                          double trans_top, LayerSol<N>& S) {
    constexpr int NSTR = 2 * N;
    const double* lp = T.lp_mu + (size_t)m * N * NSTR;
    S.status = 0;
    double Sp[N * N], Sm[N * N];  // symmetrised S~+ and S~-
    double d[N];
    for (int i = 0; i < N; ++i) d[i] = sqrt(T.wt[i] * T.mu[i]);
    for (int i = 0; i < N; ++i) {
        for (int j = 0; j <= i; ++j) {
            double even = 0.0, odd = 0.0;
            for (int l = m; l < NSTR; ++l) {
                double pp = beta[l] * lp[i * NSTR + l] * lp[j * NSTR + l];
                if ((l - m) & 1)
                    odd += pp;
                else
                    even += pp;
            }
            double f = ssa * sqrt(T.wt[i] * T.wt[j] / (T.mu[i] * T.mu[j]));
            double sp = -f * even, sm = -f * odd;
            if (i == j) {
                sp += 1.0 / T.mu[i];
                sm += 1.0 / T.mu[i];
            }
            Sp[i * N + j] = Sp[j * N + i] = sp;
            Sm[i * N + j] = Sm[j * N + i] = sm;
        }
    }
    // Cholesky S~- = H H^T (H lower, stored in Hm)
    double H[N * N];
    for (int i = 0; i < N * N; ++i) H[i] = 0.0;
    for (int j = 0; j < N; ++j) {
        double s = Sm[j * N + j];
        for (int q = 0; q < j; ++q) s -= H[j * N + q] * H[j * N + q];
        if (!(s > 0.0)) {
            S.status = 1;
            s = 1e-300;
        }
        double hjj = sqrt(s);
        H[j * N + j] = hjj;
        for (int i = j + 1; i < N; ++i) {
            double t = Sm[i * N + j];
            for (int q = 0; q < j; ++q) t -= H[i * N + q] * H[j * N + q];
            H[i * N + j] = t / hjj;
        }
    }
    // Tm = S~+ H ;  C = H^T Tm
    double Tm[N * N], C[N * N];
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) {
            double s = 0.0;
            for (int q = j; q < N; ++q) s += Sp[i * N + q] * H[q * N + j];
            Tm[i * N + j] = s;
        }
    for (int i = 0; i < N; ++i)
        for (int j = 0; j <= i; ++j) {
            double s = 0.0;
            for (int q = i; q < N; ++q) s += H[q * N + i] * Tm[q * N + j];
            C[i * N + j] = C[j * N + i] = s;
        }
    // C = Z diag(k^2) Z^T ; H <- H Z (= D X), Tm <- Tm Z (= D S+ X)
    jacobi_eig<N>(C, H, Tm);
    for (int j = 0; j < N; ++j) {
        double ksq = C[j * N + j];
        if (!(ksq > 0.0)) {
            S.status = 2;
            ksq = fabs(ksq) + 1e-300;
        }
        S.k[j] = sqrt(ksq);
        S.theta[j] = exp(-S.k[j] * od);
    }
    for (int i = 0; i < N; ++i) {
        double di = 1.0 / d[i];
        for (int j = 0; j < N; ++j) {
            double x = H[i * N + j] * di;
            double xm = Tm[i * N + j] * di / S.k[j];
            S.Wp[i * N + j] = 0.5 * (x + xm);
            S.Wm[i * N + j] = 0.5 * (x - xm);
        }
    }
    // Green's function particular solution
    double Qp[N], Qm[N];
    const double* lpc = T.lp_csz + (size_t)m * NSTR;
    for (int i = 0; i < N; ++i) {
        double sp = 0.0, sm = 0.0;
        for (int l = m; l < NSTR; ++l) {
            double pp = beta[l] * lp[i * NSTR + l] * lpc[l];
            sp += pp;
            sm += ((l - m) & 1) ? -pp : pp;
        }
        double factor = (m == 0 ? 1.0 : 2.0) * (1.0 / (4.0 * kPi)) * T.wt[i] * ssa;
        Qp[i] = sp * factor;
        Qm[i] = sm * factor;
    }
    for (int i = 0; i < N; ++i) S.Gpt[i] = S.Gmt[i] = S.Gpb[i] = S.Gmb[i] = 0.0;
    double exp_sec = exp(-od * secant);
    for (int j = 0; j < N; ++j) {
        double norm = 0.0, ap = 0.0, am = 0.0;
        for (int i = 0; i < N; ++i) {
            double wp = S.Wp[i * N + j], wm = S.Wm[i * N + j];
            norm += T.wt[i] * T.mu[i] * (wp * wp - wm * wm);
            ap += Qp[i] * wp + Qm[i] * wm;
            am += Qm[i] * wp + Qp[i] * wm;
        }
        ap /= norm;
        am /= norm;
        S.Ap[j] = ap;
        S.Am[j] = am;
        double kj = S.k[j], exp_k = S.theta[j];
        double Cp, Cm;
        if (fabs(secant - kj) > kGreensEps)
            Cp = trans_top * od * psi_value(od, kj, secant, exp_k, exp_sec);
        else
            Cp = trans_top * exp_k * od * (1.0 - od / 2.0 * (secant - kj));
        if (fabs(secant + kj) > kGreensEps)
            Cm = trans_top * (1.0 - exp_sec * exp_k) / (secant + kj);
        else
            Cm = trans_top * od * (1.0 - od / 2.0 * (secant + kj));
        double amc = am * Cm, apc = ap * Cp;
        for (int i = 0; i < N; ++i) {
            S.Gpt[i] += amc * S.Wm[i * N + j];
            S.Gmt[i] += amc * S.Wp[i * N + j];
            S.Gpb[i] += apc * S.Wp[i * N + j];
            S.Gmb[i] += apc * S.Wm[i * N + j];
        }
    }
}

// Thermal source S(x) = b0 exp(-b1 x) of a layer (order 0 only): Green's function particular solution added on top of
// the solar one (RTESolver::solveParticularGreenThermal, sktran_do_rte.cpp:1335-1617).  The source is isotropic, so
// Q+- = w_i (1 - ssa) and A+ = A-.
template <int N>
DISCO_HD void thermal_particular(const Tables& T, double od, double ssa, double b0, double b1, LayerSol<N>& S) {
    const double e_b1 = exp(-od * b1);
    for (int j = 0; j < N; ++j) {
        double norm = 0.0, a = 0.0;
        for (int i = 0; i < N; ++i) {
            const double wp = S.Wp[i * N + j], wm = S.Wm[i * N + j];
            norm += T.wt[i] * T.mu[i] * (wp * wp - wm * wm);
            a += T.wt[i] * (1.0 - ssa) * (wp + wm);
        }
        a /= norm;
        S.Ath[j] = a;
        const double kj = S.k[j], e_k = S.theta[j];
        double Cp, Cm;
        if (fabs(b1 - kj) > kGreensEps)
            Cp = b0 * (e_k - e_b1) / (b1 - kj);
        else
            Cp = b0 * e_k * od * (1.0 - od / 2.0 * (b1 - kj));
        if (fabs(b1 + kj) > kGreensEps)
            Cm = b0 * (1.0 - e_b1 * e_k) / (b1 + kj);
        else
            Cm = b0 * od * (1.0 - od / 2.0 * (b1 + kj));
        const double amc = a * Cm, apc = a * Cp;
        for (int i = 0; i < N; ++i) {
            S.Gpt[i] += amc * S.Wm[i * N + j];
            S.Gmt[i] += amc * S.Wp[i * N + j];
            S.Gpb[i] += apc * S.Wp[i * N + j];
            S.Gmb[i] += apc * S.Wm[i * N + j];
        }
    }
}

// Source-function multipliers of one layer toward one line of sight (observer above the atmosphere):
//   source = sum_j cpos[j] L_j + cneg[j] M_j + v
// i.e. cpos = Y+ h+, cneg = Y- h-, v = V + Q E  (sktran_do_opticallayer.cpp:284-321, 384-393, 513).
template <int N>
DISCO_HD void los_layer_terms(const Tables& T, int m, int los, double od, double ssa, const double* beta,
                              double secant, double trans_top, bool include_ss, const LayerSol<N>& S, double* cpos,
                              double* cneg, double& v, bool thermal = false, double b0 = 0.0, double b1 = 0.0) {
    constexpr int NSTR = 2 * N;
    const double mu = T.los_mu[los];
    const double* lp = T.lp_mu + (size_t)m * N * NSTR;
    const double* lpl = T.lp_los + ((size_t)los * NSTR + m) * NSTR;
    const double* lpc = T.lp_csz + (size_t)m * NSTR;
    double lps_plus[N], lps_minus[N];
    for (int q = 0; q < N; ++q) {
        double a = 0.0, b = 0.0;
        for (int l = m; l < NSTR; ++l) {
            double pp = beta[l] * lpl[l] * lp[q * NSTR + l];
            a += pp;
            b += ((l - m) & 1) ? -pp : pp;
        }
        lps_minus[q] = a * 0.5 * T.wt[q] * ssa;
        lps_plus[q] = b * 0.5 * T.wt[q] * ssa;
    }
    double Q = 0.0;
    if (include_ss) {
        double acc = 0.0;
        for (int l = m; l < NSTR; ++l) {
            double pp = beta[l] * lpl[l] * lpc[l];
            acc += ((l - m) & 1) ? -pp : pp;
        }
        Q = acc * (m == 0 ? 1.0 : 2.0) * (1.0 / (4.0 * kPi)) * ssa;
    }
    double att = exp(-od / mu);
    double expfactor = exp(-od * secant);
    double E = trans_top / (1.0 + mu * secant) * (1.0 - expfactor * att);
    // thermal source (order 0): E_thermal at x = 0 (sktran_do_opticallayer.cpp:941-957)
    const double e_b1 = thermal ? exp(-od * b1) : 0.0;
    const double E_th = thermal ? b0 / (1.0 + mu * b1) * (1.0 - e_b1 * att) : 0.0;
    double V = 0.0;
    for (int j = 0; j < N; ++j) {
        double Yp = 0.0, Ym = 0.0;
        for (int q = 0; q < N; ++q) {
            double wp = S.Wp[q * N + j], wm = S.Wm[q * N + j];
            Yp += lps_plus[q] * wp + lps_minus[q] * wm;
            Ym += lps_plus[q] * wm + lps_minus[q] * wp;
        }
        double k = S.k[j];
        double hp, hm;
        {
            double den = 1.0 + mu * k;
            if (fabs(den) > 0.0001)
                hp = (1.0 - S.theta[j] * att) / den;
            else
                hp = od / mu * (1.0 - od * (k + 1.0 / mu));
        }
        {
            double den = 1.0 - mu * k;
            if (fabs(den) > 0.0001)
                hm = od / mu * psi_value(od, k, 1.0 / mu, S.theta[j], att);
            else
                hm = S.theta[j] * od / mu * (1.0 - od * (k - 1.0 / mu));
        }
        cpos[j] = Yp * hp;
        cneg[j] = Ym * hm;
        double Dp = (-trans_top * expfactor * hm + E) / (secant + k);
        double Dm = trans_top * (mu * hp - od * att * psi_value(od, k, secant, S.theta[j], expfactor)) / (1.0 + mu * secant);
        V += S.Ap[j] * Yp * Dm + S.Am[j] * Ym * Dp;
        if (thermal) {  // sktran_do_opticallayer.cpp:421-478 (the reference has no series branch for b1 -> k here)
            const double Dp_th = (-b0 * e_b1 * hm + E_th) / (b1 + k);
            const double Dm_th = (b0 * hp - E_th) / (b1 - k);
            V += S.Ath[j] * (Yp * Dm_th + Ym * Dp_th);
        }
    }
    v = V + Q * E;
    if (thermal) v += E_th * (1.0 - ssa);  // :524-531
}

}  // namespace disco
