// =====================================================================================================
//  ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the shipped product.
//
//  CPU restatement (plain C++17, no Eigen) of SASKTRAN2's scalar (NSTOKES=1) discrete-ordinates
//  radiance solve for plane-parallel / pseudo-spherical geometry with a Lambertian surface.
//  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.
//
//  Every function cites the reference file:line (relative to /root/reference) whose arithmetic it
//  restates.  Parity pin: tests/test_oracle_golden.py checks this code against the reference's own
//  DISORT-verified tables (cpp/lib/tests/sktran_disco/legacy/test_scalar.cpp, abs tol 1e-8).
//
//  Third-party arithmetic not under /root/reference: the reference eigen-solves S^-S^+ with Eigen 3.4's
//  real Schur solver or LAPACK dgeev (sktran_do_rte.cpp:437-524); the oracle calls LAPACK dgeev from the
//  OpenBLAS 0.3.x bundled with scipy (symbol scipy_dgeev_), handed in as a function pointer.  The banded
//  BVP solve restates LAPACK dgbtf2/dgbtrs (== the reference's in-tree dgbtf2_unblocked,
//  cpp/lib/sktran_disco/sktran_do_banded_lu.cpp:8-146).
//
//  Weighting functions: forward-mode dual numbers over the reference's *layer-level* derivative lanes
//  (per layer: one lane per scattering group, optical depth, SSA; + albedo on the last layer —
//  sktran_do_layerarray.cpp:487-652) followed by the reference's layer->native mapping (:660-868,
//  do_source_planeparallel.cpp:154-180), including its quirks (see map_to_native()).
// =====================================================================================================
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <vector>

namespace oracle {

constexpr double PI = 3.14159265358979323846;
constexpr double GREENS_EPS = 1e-4;  // SKTRAN_DO_GREENS_EPS, sktran_do_types.h:11
constexpr double SSA_DITHER = 1e-9;  // sktran_do_specs.h:104

// LAPACK dgeev signature (Fortran ABI)
typedef void (*dgeev_fn)(const char* jobvl, const char* jobvr, const int* n, double* a, const int* lda,
                         double* wr, double* wi, double* vl, const int* ldvl, double* vr, const int* ldvr,
                         double* work, const int* lwork, int* info);

// ---------------------------------------------------------------------------------------------------
//  Dual numbers (dense derivative vector of runtime length g_nd)
// ---------------------------------------------------------------------------------------------------
inline int& nd_ref() {
    static thread_local int nd = 0;
    return nd;
}

struct Dual {
    double v;
    std::vector<double> d;
    Dual() : v(0.0), d(nd_ref(), 0.0) {}
    Dual(double x) : v(x), d(nd_ref(), 0.0) {}  // NOLINT implicit on purpose
    Dual& operator+=(const Dual& o) {
        v += o.v;
        for (size_t i = 0; i < d.size(); ++i) d[i] += o.d[i];
        return *this;
    }
    Dual& operator-=(const Dual& o) {
        v -= o.v;
        for (size_t i = 0; i < d.size(); ++i) d[i] -= o.d[i];
        return *this;
    }
    Dual& operator*=(const Dual& o) {
        for (size_t i = 0; i < d.size(); ++i) d[i] = d[i] * o.v + v * o.d[i];
        v *= o.v;
        return *this;
    }
    Dual& operator/=(const Dual& o) {
        double q = v / o.v;
        for (size_t i = 0; i < d.size(); ++i) d[i] = (d[i] - q * o.d[i]) / o.v;
        v = q;
        return *this;
    }
};
inline Dual operator+(Dual a, const Dual& b) { return a += b; }
inline Dual operator-(Dual a, const Dual& b) { return a -= b; }
inline Dual operator*(Dual a, const Dual& b) { return a *= b; }
inline Dual operator/(Dual a, const Dual& b) { return a /= b; }
inline Dual operator-(Dual a) {
    a.v = -a.v;
    for (auto& x : a.d) x = -x;
    return a;
}
inline Dual exp(const Dual& a) {
    Dual r;
    r.v = std::exp(a.v);
    for (size_t i = 0; i < a.d.size(); ++i) r.d[i] = r.v * a.d[i];
    return r;
}
inline Dual sqrt(const Dual& a) {
    Dual r;
    r.v = std::sqrt(a.v);
    for (size_t i = 0; i < a.d.size(); ++i) r.d[i] = 0.5 * a.d[i] / r.v;
    return r;
}
inline double val(double x) { return x; }
inline double val(const Dual& x) { return x.v; }

// Fixed-size dual number: the K "layer-local" derivative lanes of the reverse-mode linearisation (ReverseSolver
// below).  Same member names as Dual so that the layer routines are shared.
template <int K>
struct FDual {
    double v;
    double d[K];
    FDual() : v(0.0) {
        for (int i = 0; i < K; ++i) d[i] = 0.0;
    }
    FDual(double x) : v(x) {  // NOLINT implicit on purpose
        for (int i = 0; i < K; ++i) d[i] = 0.0;
    }
    FDual& operator+=(const FDual& o) {
        v += o.v;
        for (int i = 0; i < K; ++i) d[i] += o.d[i];
        return *this;
    }
    FDual& operator-=(const FDual& o) {
        v -= o.v;
        for (int i = 0; i < K; ++i) d[i] -= o.d[i];
        return *this;
    }
    FDual& operator*=(const FDual& o) {
        for (int i = 0; i < K; ++i) d[i] = d[i] * o.v + v * o.d[i];
        v *= o.v;
        return *this;
    }
    FDual& operator/=(const FDual& o) {
        const double q = v / o.v;
        for (int i = 0; i < K; ++i) d[i] = (d[i] - q * o.d[i]) / o.v;
        v = q;
        return *this;
    }
};
template <int K> inline FDual<K> operator+(FDual<K> a, const FDual<K>& b) { return a += b; }
template <int K> inline FDual<K> operator-(FDual<K> a, const FDual<K>& b) { return a -= b; }
template <int K> inline FDual<K> operator*(FDual<K> a, const FDual<K>& b) { return a *= b; }
template <int K> inline FDual<K> operator/(FDual<K> a, const FDual<K>& b) { return a /= b; }
template <int K> inline FDual<K> operator-(FDual<K> a) {
    a.v = -a.v;
    for (int i = 0; i < K; ++i) a.d[i] = -a.d[i];
    return a;
}
template <int K> inline FDual<K> exp(const FDual<K>& a) {
    FDual<K> r;
    r.v = std::exp(a.v);
    for (int i = 0; i < K; ++i) r.d[i] = r.v * a.d[i];
    return r;
}
template <int K> inline FDual<K> sqrt(const FDual<K>& a) {
    FDual<K> r;
    r.v = std::sqrt(a.v);
    for (int i = 0; i < K; ++i) r.d[i] = 0.5 * a.d[i] / r.v;
    return r;
}
template <int K> inline double val(const FDual<K>& x) { return x.v; }
// number of derivative lanes carried by a value
inline int ndual(const double&) { return 0; }
inline int ndual(const Dual& x) { return (int)x.d.size(); }
template <int K> inline int ndual(const FDual<K>&) { return K; }
inline double exp(double x) { return std::exp(x); }
inline double sqrt(double x) { return std::sqrt(x); }

// ---- "stable multipliers" variant (NOT the reference's formulas; off by default) -------------------------
// The reference evaluates the particular-solution multipliers C+ (sktran_do_rte.cpp:1203-1226), D- and h-
// (sktran_do_opticallayer.cpp:339-344, 897-938) as differences of exponentials divided by (secant - k) or
// (1 - mu k); within ~1e-3 of those degeneracies the values lose digits and the derivatives w.r.t. the secant
// lose twice as many, which the layer -> optical-depth chain then divides by the layer optical depth.  With
//   phi(x) = (1 - exp(-x)) / x,   psi(a; k1, k2) = (e^{-a k1} - e^{-a k2}) / (a (k2 - k1)) = e^{-a min} phi(a |k2 - k1|)
// the same quantities are  C+ = t a psi(a; k, s),  h- = (a/mu) psi(a; k, 1/mu),
// D- = t (mu h+ - a e^{-a/mu} psi(a; k, s)) / (1 + mu s), without a removable singularity.  Where the reference
// switches to its own Taylor branches (|secant - k| <= 1e-4, |1 - mu k| <= 1e-4) those are kept verbatim, so the
// two variants differ only by the rounding noise of the reference's direct formulas.  Used by the tests to tell
// the CUDA path's accuracy apart from that noise.
inline int& stable_multipliers_ref() {
    static int flag = 0;
    return flag;
}
inline double phi_value(double x) { return x == 0.0 ? 1.0 : -std::expm1(-x) / x; }
inline double phi_deriv(double x) {
    if (std::abs(x) > 0.01) return (std::exp(-x) - phi_value(x)) / x;
    // phi'(x) = sum_n (-1)^(n+1) (n+1) x^n / (n+2)!
    double term = -0.5, sum = -0.5;
    for (int n = 1; n < 9; ++n) {
        term *= -x * (n + 1.0) / (n * (n + 2.0));
        sum += term;
    }
    return sum;
}
inline double phi(double x) { return phi_value(x); }
inline Dual phi(const Dual& x) {
    Dual r;
    r.v = phi_value(x.v);
    const double dp = phi_deriv(x.v);
    for (size_t i = 0; i < x.d.size(); ++i) r.d[i] = dp * x.d[i];
    return r;
}
template <int K> inline FDual<K> phi(const FDual<K>& x) {
    FDual<K> r;
    r.v = phi_value(x.v);
    const double dp = phi_deriv(x.v);
    for (int i = 0; i < K; ++i) r.d[i] = dp * x.d[i];
    return r;
}
// psi(a; k1, k2) with e1 = exp(-a k1), e2 = exp(-a k2) supplied by the caller
template <class T>
inline T psi(const T& a, const T& k1, const T& k2, const T& e1, const T& e2) {
    if (val(k2) >= val(k1)) return e1 * phi(a * (k2 - k1));
    return e2 * phi(a * (k1 - k2));
}

// ---------------------------------------------------------------------------------------------------
//  Geometry-only plan
// ---------------------------------------------------------------------------------------------------
// Gauss-Legendre nodes/weights on (-1,1), ascending.  The reference takes them from the gauss-quad 0.2.4
// crate or from 25-digit tables (sktran_do_quadrature.cpp:25-63,105-225); Newton on P_n reproduces both
// to rounding.
inline void gauss_legendre(int n, std::vector<double>& x, std::vector<double>& w) {
    x.assign(n, 0.0);
    w.assign(n, 0.0);
    for (int i = 0; i < (n + 1) / 2; ++i) {
        long double z = std::cos(3.14159265358979323846264338327950288L * (i + 0.75L) / (n + 0.5L));
        long double pp = 0;
        for (int it = 0; it < 100; ++it) {
            long double p1 = 1.0L, p2 = 0.0L;
            for (int j = 1; j <= n; ++j) {
                long double p3 = p2;
                p2 = p1;
                p1 = ((2.0L * j - 1.0L) * z * p2 - (j - 1.0L) * p3) / j;
            }
            pp = n * (z * p1 - p2) / (z * z - 1.0L);
            long double z1 = z;
            z = z1 - p1 / pp;
            if (fabsl(z - z1) < 1e-19L) break;
        }
        x[i] = (double)(-z);
        x[n - 1 - i] = (double)z;
        w[i] = (double)(2.0L / ((1.0L - z * z) * pp * pp));
        w[n - 1 - i] = w[i];
    }
}

// Double-Gauss streams: sktran_do_quadrature.cpp:5-72.  First N entries are the upwelling (mu>0) streams.
inline void streams_and_weights(int nstr, std::vector<double>& mu, std::vector<double>& wt) {
    mu.assign(nstr, 0.0);
    wt.assign(nstr, 0.0);
    int order = nstr / 2;
    if (nstr == 2) {  // :16-23
        mu[0] = 0.5;
        mu[1] = -0.5;
        wt[0] = wt[1] = 1.0;
        return;
    }
    std::vector<double> x, w;
    gauss_legendre(order, x, w);
    for (int i = 0; i < order; ++i) {  // :65-71
        mu[i] = 0.5 * x[i] + 0.5;
        mu[i + order] = -0.5 * x[i] - 0.5;
        wt[i] = 0.5 * w[i];
        wt[i + order] = 0.5 * w[i];
    }
}

// Wigner d^l_{m0}(theta), cpp/include/sasktran2/math/wigner.h:56-169 (n = 0 branch).
inline double wigner_d_m0(int m, int l, double theta) {
    if (l < m) return 0.0;
    int zeta = (m == 0) ? 1 : ((m % 2 == 0) ? 1 : -1);  // :103-112 with n = 0
    // recurrence_start_factor(): (2m)! / (m! m!) built by the same descending loop (:56-82)
    double factorial = 1;
    for (int i = 2 * m; i > 1; --i) {
        factorial *= double(i);
        if (i <= m) factorial /= double(i);
        if (i <= m) factorial /= double(i);
    }
    double start_factor = zeta * std::pow(2.0, -double(m)) * std::sqrt(factorial);
    double x = std::cos(theta);
    double val_l = start_factor * std::pow(1 - x, double(m) / 2.0) * std::pow(1 + x, double(m) / 2.0);
    double val_lm1 = 0.0;
    for (int lidx = m + 1; lidx <= l; ++lidx) {  // :137-149
        double multiplier = 1.0 / (std::sqrt(double(lidx * lidx - m * m)) * lidx);
        double curfactor = (2 * lidx - 1) * (lidx * x);
        double priorfactor = lidx * std::sqrt(double((lidx - 1) * (lidx - 1) - m * m));
        double temp = val_l;
        val_l = multiplier * (curfactor * val_l - priorfactor * val_lm1);
        val_lm1 = temp;
    }
    return val_l;
}

struct Plan {
    int nstr = 0, N = 0, L = 0, nloc = 0, nlos = 0;
    double csz = 0;
    std::vector<double> mu, wt;     // [nstr]
    std::vector<double> lp_mu;      // [m][i<N][l]   sktran_do_specs.cpp:58-103
    std::vector<double> lp_csz;     // [m][l]        sktran_do_pconfig.cpp:41-46
    std::vector<double> lp_los;     // [los][m][l]   do_source_planeparallel.cpp:631-639
    std::vector<double> los_mu, los_az;
    std::vector<double> ceil_h, floor_h;  // [L]
    std::vector<double> W;                // [L][nloc] optical interpolator
    std::vector<double> chapman;          // [L][L]
    double LPmu(int m, int i, int l) const { return lp_mu[(size_t(m) * N + i) * nstr + l]; }
    double LPcsz(int m, int l) const { return lp_csz[size_t(m) * nstr + l]; }
    double LPlos(int j, int m, int l) const { return lp_los[(size_t(j) * nstr + m) * nstr + l]; }
};

// Grid::calculate_interpolation_weights, cpp/lib/grids/grid.cpp:43-300 (in-bounds part only; the layer
// mid-points are always inside the grid).  interp: 0 shell, 1 linear, 2 lower.
inline void interp_weights(const std::vector<double>& g, int interp, double x, int idx[2], double w[2], int& n) {
    int ng = (int)g.size();
    if (interp == 2) {  // :237 (+0.1 m tolerance)
        for (int i = 0; i < ng - 1; ++i) {
            if (x + 0.1 >= g[i] && x < g[i + 1]) {
                idx[0] = i; idx[1] = 0; w[0] = 1.0; w[1] = 0.0; n = 1;
                return;
            }
        }
        throw std::runtime_error("interp_weights: out of bounds (lower)");
    }
    // automatic spacing detection, grid.cpp:9-27 (Eigen isApproxToConstant, prec 1e-12)
    bool constant = true;
    double d0 = g[1] - g[0];
    for (int i = 1; i < ng; ++i) {
        double di = g[i] - g[i - 1];
        if (std::abs(di - d0) > 1e-12 * std::min(std::abs(di), std::abs(d0))) constant = false;
    }
    int i;
    if (constant) {  // :68-127
        i = int(std::floor((x - g[0]) / d0));
        if (i >= ng - 1) throw std::runtime_error("interp_weights: out of bounds");
        idx[0] = i; idx[1] = i + 1; n = 2;
        if (interp == 1) {
            w[1] = (x - g[i]) / d0;
            w[0] = 1 - w[1];
        } else {
            w[0] = w[1] = 0.5;
        }
    } else {  // :185-212
        i = int(std::lower_bound(g.begin(), g.end(), x) - g.begin());
        if (i == 0) i += 1;
        idx[0] = i - 1; idx[1] = i; n = 2;
        if (interp == 0) {
            w[0] = w[1] = 0.5;
        } else {
            w[1] = (x - g[i - 1]) / (g[i] - g[i - 1]);
            w[0] = 1 - w[1];
        }
    }
}

// ---- chapman factors by ray tracing: GeometryLayerArray::calculate_chapman_factors_raytracer
//      (sktran_do_geometrylayerarray.cpp:122-186) over SphericalShellRayTracer::trace_ray for an observer inside the
//      atmosphere looking up (cpp/lib/raytracing/spherical_shell.cpp:6-76, 424-455: complete shells from the top of
//      the atmosphere down to the first grid altitude above the observer, then one partial shell) with straight-line
//      layer distances (:193-215) and the pseudo-spherical coordinates of cpp/lib/geometry/geometry.cpp:8-35, 155-160
//      (observer on the z axis at the layer floor, look vector = sun unit vector).  A row of -1 flags "the sun ray
//      hits the ground" (cos_sza <= 0 is not traced here: the reference's looking-down branches are not restated).
struct Vec3 {
    double x, y, z;
    Vec3 operator+(const Vec3& o) const { return {x + o.x, y + o.y, z + o.z}; }
    Vec3 operator*(double f) const { return {x * f, y * f, z * f}; }
    double dot(const Vec3& o) const { return x * o.x + y * o.y + z * o.z; }
    double norm() const { return std::sqrt(x * x + y * y + z * z); }
};
inline void chapman_raytraced(const std::vector<double>& alt, const std::vector<double>& floor_h,
                              const std::vector<double>& ceil_h, double cos_sza, double earth_radius,
                              std::vector<double>& chapman) {
    const int L = (int)floor_h.size(), ng = (int)alt.size();
    chapman.assign(size_t(L) * L, 0.0);
    if (!(cos_sza > 0)) throw std::runtime_error("chapman_raytraced: cos_sza must be positive");
    const Vec3 sun{std::sqrt(1 - cos_sza * cos_sza), 0.0, cos_sza};  // geometry.cpp:19-22 with saa = 0
    struct Shell { double r_entrance, r_exit; };
    for (int p = 0; p < L; ++p) {
        const Vec3 obs{0.0, 0.0, floor_h[p] + earth_radius};              // solar_coordinate_vector, geometry.cpp:157-160
        const double robs = obs.norm();
        const double cosv = (obs * (1.0 / robs)).dot(sun);               // ViewingRay::cos_viewing
        const double rt = robs * std::sqrt(std::max(0.0, 1 - cosv * cosv));  // spherical_shell.cpp:17-19
        // trace_ray_observer_inside_looking_up (:424-455)
        const int start_index = int(std::upper_bound(alt.begin(), alt.end(), robs - earth_radius) - alt.begin());
        std::vector<Shell> layers;
        for (int i = ng - 1; i != start_index; --i)                       // complete_layer(exit i, ViewingDirection::up = -1), :254-276
            layers.push_back({alt[i - 1] + earth_radius, alt[i] + earth_radius});
        layers.push_back({robs, alt[start_index] + earth_radius});       // partial_layer (:278-298)
        // finalize_ray_geometry (:85-215): from the observer outwards, straight ray
        Vec3 entrance = obs;
        for (int i = 0; i < (int)layers.size(); ++i) {
            const Shell& ly = layers[layers.size() - i - 1];
            const double dist = std::abs(std::sqrt(std::fmax(ly.r_entrance * ly.r_entrance - rt * rt, 0.0)) -
                                         std::sqrt(std::fmax(ly.r_exit * ly.r_exit - rt * rt, 0.0)));
            const Vec3 exitp = entrance + sun * dist;
            const double average_altitude = (entrance.norm() + exitp.norm()) / 2.0 - earth_radius;
            int q = 0;
            for (; q < L; ++q)
                if (average_altitude >= floor_h[q] && average_altitude <= ceil_h[q]) break;
            if (q == L) q = L - 1;
            chapman[size_t(p) * L + q] += dist * 1.0 / (ceil_h[q] - floor_h[q]);
            entrance = exitp;
        }
    }
}

// geotype: 0 plane-parallel, 1 pseudo-spherical.  sktran_do_geometrylayerarray.cpp:8-119.
// Pseudo-spherical chapman factors: ray traced like the reference (chapman_raytraced above) unless
// chapman_straight_line_ref() is set, which selects the closed formula of calculate_chapman_factors (:69-119).
inline int& chapman_straight_line_ref() {
    static int flag = 0;
    return flag;
}
inline Plan make_plan(int nstr, const std::vector<double>& alt, int interp, int geotype, double cos_sza,
                      double earth_radius, const std::vector<double>& los_cos_vza,
                      const std::vector<double>& los_rel_az) {
    Plan P;
    P.nstr = nstr;
    P.N = nstr / 2;
    P.nloc = (int)alt.size();
    P.L = P.nloc - 1;
    P.nlos = (int)los_cos_vza.size();
    P.csz = cos_sza;
    streams_and_weights(nstr, P.mu, P.wt);
    P.lp_mu.assign(size_t(nstr) * P.N * nstr, 0.0);
    P.lp_csz.assign(size_t(nstr) * nstr, 0.0);
    P.lp_los.assign(size_t(P.nlos) * nstr * nstr, 0.0);
    for (int m = 0; m < nstr; ++m)
        for (int l = 0; l < nstr; ++l) {
            for (int i = 0; i < P.N; ++i) P.lp_mu[(size_t(m) * P.N + i) * nstr + l] = wigner_d_m0(m, l, std::acos(P.mu[i]));
            P.lp_csz[size_t(m) * nstr + l] = wigner_d_m0(m, l, std::acos(cos_sza));
            for (int j = 0; j < P.nlos; ++j)
                P.lp_los[(size_t(j) * nstr + m) * nstr + l] = wigner_d_m0(m, l, std::acos(los_cos_vza[j]));
        }
    P.los_mu = los_cos_vza;
    P.los_az.resize(P.nlos);
    for (int j = 0; j < P.nlos; ++j) P.los_az[j] = -los_rel_az[j];  // do_source_planeparallel.cpp:619
    P.ceil_h.resize(P.L);
    P.floor_h.resize(P.L);
    for (int p = 0; p < P.L; ++p) {  // :20-28
        P.ceil_h[p] = alt[P.nloc - 1 - p];
        P.floor_h[p] = alt[P.nloc - 2 - p];
    }
    P.W.assign(size_t(P.L) * P.nloc, 0.0);
    for (int p = 0; p < P.L; ++p) {  // :43-55
        double c = (P.ceil_h[p] + P.floor_h[p]) / 2.0;
        int idx[2];
        double w[2];
        int n;
        interp_weights(alt, interp, c, idx, w, n);
        for (int q = 0; q < n; ++q) P.W[size_t(p) * P.nloc + idx[q]] = w[q];
    }
    P.chapman.assign(size_t(P.L) * P.L, 0.0);
    if (geotype == 0) {  // :57-61
        for (int p = 0; p < P.L; ++p)
            for (int q = 0; q <= p; ++q) P.chapman[size_t(p) * P.L + q] = 1 / cos_sza;
    } else if (!chapman_straight_line_ref()) {
        chapman_raytraced(alt, P.floor_h, P.ceil_h, cos_sza, earth_radius, P.chapman);
    } else {  // :69-119
        double sinthetasq = 1 - cos_sza * cos_sza;
        for (int p = 0; p < P.L; ++p) {
            double rp = earth_radius + P.floor_h[p];
            for (int q = 0; q <= p; ++q) {
                double rfloor = earth_radius + P.floor_h[q];
                double rceil = earth_radius + P.ceil_h[q];
                P.chapman[size_t(p) * P.L + q] =
                    (std::sqrt(rceil * rceil - rp * rp * sinthetasq) - std::sqrt(rfloor * rfloor - rp * rp * sinthetasq)) /
                    (rceil - rfloor);
            }
        }
    }
    return P;
}

// ---------------------------------------------------------------------------------------------------
//  Banded LU with partial pivoting in LAPACK general-band storage (kl = ku), restating dgbtf2 / dgbtrs
//  ('N' and 'T').  ab is ldab x n column-major, ldab = 2*kl+ku+1, A(i,j) at ab[kl+ku+i-j + j*ldab]
//  (== la::BVPMatrix::operator(), sktran_do_linalg.h:35-40).
// ---------------------------------------------------------------------------------------------------
struct BandLU {
    int n = 0, kl = 0, ku = 0, ldab = 0;
    std::vector<double> ab;
    std::vector<int> ipiv;
    void init(int n_, int kl_, int ku_) {
        n = n_; kl = kl_; ku = ku_; ldab = 2 * kl + ku + 1;
        ab.assign(size_t(ldab) * n, 0.0);
        ipiv.assign(n, 0);
    }
    double& at(int i, int j) { return ab[size_t(kl + ku + i - j) + size_t(j) * ldab]; }
    // returns LAPACK-style info (0 ok, j+1 if U(j,j) == 0)
    int factor() {
        int kv = ku + kl, info = 0, ju = 0;
        for (int j = 0; j < n; ++j) {
            int km = std::min(kl, n - 1 - j);
            double* col = &ab[size_t(j) * ldab + kv];  // diagonal element of column j
            int jp = 0;
            double amax = std::abs(col[0]);
            for (int i = 1; i <= km; ++i)
                if (std::abs(col[i]) > amax) { amax = std::abs(col[i]); jp = i; }
            ipiv[j] = jp + j;
            if (col[jp] != 0.0) {
                ju = std::max(ju, std::min(j + ku + jp, n - 1));
                if (jp != 0)
                    for (int c = j; c <= ju; ++c) std::swap(at(j + jp, c), at(j, c));
                if (km > 0) {
                    double r = 1.0 / col[0];
                    for (int i = 1; i <= km; ++i) col[i] *= r;
                    for (int c = j + 1; c <= ju; ++c) {
                        double t = at(j, c);
                        if (t != 0.0) {
                            double* cc = &ab[size_t(c) * ldab + kv + j - c];  // element (j, c)
                            for (int i = 1; i <= km; ++i) cc[i] -= col[i] * t;
                        }
                    }
                }
            } else if (info == 0) {
                info = j + 1;
            }
        }
        return info;
    }
    void solve(double* b) const {
        int kv = ku + kl;
        for (int j = 0; j < n - 1; ++j) {
            int lm = std::min(kl, n - 1 - j);
            int l = ipiv[j];
            if (l != j) std::swap(b[l], b[j]);
            const double* col = &ab[size_t(j) * ldab + kv];
            double bj = b[j];
            for (int i = 1; i <= lm; ++i) b[j + i] -= bj * col[i];
        }
        for (int j = n - 1; j >= 0; --j) {
            const double* col = &ab[size_t(j) * ldab + kv];
            b[j] /= col[0];
            double bj = b[j];
            int lo = std::max(0, j - kv);
            for (int i = lo; i < j; ++i) b[i] -= bj * col[i - j];
        }
    }
    void solve_transposed(double* b) const {
        int kv = ku + kl;
        for (int j = 0; j < n; ++j) {
            const double* col = &ab[size_t(j) * ldab + kv];
            double t = b[j];
            int lo = std::max(0, j - kv);
            for (int i = lo; i < j; ++i) t -= col[i - j] * b[i];
            b[j] = t / col[0];
        }
        for (int j = n - 2; j >= 0; --j) {
            int lm = std::min(kl, n - 1 - j);
            const double* col = &ab[size_t(j) * ldab + kv];
            double t = b[j];
            for (int i = 1; i <= lm; ++i) t -= col[i] * b[j + i];
            b[j] = t;
            int l = ipiv[j];
            if (l != j) std::swap(b[l], b[j]);
        }
    }
};

// dense LU with partial pivoting (bordered eigen-derivative system, sktran_do_rte.cpp:254-256)
struct DenseLU {
    int n = 0;
    std::vector<double> a;
    std::vector<int> piv;
    bool factor(int n_, const std::vector<double>& m) {  // m row-major
        n = n_; a = m; piv.resize(n);
        for (int k = 0; k < n; ++k) {
            int p = k;
            for (int i = k + 1; i < n; ++i)
                if (std::abs(a[i * n + k]) > std::abs(a[p * n + k])) p = i;
            piv[k] = p;
            if (a[p * n + k] == 0.0) return false;
            if (p != k)
                for (int c = 0; c < n; ++c) std::swap(a[p * n + c], a[k * n + c]);
            for (int i = k + 1; i < n; ++i) {
                a[i * n + k] /= a[k * n + k];
                double f = a[i * n + k];
                for (int c = k + 1; c < n; ++c) a[i * n + c] -= f * a[k * n + c];
            }
        }
        return true;
    }
    void solve(double* b) const {
        // full rows were swapped in factor() (LAPACK getrf convention): permute b first, then L, then U
        for (int k = 0; k < n; ++k)
            if (piv[k] != k) std::swap(b[k], b[piv[k]]);
        for (int k = 0; k < n; ++k)
            for (int i = k + 1; i < n; ++i) b[i] -= a[i * n + k] * b[k];
        for (int k = n - 1; k >= 0; --k) {
            for (int c = k + 1; c < n; ++c) b[k] -= a[k * n + c] * b[c];
            b[k] /= a[k * n + k];
        }
    }
};

// ---------------------------------------------------------------------------------------------------
//  Inputs for one wavelength
// ---------------------------------------------------------------------------------------------------
struct WavelInputs {
    const double* ext;   // [nloc] total extinction
    const double* ssa;   // [nloc]
    const double* leg;   // [nleg][nloc] (leg fastest: leg[l + nleg*q])
    int nleg;
    const double* f;     // [nloc] delta-M fraction or nullptr
    const double* d_f = nullptr;  // [nloc * g + q] derivative of f per scattering group, nullptr: no scaling applied
    double solar;        // solar irradiance
    double albedo;       // Lambertian albedo
    // scattering derivative groups: d_leg[l + nleg*(q + nloc*g)] or nullptr
    const double* d_leg;
    int ngroups;
    bool include_ss;     // single_scatter_source == discrete_ordinates
    int num_azimuth;     // nstr or num_do_forced_azimuth
    // surface BRDF: 0 Lambertian (albedo above), 1 snow (Kokhanovsky), 2 MODIS kernels; args [nargs] of this wavelength
    int brdf_kind = 0;
    const double* brdf_args = nullptr;
    // thermal emission (config.emission_source == discrete_ordinates): emission_source at the grid points [nloc] of
    // this wavelength (nullptr: no atmospheric emission) and the surface emission (Surface::emission, 0 by default;
    // the reference adds it whatever the emission source is, sktran_do_rte.h:229-235)
    const double* emission = nullptr;
    double surface_emission = 0.0;
};

// ---------------------------------------------------------------------------------------------------
//  Surface BRDF models (cpp/include/sasktran2/atmosphere/surface.h:112-362) and their azimuthal Fourier
//  expansion (SurfaceStorage::compute_expansion, cpp/include/sktran_disco/sktran_do_surface.h:49-91)
// ---------------------------------------------------------------------------------------------------
inline double brdf_value(int kind, const double* args, double mu_in, double mu_out, double phi_diff) {
    if (kind == 1) {  // SnowKokhanovsky, surface.h:151-199
        auto K0 = [](double mu) { return (3.0 / 7.0) * (1.0 + 2.0 * mu); };
        double mus = mu_in, muv = mu_out;
        double ss = std::sqrt(1 - mus * mus), sv = std::sqrt(1 - muv * muv);
        double cost = std::max(-1.0, std::min(1.0, -mus * muv + ss * sv * std::cos(phi_diff)));
        double theta = std::acos(cost) * 180.0 / PI;
        double alpha = std::sqrt(4 * PI * args[0]);
        double pth = 11.1 * std::exp(-0.087 * theta) + 1.1 * std::exp(-0.014 * theta);
        double r0 = (1.247 + 1.186 * (mus + muv) + 5.157 * mus * muv + pth) / (4.0 * (mus + muv));
        return r0 * std::exp(-alpha * K0(mus) * K0(muv) / r0) / PI;
    }
    if (kind == 2) {  // MODIS (Ross-thick / Li-sparse-R), surface.h:246-294
        double csza = mu_in, cvza = mu_out;
        double ssza = std::sqrt(1 - csza * csza), svza = std::sqrt(1 - cvza * cvza);
        double tsza = ssza / csza, tvza = svza / cvza;
        double craa = -std::cos(phi_diff), sraa = std::sin(phi_diff);
        double csa = std::max(-1.0, std::min(1.0, csza * cvza + ssza * svza * craa));
        double sa = std::acos(csa), ssa = std::sin(sa);
        double k_vol = ((0.5 * PI - sa) * csa + ssa) / (csza + cvza) - 0.25 * PI;
        double d2 = tsza * tsza + tvza * tvza - 2 * tsza * tvza * craa;
        double ct = std::max(-1.0, std::min(1.0, 2 * std::sqrt(d2 + tsza * tsza * tvza * tvza * sraa * sraa) * csza * cvza / (csza + cvza)));
        double t = std::acos(ct), st = std::sin(t);
        double o = (t - st * ct) * (csza + cvza) / (PI * csza * cvza);
        double k_geo = o - (csza + cvza - 0.5 * (1 + csa)) / (csza * cvza);
        return (args[0] + args[1] * k_vol + args[2] * k_geo) / PI;
    }
    return args[0] / PI;  // Lambertian, surface.h:112-122
}
inline void gauss_legendre(int n, std::vector<double>& x, std::vector<double>& w);
inline double compute_expansion(int m, int kind, const double* args, double mu_out, double mu_in) {
    if (kind == 0) return m == 0 ? args[0] : 0.0;  // max_azimuthal_order() == 1: brdf * pi for m = 0, nothing above
    static std::vector<double> qx, qw;  // 512-point Gauss-Legendre rule (getQuadratureAbscissae / Weights (512))
    if (qx.empty()) {
        std::vector<double> x, w;
        gauss_legendre(512, x, w);
#pragma omp critical(oracle_brdf_quadrature)
        if (qx.empty()) {
            qw = w;
            qx = x;
        }
    }
    double result = 0;
    for (size_t i = 0; i < qx.size() / 2; ++i) {
        const double a[4] = {0.5 * qx[i] + 0.5, -0.5 * qx[i] + 0.5, 0.5 * qx[i] - 0.5, -0.5 * qx[i] - 0.5};
        const double w = 0.5 * qw[i];
        for (int k = 0; k < 4; ++k) result += w * brdf_value(kind, args, mu_in, mu_out, PI * a[k]) * std::cos(m * PI * a[k]);
    }
    return result * 0.5 * PI * (2.0 - (m == 0 ? 1.0 : 0.0));
}
// Expansion coefficients of one azimuth order for the stream / solar / line-of-sight pairs the solver needs
// (Surface::calculate, sktran_do_surface.h:153-217)
struct SurfaceExpansion {
    bool general = false;
    std::vector<double> ss, sun, los, lsun;  // [N*N] (out i, in q), [N], [nlos*N], [nlos]
};

template <class T>
struct LayerSolution {
    std::vector<T> k;              // [N] eigval
    std::vector<T> Wp, Wm;         // [N*N] column-major W(i + N*j): stream i, solution j
    std::vector<T> Ap, Am;         // [N]
    std::vector<T> Ath;            // [N] thermal Green's coefficient (A+ = A- for the scalar thermal source), empty: none
    std::vector<T> Gpt, Gpb, Gmt, Gmb;  // [N]
    std::vector<T> Lc, Mc;         // [N] BVP coefficients
};

template <class T>
struct Layers {
    std::vector<T> od, ssa;           // [L]
    std::vector<std::vector<T>> beta; // [L][nstr]
    std::vector<T> secant;            // [L]
    std::vector<T> trans;             // [L+1] beam transmittance at boundaries (incl. F0)
    std::vector<double> tot_ext, scat_ext, ssa_value;  // per-layer scalars used by the WF mapping
    std::vector<double> b0, b1;       // [L] thermal source b0 exp(-b1 x) of the layer, empty: no thermal emission
    double surface_emission = 0.0;
};

// Lane bookkeeping for T = Dual: lane index of (layer p, kind) in the reference's sorted order
// (sktran_do_types.h:233-251): per layer [scat g=0..G-1][od][ssa]; last layer + [albedo].
struct Lanes {
    int L = 0, G = 0;
    bool local = false;  // layer-local lanes of ReverseSolver: the same G + 2 lane numbers for every layer
    int per_layer() const { return G + 2; }
    int total() const { return L * (G + 2) + 1; }
    int scat(int p, int g) const { return local ? g : p * (G + 2) + g; }
    int od(int p) const { return local ? G : p * (G + 2) + G; }
    int ssa(int p) const { return local ? G + 1 : p * (G + 2) + G + 1; }
    int albedo() const { return L * (G + 2); }
};

inline void seed(double&, int) {}
inline void seed(Dual& x, int lane) { x.d[lane] = 1.0; }
template <int K> inline void seed(FDual<K>& x, int lane) { x.d[lane] = 1.0; }
inline void seed_dir(double&, int, double) {}
inline void seed_dir(Dual& x, int lane, double v) { x.d[lane] = v; }
template <int K> inline void seed_dir(FDual<K>& x, int lane, double v) { x.d[lane] = v; }

template <class T>
struct Solver {
    const Plan& P;
    dgeev_fn dgeev;
    Lanes lanes;
    const SurfaceExpansion* surf = nullptr;  // non-Lambertian surface of the order being solved (values only)
    Solver(const Plan& p, dgeev_fn f) : P(p), dgeev(f) {}

    // ---- layer optics: OpticalLayerArray ctor, sktran_do_layerarray.cpp:332-477 + OpticalLayer ctor
    //      sktran_do_opticallayer.cpp:17-40; scattering-lane directions :773-800 (with the reference's
    //      behaviour that the direction comes from the LAST contributing grid point).
    void layer_optics(const WavelInputs& in, Layers<T>& Ly) const {
        const int L = P.L, nloc = P.nloc, nstr = P.nstr;
        Ly.od.assign(L, T(0.0));
        Ly.ssa.assign(L, T(0.0));
        Ly.beta.assign(L, std::vector<T>(nstr, T(0.0)));
        Ly.tot_ext.assign(L, 0.0);
        Ly.scat_ext.assign(L, 0.0);
        Ly.ssa_value.assign(L, 0.0);
        double ceiling_depth = 0, floor_depth = 0;
        Ly.b0.clear();
        Ly.b1.clear();
        Ly.surface_emission = in.surface_emission;
        if (in.emission) {
            Ly.b0.assign(L, 0.0);
            Ly.b1.assign(L, 0.0);
        }
        for (int p = 0; p < L; ++p) {
            double dh = P.ceil_h[p] - P.floor_h[p];
            double od = 0, ssa = 0;
            std::vector<double> leg(nstr, 0.0);
            int last_q = -1;
            for (int q = 0; q < nloc; ++q) {
                double w = P.W[size_t(p) * nloc + q];
                if (w > 0) {
                    double kext = in.ext[q];
                    double kscat = in.ssa[q] * kext;
                    od += kext * w;
                    ssa += kscat * w;
                    double f = in.f ? in.f[q] : 0.0;
                    for (int k = 0; k < nstr; ++k) {
                        double ph = (k < in.nleg) ? in.leg[k + size_t(in.nleg) * q] : 0.0;
                        leg[k] += w * kscat * (ph - (2 * k + 1) * f / (1 - f));
                    }
                    last_q = q;
                }
            }
            if (ssa > 0) {
                for (int k = 0; k < nstr; ++k) leg[k] /= ssa;
            } else {
                leg[0] = 0;
            }
            ssa /= od;
            od *= dh;
            floor_depth += od;
            if (in.emission) {
                // thermal source S(x) = b0 exp(-b1 x): emission at the highest / lowest contributing grid point
                // (sktran_do_layerarray.cpp:341-370, 459-470)
                int min_q = -1, max_q = -1;
                for (int q = 0; q < nloc; ++q)
                    if (P.W[size_t(p) * nloc + q] > 0) {
                        if (min_q < 0) min_q = q;
                        max_q = q;
                    }
                double b0_top = max_q >= 0 ? in.emission[max_q] : 0.0;
                double b0_bot = (min_q >= 0 && min_q != max_q) ? in.emission[min_q] : b0_top;
                double b1 = 0.0;
                if (od > 1e-10 && b0_top > 1e-30 && b0_bot > 1e-30 &&
                    std::abs(b0_top - b0_bot) > 1e-15 * std::max(b0_top, b0_bot))
                    b1 = std::log(b0_top / b0_bot) / od;
                Ly.b0[p] = b0_top;
                Ly.b1[p] = b1;
            }
            double total_ext = od / dh;
            double scat_ext = total_ext * ssa;
            scat_ext = std::max(scat_ext, total_ext * SSA_DITHER);
            double m_ssa = scat_ext / total_ext;
            if (1 - m_ssa < SSA_DITHER) m_ssa = 1 - SSA_DITHER;
            double thickness = floor_depth - ceiling_depth;  // M_OPTICAL_THICKNESS
            ceiling_depth = floor_depth;

            Ly.tot_ext[p] = total_ext;
            Ly.scat_ext[p] = scat_ext;
            Ly.ssa_value[p] = m_ssa;
            Ly.od[p] = T(thickness);
            Ly.ssa[p] = T(m_ssa);
            seed(Ly.od[p], lanes.od(p));
            seed(Ly.ssa[p], lanes.ssa(p));
            for (int k = 0; k < nstr; ++k) Ly.beta[p][k] = T(leg[k]);
            if (in.d_leg && last_q >= 0) {
                double f = in.f ? in.f[last_q] : 0.0;
                for (int g = 0; g < in.ngroups; ++g)
                    for (int l = 0; l < nstr; ++l) {
                        double ph = (l < in.nleg) ? in.leg[l + size_t(in.nleg) * last_q] : 0.0;
                        double dl = (l < in.nleg) ? in.d_leg[l + size_t(in.nleg) * (last_q + size_t(nloc) * g)] : 0.0;
                        double dir = dl + (ph - (2 * l + 1) * f / (1 - f) - leg[l]);
                        // applied_f_order > 0 (sktran_do_layerarray.cpp:792-800)
                        if (in.d_f) dir += -(2 * l + 1) / (1 - f) / (1 - f) * in.d_f[last_q + size_t(nloc) * g];
                        seed_dir(Ly.beta[p][l], lanes.scat(p, g), dir);
                    }
            }
        }
        // configureTransmission, sktran_do_layerarray.cpp:891-979
        std::vector<T> slant(L + 1, T(0.0));
        Ly.secant.assign(L, T(0.0));
        Ly.trans.assign(L + 1, T(0.0));
        for (int p = 0; p < L; ++p) {
            T acc(0.0);
            for (int q = 0; q < L; ++q) {
                double c = P.chapman[size_t(p) * L + q];
                if (c != 0.0) acc += Ly.od[q] * T(c);
            }
            slant[p + 1] = acc;
            Ly.secant[p] = (slant[p + 1] - slant[p]) / Ly.od[p];
        }
        Ly.trans[0] = T(in.solar);
        for (int p = 0; p < L; ++p) Ly.trans[p + 1] = exp(-slant[p + 1]) * T(in.solar);
    }

    // ---- homogeneous solution: lp_triple_product (sktran_do_lpproduct.h:165-263) + solveHomogeneous
    //      (sktran_do_rte.cpp:383-553) + linearizeHomogeneous (:198-298)
    void homogeneous(int m, const T& ssa, const std::vector<T>& beta, LayerSolution<T>& S) const {
        const int N = P.N, nstr = P.nstr;
        std::vector<T> Sp(N * N), Sm(N * N);  // row-major (i,j)
        for (int i = 0; i < N; ++i)
            for (int j = 0; j < N; ++j) {
                T sp(0.0), eta(0.0);
                for (int l = m; l < nstr; ++l) {
                    double pp = P.LPmu(m, i, l) * P.LPmu(m, j, l);
                    sp += beta[l] * T(pp);
                    eta += beta[l] * T(((l - m) % 2 != 0) ? -pp : pp);
                }
                double q = -0.5 * P.wt[j] / P.mu[i];
                sp = sp * T(q) * ssa;
                eta = eta * T(q) * ssa;
                if (i == j) sp += T(1 / P.mu[i]);
                Sp[i * N + j] = sp + eta;
                Sm[i * N + j] = sp - eta;
            }
        std::vector<T> E(N * N);
        for (int i = 0; i < N; ++i)
            for (int j = 0; j < N; ++j) {
                T acc(0.0);
                for (int q = 0; q < N; ++q) acc += Sm[i * N + q] * Sp[q * N + j];
                E[i * N + j] = acc;
            }
        // eigen-decomposition of the value matrix (LAPACK dgeev, column-major input)
        std::vector<double> a(N * N), wr(N), wi(N), vr(N * N), work(8 * N + 16);
        for (int i = 0; i < N; ++i)
            for (int j = 0; j < N; ++j) a[i + j * N] = val(E[i * N + j]);
        int n = N, one = 1, lwork = (int)work.size(), info = 0;
        dgeev("N", "V", &n, a.data(), &n, wr.data(), wi.data(), nullptr, &one, vr.data(), &n, work.data(), &lwork, &info);
        if (info != 0) throw std::runtime_error("dgeev failed");
        std::vector<T> X(N * N);  // column-major X(i + N*j)
        std::vector<T> ksq(N);
        for (int j = 0; j < N; ++j) {
            double nrm = 0;
            for (int i = 0; i < N; ++i) nrm += vr[i + j * N] * vr[i + j * N];
            nrm = std::sqrt(nrm);
            for (int i = 0; i < N; ++i) X[i + j * N] = T(vr[i + j * N] / nrm);
            if (wr[j] <= 0) throw std::runtime_error("imaginary homogeneous solution");
            ksq[j] = T(wr[j]);
        }
        S.k.assign(N, T(0.0));
        for (int j = 0; j < N; ++j) S.k[j] = T(std::sqrt(std::abs(wr[j])));
        eig_derivatives(E, X, S.k);
        S.Wp.assign(N * N, T(0.0));
        S.Wm.assign(N * N, T(0.0));
        for (int j = 0; j < N; ++j)
            for (int i = 0; i < N; ++i) {
                T xm(0.0);
                for (int q = 0; q < N; ++q) xm += Sp[i * N + q] * X[q + j * N];
                S.Wp[i + j * N] = T(0.5) * (X[i + j * N] + xm / S.k[j]);
                S.Wm[i + j * N] = T(0.5) * (X[i + j * N] - xm / S.k[j]);
            }
    }
    void eig_derivatives(const std::vector<double>&, std::vector<double>&, std::vector<double>&) const {}
    template <class D>
    void eig_derivatives(const std::vector<D>& E, std::vector<D>& X, std::vector<D>& k) const {
        const int N = P.N, nd = ndual(E[0]);
        if (nd == 0) return;
        std::vector<double> lhs((N + 1) * (N + 1)), rhs(N + 1);
        DenseLU lu;
        for (int j = 0; j < N; ++j) {
            // which lanes have a non-zero RHS?
            bool any = false;
            std::vector<char> active(nd, 0);
            for (int d = 0; d < nd; ++d) {
                for (int i = 0; i < N && !active[d]; ++i) {
                    double s = 0;
                    for (int q = 0; q < N; ++q) s += E[i * N + q].d[d] * X[q + j * N].v;
                    if (s != 0.0) active[d] = 1;
                }
                any = any || active[d];
            }
            if (!any) continue;  // exactly-zero RHS shortcut, sktran_do_rte.cpp:249-253
            double kj = k[j].v;
            for (int r = 0; r < N; ++r) {
                for (int c = 0; c < N; ++c) lhs[r * (N + 1) + c] = E[r * N + c].v - (r == c ? kj * kj : 0.0);
                lhs[r * (N + 1) + N] = -2 * kj * X[r + j * N].v;
            }
            for (int c = 0; c < N; ++c) lhs[N * (N + 1) + c] = X[c + j * N].v;
            lhs[N * (N + 1) + N] = 0;
            if (!lu.factor(N + 1, lhs)) throw std::runtime_error("singular bordered eigen system");
            for (int d = 0; d < nd; ++d) {
                if (!active[d]) continue;
                for (int i = 0; i < N; ++i) {
                    double s = 0;
                    for (int q = 0; q < N; ++q) s += E[i * N + q].d[d] * X[q + j * N].v;
                    rhs[i] = -s;
                }
                rhs[N] = 0;
                lu.solve(rhs.data());
                for (int i = 0; i < N; ++i) X[i + j * N].d[d] = rhs[i];
                k[j].d[d] = rhs[N];
            }
        }
    }

    // ---- particular solution: assignParticularQ (sktran_do_rte.cpp:556-580), single_scat_st
    //      (sktran_do_lpproduct.h:339-384), solveParticularGreen (:903-1332)
    void particular(int m, const T& ssa, const std::vector<T>& beta, const T& od, const T& secant, const T& trans_top,
                    LayerSolution<T>& S) const {
        const int N = P.N, nstr = P.nstr;
        std::vector<T> Qp(N, T(0.0)), Qm(N, T(0.0));
        bool allzero = true;
        for (int i = 0; i < N; ++i) {
            T sp(0.0), sm(0.0);
            for (int l = m; l < nstr; ++l) {
                double pp = P.LPmu(m, i, l) * P.LPcsz(m, l);
                sp += beta[l] * T(pp);
                sm += beta[l] * T(((l - m) % 2 != 0) ? -pp : pp);
            }
            double factor = (2.0 - (m == 0 ? 1.0 : 0.0)) * (1.0 / (4.0 * PI)) * P.wt[i];
            Qp[i] = sp * T(factor) * ssa;
            Qm[i] = sm * T(factor) * ssa;
            if (!is_zero(Qp[i]) || !is_zero(Qm[i])) allzero = false;
        }
        S.Ap.assign(N, T(0.0)); S.Am.assign(N, T(0.0));
        S.Gpt.assign(N, T(0.0)); S.Gpb.assign(N, T(0.0)); S.Gmt.assign(N, T(0.0)); S.Gmb.assign(N, T(0.0));
        if (allzero) return;  // :956-959
        T exp_sec = exp(-od * secant);
        for (int j = 0; j < N; ++j) {
            T norm(0.0), ap(0.0), am(0.0);
            for (int i = 0; i < N; ++i) {
                const T& wp = S.Wp[i + j * N];
                const T& wm = S.Wm[i + j * N];
                norm += T(P.wt[i] * P.mu[i]) * (wp * wp - wm * wm);
                ap += Qp[i] * wp + Qm[i] * wm;
                am += Qm[i] * wp + Qp[i] * wm;
            }
            ap = ap / norm;
            am = am / norm;
            S.Ap[j] = ap;
            S.Am[j] = am;
            T exp_k = exp(-od * S.k[j]);
            T Cp, Cm;
            if (std::abs(val(secant) - val(S.k[j])) > GREENS_EPS)
                Cp = stable_multipliers_ref() ? trans_top * od * psi(od, S.k[j], secant, exp_k, exp_sec)
                                              : trans_top * (exp_k - exp_sec) / (secant - S.k[j]);
            else
                Cp = trans_top * exp_k * od * (T(1.0) - od / T(2.0) * (secant - S.k[j]));
            if (std::abs(val(secant) + val(S.k[j])) > GREENS_EPS)
                Cm = trans_top * (T(1.0) - exp_sec * exp_k) / (secant + S.k[j]);
            else
                Cm = trans_top * od * (T(1.0) - od / T(2.0) * (secant + S.k[j]));
            for (int i = 0; i < N; ++i) {
                S.Gpt[i] += am * Cm * S.Wm[i + j * N];
                S.Gmt[i] += am * Cm * S.Wp[i + j * N];
                S.Gpb[i] += ap * Cp * S.Wp[i + j * N];
                S.Gmb[i] += ap * Cp * S.Wm[i + j * N];
            }
        }
    }
    // ---- thermal particular solution, order 0 only: solveParticularGreenThermal (sktran_do_rte.cpp:1335-1617), added on
    //      top of the solar G+-.  Values only (the reference's derivative lanes of b0 / b1 are not restated).
    void particular_thermal(const T& ssa, const T& od, double b0, double b1, LayerSolution<T>& S) const {
        const int N = P.N;
        S.Ath.assign(N, T(0.0));
        const double tau = val(od), e_b1 = std::exp(-tau * b1);
        for (int j = 0; j < N; ++j) {
            T norm(0.0), a(0.0);
            for (int i = 0; i < N; ++i) {
                const T& wp = S.Wp[i + j * N];
                const T& wm = S.Wm[i + j * N];
                norm += T(P.wt[i] * P.mu[i]) * (wp * wp - wm * wm);
                a += T(P.wt[i]) * (T(1.0) - ssa) * (wp + wm);
            }
            a = a / norm;
            S.Ath[j] = a;
            const double k = val(S.k[j]), e_k = std::exp(-tau * k);
            double Cp, Cm;
            if (std::abs(b1 - k) > GREENS_EPS)
                Cp = b0 * (e_k - e_b1) / (b1 - k);
            else
                Cp = b0 * e_k * tau * (1 - tau / 2 * (b1 - k));
            if (std::abs(b1 + k) > GREENS_EPS)
                Cm = b0 * (1 - e_b1 * e_k) / (b1 + k);
            else
                Cm = b0 * tau * (1 - tau / 2 * (b1 + k));
            for (int i = 0; i < N; ++i) {
                S.Gpt[i] += a * T(Cm) * S.Wm[i + j * N];
                S.Gmt[i] += a * T(Cm) * S.Wp[i + j * N];
                S.Gpb[i] += a * T(Cp) * S.Wp[i + j * N];
                S.Gmb[i] += a * T(Cp) * S.Wm[i + j * N];
            }
        }
    }
    static bool is_zero(double x) { return x == 0.0; }
    template <class D>
    static bool is_zero(const D& x) {
        if (x.v != 0.0) return false;
        for (int i = 0; i < ndual(x); ++i)
            if (x.d[i] != 0.0) return false;
        return true;
    }

    // ---- BVP: solveBVP + bvp*Condition (sktran_do_rte.cpp:1621-1790, 1898-2294), v_plus/v_minus/u_minus/
    //      ground_direct_sun (sktran_do_rte.h:116-345)
    void bvp(int m, const Layers<T>& Ly, const T& albedo, std::vector<LayerSolution<T>>& sol) const {
        const int N = P.N, L = P.L, n = 2 * N * L, kl = 3 * N - 1;
        // assemble entries as T, then factor the value matrix
        struct Entry { int r, c; T a; };
        std::vector<Entry> ent;
        ent.reserve(size_t(L) * 8 * N * N);
        std::vector<T> b(n, T(0.0));
        std::vector<std::vector<T>> theta(L, std::vector<T>(N));
        for (int p = 0; p < L; ++p)
            for (int j = 0; j < N; ++j) theta[p][j] = exp(-S_abs(sol[p].k[j]) * Ly.od[p]);
        // TOA (:1898-1942, :2131-2172)
        for (int i = 0; i < N; ++i) {
            for (int j = 0; j < N; ++j) {
                ent.push_back({i, j, sol[0].Wp[i + j * N]});
                ent.push_back({i, j + N, sol[0].Wm[i + j * N] * theta[0][j]});
            }
            b[i] = -sol[0].Gpt[i];
        }
        // continuity (:1945-2072, :2175-2267)
        for (int bd = 1; bd < L; ++bd) {
            int r0 = N + (bd - 1) * 2 * N, c0 = (bd - 1) * 2 * N;
            const auto& U = sol[bd - 1];
            const auto& Lo = sol[bd];
            for (int i = 0; i < N; ++i) {
                for (int j = 0; j < N; ++j) {
                    ent.push_back({r0 + i + N, c0 + j, U.Wp[i + j * N] * theta[bd - 1][j]});
                    ent.push_back({r0 + i + N, c0 + 2 * N + j, -Lo.Wp[i + j * N]});
                    ent.push_back({r0 + i, c0 + j, U.Wm[i + j * N] * theta[bd - 1][j]});
                    ent.push_back({r0 + i, c0 + 2 * N + j, -Lo.Wm[i + j * N]});
                    ent.push_back({r0 + i + N, c0 + N + j, U.Wm[i + j * N]});
                    ent.push_back({r0 + i + N, c0 + 3 * N + j, -(Lo.Wm[i + j * N] * theta[bd][j])});
                    ent.push_back({r0 + i, c0 + N + j, U.Wp[i + j * N]});
                    ent.push_back({r0 + i, c0 + 3 * N + j, -(Lo.Wp[i + j * N] * theta[bd][j])});
                }
                b[r0 + i] = -U.Gmb[i] + Lo.Gmt[i];
                b[r0 + i + N] = -U.Gpb[i] + Lo.Gpt[i];
            }
        }
        // ground (:2075-2128, :2270-2294).  Lambertian: rho = albedo for every pair, only m = 0; a general BRDF
        // reflects every order with rho_m(mu_i, mu_q) (sktran_do_rte.h:116-345)
        {
            int r0 = N + (L - 1) * 2 * N, c0 = n - 2 * N;
            const auto& B = sol[L - 1];
            const bool gen = surf && surf->general;
            bool refl = gen || (m == 0);
            double kd = (m == 0) ? 2.0 : 1.0;
            auto rho = [&](int i, int q) { return gen ? T(surf->ss[i * N + q]) : albedo; };
            for (int i = 0; i < N; ++i) {
                for (int j = 0; j < N; ++j) {
                    T vm = B.Wm[i + j * N], vp = B.Wp[i + j * N];
                    if (refl)
                        for (int q = 0; q < N; ++q) {
                            vm -= T(kd) * rho(i, q) * T(P.wt[q] * P.mu[q]) * B.Wp[q + j * N];
                            vp -= T(kd) * rho(i, q) * T(P.wt[q] * P.mu[q]) * B.Wm[q + j * N];
                        }
                    ent.push_back({r0 + i, c0 + j, vm * theta[L - 1][j]});
                    ent.push_back({r0 + i, c0 + N + j, vp});
                }
                T gds(0.0);
                if (m == 0) gds = T(Ly.surface_emission);   // ground_direct_sun: thermal source, sktran_do_rte.h:229-235
                if (refl) gds += T(P.csz) * (gen ? T(surf->sun[i]) : albedo) / T(PI) * Ly.trans[L];
                T um = B.Gmb[i];
                if (refl)
                    for (int q = 0; q < N; ++q) um -= T(kd) * rho(i, q) * T(P.wt[q] * P.mu[q]) * B.Gpb[q];
                b[r0 + i] = gds - um;
            }
        }
        BandLU lu;
        lu.init(n, kl, kl);
        for (const auto& e : ent) lu.at(e.r, e.c) = val(e.a);
        int info = lu.factor();
        if (info != 0) throw std::runtime_error("BVP matrix singular");
        std::vector<double> x(n);
        for (int i = 0; i < n; ++i) x[i] = val(b[i]);
        lu.solve(x.data());
        std::vector<T> xs(n);
        for (int i = 0; i < n; ++i) xs[i] = T(x[i]);
        bvp_derivs(ent, b, lu, x, xs);
        for (int p = 0; p < L; ++p) {
            sol[p].Lc.assign(N, T(0.0));
            sol[p].Mc.assign(N, T(0.0));
            for (int j = 0; j < N; ++j) {
                sol[p].Lc[j] = xs[p * 2 * N + j];
                sol[p].Mc[j] = xs[p * 2 * N + N + j];
            }
        }
    }
    static double S_abs(double x) { return std::abs(x); }
    template <class D>
    static D S_abs(const D& x) { return x.v >= 0 ? x : -x; }
    template <class E>
    void bvp_derivs(const std::vector<E>&, const std::vector<double>&, const BandLU&, const std::vector<double>&,
                    std::vector<double>&) const {}
    // x' = A^-1 (b' - A' x), sktran_do_rte.cpp:1730-1789
    template <class E>
    void bvp_derivs(const std::vector<E>& ent, const std::vector<Dual>& b, const BandLU& lu, const std::vector<double>& x,
                    std::vector<Dual>& xs) const {
        const int nd = nd_ref(), n = (int)x.size();
        std::vector<double> rhs(n);
        for (int d = 0; d < nd; ++d) {
            bool any = false;
            for (int i = 0; i < n; ++i) {
                rhs[i] = b[i].d[d];
                any = any || rhs[i] != 0.0;
            }
            for (const auto& e : ent) {
                double da = e.a.d[d];
                if (da != 0.0) {
                    rhs[e.r] -= da * x[e.c];
                    any = true;
                }
            }
            if (!any) continue;
            lu.solve(rhs.data());
            for (int i = 0; i < n; ++i) xs[i].d[d] = rhs[i];
        }
    }

    // ---- post-processing for one (m, LOS): computeReflectedIntensities (sktran_do_layerarray.cpp:5-288),
    //      integrate_source / h_plus / h_minus / E (sktran_do_opticallayer.cpp:94-555, 785-938), upward
    //      recursion (do_source_planeparallel.cpp:69-146).  Observer above the top of the atmosphere.
    // Ground-leaving radiance of order m toward any line of sight (Lambertian: m = 0 only).  gL / gM (optional):
    // its partial derivatives w.r.t. the bottom layer's L_q, M_q at fixed layer quantities.
    T ground_term(int m, const T& trans_bottom, const T& od_last, const T& albedo_in, const LayerSolution<T>& B,
                  bool include_ss, double* gL = nullptr, double* gM = nullptr, int los = -1) const {
        const int N = P.N;
        if (gL)
            for (int q = 0; q < N; ++q) gL[q] = gM[q] = 0.0;
        const bool gen = surf && surf->general && los >= 0;
        if (m != 0 && !gen) return T(0.0);  // Lambertian: max_azimuthal_order == 1
        T diffuse(0.0);
        for (int i = 0; i < N; ++i) {
            const T albedo = gen ? T(surf->los[los * N + i]) : albedo_in;   // rho_m(mu_los, mu_i)
            T sc = B.Gpb[i];
            double factor = (m == 0 ? 2.0 : 1.0) * P.mu[i] * P.wt[i];
            for (int q = 0; q < N; ++q) {
                T th = exp(-S_abs(B.k[q]) * od_last);
                sc += B.Lc[q] * B.Wp[i + q * N] * th;
                sc += B.Mc[q] * B.Wm[i + q * N];
                if (gL) {
                    gL[q] += factor * val(albedo) * val(B.Wp[i + q * N]) * val(th);
                    gM[q] += factor * val(albedo) * val(B.Wm[i + q * N]);
                }
            }
            diffuse += T(factor) * sc * albedo;
        }
        if (include_ss) {
            T direct = T(P.csz / PI) * trans_bottom * (gen ? T(surf->lsun[los]) : albedo_in);
            return direct + diffuse;
        }
        return diffuse;
    }
    // Source of layer p toward line of sight j for order m, J + V + Q E (full layer, x = 0).  wL / wM (optional): its
    // partial derivatives w.r.t. L_i, M_i at fixed layer quantities (Y+ h+ and Y- h-).
    T layer_source(int m, int j, int p, const Layers<T>& Ly, const LayerSolution<T>& S, bool include_ss,
                   double* wL = nullptr, double* wM = nullptr) const {
        const int N = P.N, nstr = P.nstr;
        const double mu = P.los_mu[j];
        const T& od = Ly.od[p];
        const T& s = Ly.secant[p];
        const T& t = Ly.trans[p];
        // scat_phase_f (sktran_do_lpproduct.h:265-337); note the swapped plus/minus arguments at
        // sktran_do_opticallayer.cpp:128-130: "plus" carries the (-1)^(l-m) factor.
        std::vector<T> lps_plus(N), lps_minus(N);
        for (int q = 0; q < N; ++q) {
            T a(0.0), bneg(0.0);
            for (int l = m; l < nstr; ++l) {
                double pp = P.LPlos(j, m, l) * P.LPmu(m, q, l);
                a += Ly.beta[p][l] * T(pp);
                bneg += Ly.beta[p][l] * T(((l - m) % 2 != 0) ? -pp : pp);
            }
            lps_minus[q] = a * T(0.5 * P.wt[q]) * Ly.ssa[p];
            lps_plus[q] = bneg * T(0.5 * P.wt[q]) * Ly.ssa[p];
        }
        T Q(0.0);
        if (include_ss) {
            T acc(0.0);
            for (int l = m; l < nstr; ++l) {
                double pp = P.LPlos(j, m, l) * P.LPcsz(m, l);
                acc += Ly.beta[p][l] * T(((l - m) % 2 != 0) ? -pp : pp);
            }
            double factor = (2.0 - (m == 0 ? 1.0 : 0.0)) * (1.0 / (4.0 * PI));
            Q = acc * T(factor) * Ly.ssa[p];
        }
        // E (x = 0)
        T e2s = exp(-od * s) * exp(-od / T(mu));
        T E = t / (T(1.0) + T(mu) * s) * (T(1.0) - e2s);
        T expfactor = exp(-od * s);
        // E_thermal at x = 0 (sktran_do_opticallayer.cpp:941-957)
        const bool thermal = (m == 0) && !Ly.b0.empty() && !S.Ath.empty();
        T E_th(0.0);
        if (thermal) {
            const double b0 = Ly.b0[p], b1 = Ly.b1[p];
            E_th = T(b0 / (1.0 + mu * b1)) * (T(1.0) - exp(-od * T(b1)) * exp(-od / T(mu)));
        }
        T J(0.0), V(0.0);
        for (int i = 0; i < N; ++i) {
            T Yp(0.0), Ym(0.0);
            for (int q = 0; q < N; ++q) {
                Yp += lps_plus[q] * S.Wp[q + i * N] + lps_minus[q] * S.Wm[q + i * N];
                Ym += lps_plus[q] * S.Wm[q + i * N] + lps_minus[q] * S.Wp[q + i * N];
            }
            const T& k = S.k[i];
            T hp, hm;
            {
                T den = T(1.0) + T(mu) * k;
                if (std::abs(val(den)) > 0.0001) {
                    T e2 = exp(-od * k) * exp(-od / T(mu));
                    hp = (T(1.0) - e2) / den;
                } else {
                    hp = od / T(mu) * (T(1.0) - od * (k + T(1.0 / mu)));
                }
            }
            {
                T den = T(1.0) - T(mu) * k;
                if (std::abs(val(den)) > 0.0001 && stable_multipliers_ref()) {
                    hm = od / T(mu) * psi(od, k, T(1.0 / mu), exp(-k * od), exp(-od / T(mu)));
                } else if (std::abs(val(den)) > 0.0001) {
                    T e1 = exp(-k * od);
                    T e2 = exp(-od / T(mu));
                    hm = (e1 - e2) / den;
                } else {
                    T e1 = exp(-k * od);
                    hm = e1 * od / T(mu) * (T(1.0) - od * (k - T(1.0 / mu)));
                }
            }
            J += Yp * hp * S.Lc[i];
            J += Ym * hm * S.Mc[i];
            if (wL) {
                wL[i] = val(Yp) * val(hp);
                wM[i] = val(Ym) * val(hm);
            }
            T Dp = (-t * expfactor * hm + E) / (s + k);
            T Dm = stable_multipliers_ref()
                       ? t * (T(mu) * hp - od * exp(-od / T(mu)) * psi(od, k, s, exp(-k * od), expfactor)) /
                             (T(1.0) + T(mu) * s)
                       : (t * hp - E) / (s - k);
            V += S.Ap[i] * Yp * Dm + S.Am[i] * Ym * Dp;
            if (thermal) {
                // thermal part of V (sktran_do_opticallayer.cpp:421-478); the reference has no series branch for
                // b1 -> k_i here
                const double b0 = Ly.b0[p], b1 = Ly.b1[p];
                T e_b1 = exp(-od * T(b1));
                T Dp_th = (-T(b0) * e_b1 * hm + E_th) / (T(b1) + k);
                T Dm_th = (T(b0) * hp - E_th) / (T(b1) - k);
                V += S.Ath[i] * Yp * Dm_th + S.Ath[i] * Ym * Dp_th;
            }
        }
        T src = J + V + Q * E;
        if (thermal) src += E_th * (T(1.0) - Ly.ssa[p]);   // :524-531
        return src;
    }
    T los_component(int m, int j, const Layers<T>& Ly, const T& albedo, const std::vector<LayerSolution<T>>& sol,
                    bool include_ss) const {
        const int L = P.L;
        const double mu = P.los_mu[j];
        T I = ground_term(m, Ly.trans[L], Ly.od[L - 1], albedo, sol[L - 1], include_ss, nullptr, nullptr, j);
        // surface emission leaves the ground unreflected; the reference adds it inside its direct-bounce branch
        // (sktran_do_layerarray.cpp:225-266)
        if (m == 0 && include_ss) I += T(Ly.surface_emission);
        for (int p = L - 1; p >= 0; --p) {
            I = I * exp(-Ly.od[p] / T(mu));
            I += layer_source(m, j, p, Ly, sol[p], include_ss);
        }
        return I;
    }

    // ---- one wavelength.  radiance[nlos]; layer-lane derivatives dlane[nlos][nd] (T = Dual only)
    void solve_wavelength(const WavelInputs& in, double* radiance, double* dlane, Layers<T>* layers_out = nullptr) {
        const int L = P.L, nlos = P.nlos;
        lanes.L = L;
        lanes.G = in.d_leg ? in.ngroups : 0;
        Layers<T> Ly;
        layer_optics(in, Ly);
        T albedo(in.albedo);
        seed(albedo, lanes.albedo());
        std::vector<T> rad(nlos, T(0.0));
        std::vector<LayerSolution<T>> sol(L);
        SurfaceExpansion sx;
        for (int m = 0; m < in.num_azimuth; ++m) {
            for (int p = 0; p < L; ++p) {
                homogeneous(m, Ly.ssa[p], Ly.beta[p], sol[p]);
                particular(m, Ly.ssa[p], Ly.beta[p], Ly.od[p], Ly.secant[p], Ly.trans[p], sol[p]);
                sol[p].Ath.clear();
                if (in.emission && m == 0) {   // sktran_do_rte.cpp:154-156
                    if (ndual(albedo) != 0) throw std::runtime_error("oracle: weighting functions with thermal emission are not restated");
                    particular_thermal(Ly.ssa[p], Ly.od[p], Ly.b0[p], Ly.b1[p], sol[p]);
                }
            }
            if (in.brdf_kind != 0) {   // Surface::calculate(m), sktran_do_surface.h:153-217
                // The BRDF does not depend on the atmosphere: its Fourier coefficients enter the ground rows and the
                // ground-leaving term as constants, and the forward-mode lanes of the atmosphere go through unchanged
                // (the albedo lane is meaningless here - derivatives w.r.t. BRDF arguments are not restated).
                const int N = P.N;
                sx.general = true;
                sx.ss.assign(size_t(N) * N, 0.0);
                sx.sun.assign(N, 0.0);
                sx.los.assign(size_t(nlos) * N, 0.0);
                sx.lsun.assign(nlos, 0.0);
                for (int i = 0; i < N; ++i) {
                    for (int q = 0; q < N; ++q) sx.ss[i * N + q] = compute_expansion(m, in.brdf_kind, in.brdf_args, P.mu[i], P.mu[q]);
                    sx.sun[i] = compute_expansion(m, in.brdf_kind, in.brdf_args, P.mu[i], P.csz);
                    for (int j = 0; j < nlos; ++j) sx.los[j * N + i] = compute_expansion(m, in.brdf_kind, in.brdf_args, P.los_mu[j], P.mu[i]);
                }
                for (int j = 0; j < nlos; ++j) sx.lsun[j] = compute_expansion(m, in.brdf_kind, in.brdf_args, P.los_mu[j], P.csz);
                surf = &sx;
            }
            bvp(m, Ly, albedo, sol);
            for (int j = 0; j < nlos; ++j) {
                T comp = los_component(m, j, Ly, albedo, sol, in.include_ss);
                rad[j] += comp * T(std::cos(m * P.los_az[j]));
            }
        }
        for (int j = 0; j < nlos; ++j) {
            radiance[j] = val(rad[j]);
            store_derivs(rad[j], dlane ? dlane + size_t(j) * nd_ref() : nullptr);
        }
        surf = nullptr;
        if (layers_out) *layers_out = Ly;
    }
    static void store_derivs(const double&, double*) {}
    static void store_derivs(const Dual& x, double* out) {
        if (out)
            for (size_t i = 0; i < x.d.size(); ++i) out[i] = x.d[i];
    }
};


// ---------------------------------------------------------------------------------------------------
//  Reverse-mode linearisation (config.do_backprop = true): RTESolver::backprop, sktran_do_rte.cpp:1793-1895,
//  with the reference's layer-sparse duals (LayerDual: a layer quantity carries derivatives w.r.t. that layer's own
//  parameters only, sktran_do_types.h:233-330) instead of the dense forward-mode lanes of Solver<Dual>.
//
//  Per layer the K = G + 5 local lanes are [eps_g (scattering-group directions) | tau | omega | t (beam transmittance
//  at the layer top) | s (average secant) | albedo (bottom layer only)].  Per (order, line of sight):
//      I = sum_p A_p S_p(x; layer p) + A_L ground(x; layer L-1),      A_p = prod_{q<p} exp(-tau_q / mu)
//      dI = (dI at fixed BVP coefficients x) + z^T (db - dA x),       A^T z = dI/dx        (dgbtrs 'T', :1812-1836)
//  and the dependence of t_p, s_p on the optical depths of the other layers (m_trans_to_C+-, m_secant_to_C+-,
//  :1844-1893) is the chain  t_p = F0 exp(-sum_q chapman[p-1][q] tau_q),  s_p = (slant_{p+1} - slant_p) / tau_p.
//  Results are the same layer lanes as Solver<Dual> (checked against it by tests/test_oracle_wf.py).
// ---------------------------------------------------------------------------------------------------
template <int G>
struct ReverseSolver {
    static constexpr int K = G + 5;
    static constexpr int iTau = G, iOm = G + 1, iT = G + 2, iS = G + 3, iAlb = G + 4;
    using T = FDual<K>;
    const Plan& P;
    Solver<T> S;
    ReverseSolver(const Plan& p, dgeev_fn f) : P(p), S(p, f) {}

    struct Entry { int r, c, layer; T a; };
    struct Rhs { int r, layer; T b; };

    // radiance[nlos], dlane[nlos][L*(G+2)+1] in the lane order of `Lanes` (non-local)
    void solve_wavelength(const WavelInputs& in, double* radiance, double* dlane, Layers<T>* layers_out = nullptr) {
        const int L = P.L, N = P.N, nlos = P.nlos, n = 2 * N * L, kl = 3 * N - 1;
        if ((in.d_leg ? in.ngroups : 0) != G) throw std::runtime_error("ReverseSolver: group count mismatch");
        // layer optics with local lanes: scat(p, g) -> g, od -> iTau, ssa -> iOm
        S.lanes.L = L;
        S.lanes.G = G;
        S.lanes.local = true;
        Layers<T> Ly;
        S.layer_optics(in, Ly);
        // beam quantities become independent local variables (their cross-layer dependence is the chain at the end)
        std::vector<double> tval(L + 1), sval(L);
        for (int p = 0; p <= L; ++p) tval[p] = Ly.trans[p].v;
        for (int p = 0; p < L; ++p) {
            sval[p] = Ly.secant[p].v;
            Ly.secant[p] = T(sval[p]);
            Ly.secant[p].d[iS] = 1.0;
            Ly.trans[p] = T(tval[p]);
            Ly.trans[p].d[iT] = 1.0;
        }
        // bottom boundary in the bottom layer's lanes: t_L = t_{L-1} exp(-s_{L-1} tau_{L-1})
        Ly.trans[L] = Ly.trans[L - 1] * exp(-Ly.secant[L - 1] * Ly.od[L - 1]);
        T albedo(in.albedo);
        albedo.d[iAlb] = 1.0;

        const Lanes out_lanes{L, G, false};
        const int nd = out_lanes.total();
        std::vector<double> rad(nlos, 0.0);
        std::vector<double> dloc(size_t(nlos) * L * K, 0.0);  // [los][layer][lane], summed over orders
        std::vector<LayerSolution<T>> sol(L);
        std::vector<Entry> ent;
        std::vector<Rhs> rhs;
        std::vector<double> bval(n), x(n), wvec(n), srcv(L), wL(N), wM(N), gL(N), gM(N), Acum(L + 1);
        std::vector<T> src(L);
        BandLU lu;
        for (int m = 0; m < in.num_azimuth; ++m) {
            for (int p = 0; p < L; ++p) {
                S.homogeneous(m, Ly.ssa[p], Ly.beta[p], sol[p]);
                S.particular(m, Ly.ssa[p], Ly.beta[p], Ly.od[p], Ly.secant[p], Ly.trans[p], sol[p]);
            }
            assemble(m, Ly, albedo, sol, ent, rhs);
            lu.init(n, kl, kl);
            for (const auto& e : ent) lu.at(e.r, e.c) = e.a.v;
            if (lu.factor() != 0) throw std::runtime_error("BVP matrix singular");
            std::fill(bval.begin(), bval.end(), 0.0);
            for (const auto& r : rhs) bval[r.r] += r.b.v;
            x = bval;
            lu.solve(x.data());
            for (int p = 0; p < L; ++p) {
                sol[p].Lc.assign(N, T(0.0));
                sol[p].Mc.assign(N, T(0.0));
                for (int j = 0; j < N; ++j) {
                    sol[p].Lc[j] = T(x[p * 2 * N + j]);       // constants: derivatives w.r.t. x go through z
                    sol[p].Mc[j] = T(x[p * 2 * N + N + j]);
                }
            }
            for (int j = 0; j < nlos; ++j) {
                const double mu = P.los_mu[j], cosm = std::cos(m * P.los_az[j]);
                double* dl = &dloc[size_t(j) * L * K];
                Acum[0] = 1.0;
                for (int p = 0; p < L; ++p) Acum[p + 1] = Acum[p] * std::exp(-Ly.od[p].v / mu);
                double I = 0.0;
                for (int p = 0; p < L; ++p) {
                    src[p] = S.layer_source(m, j, p, Ly, sol[p], in.include_ss, wL.data(), wM.data());
                    for (int i = 0; i < N; ++i) {
                        wvec[p * 2 * N + i] = Acum[p] * wL[i];
                        wvec[p * 2 * N + N + i] = Acum[p] * wM[i];
                    }
                    I += Acum[p] * src[p].v;
                }
                T gnd = S.ground_term(m, Ly.trans[L], Ly.od[L - 1], albedo, sol[L - 1], in.include_ss, gL.data(), gM.data());
                for (int i = 0; i < N; ++i) {
                    wvec[(L - 1) * 2 * N + i] += Acum[L] * gL[i];
                    wvec[(L - 1) * 2 * N + N + i] += Acum[L] * gM[i];
                }
                I += Acum[L] * gnd.v;
                rad[j] += I * cosm;
                // derivatives at fixed x: sources, ground, line-of-sight attenuation
                double below = I;
                for (int p = 0; p < L; ++p) {
                    for (int k = 0; k < K; ++k) dl[p * K + k] += cosm * Acum[p] * src[p].d[k];
                    below -= Acum[p] * src[p].v;           // everything emitted below layer p carries exp(-tau_p / mu)
                    dl[p * K + iTau] += cosm * (-1.0 / mu) * below;
                }
                for (int k = 0; k < K; ++k) dl[(L - 1) * K + k] += cosm * Acum[L] * gnd.d[k];
                // adjoint of the boundary-value problem: A^T z = dI/dx, dI += z^T (db - dA x)
                lu.solve_transposed(wvec.data());
                for (const auto& e : ent) {
                    const double f = cosm * wvec[e.r] * x[e.c];
                    if (f != 0.0)
                        for (int k = 0; k < K; ++k) dl[e.layer * K + k] -= f * e.a.d[k];
                }
                for (const auto& r : rhs) {
                    const double f = cosm * wvec[r.r];
                    if (f != 0.0)
                        for (int k = 0; k < K; ++k) dl[r.layer * K + k] += f * r.b.d[k];
                }
            }
        }
        // local lanes -> layer lanes [scat g | od | ssa] + albedo, with the cross-layer chain of t_p and s_p
        for (int j = 0; j < nlos; ++j) {
            radiance[j] = rad[j];
            if (!dlane) continue;
            const double* dl = &dloc[size_t(j) * L * K];
            double* out = dlane + size_t(j) * nd;
            std::fill(out, out + nd, 0.0);
            for (int p = 0; p < L; ++p) {
                for (int g = 0; g < G; ++g) out[out_lanes.scat(p, g)] = dl[p * K + g];
                out[out_lanes.ssa(p)] = dl[p * K + iOm];
                out[out_lanes.od(p)] += dl[p * K + iTau] - dl[p * K + iS] * sval[p] / Ly.od[p].v;
                const double dIdt = dl[p * K + iT], dIds = dl[p * K + iS] / Ly.od[p].v;
                for (int q = 0; q < L; ++q) {
                    const double c1 = P.chapman[size_t(p) * L + q];
                    const double c0 = p > 0 ? P.chapman[size_t(p - 1) * L + q] : 0.0;
                    if (c1 == 0.0 && c0 == 0.0) continue;
                    out[out_lanes.od(q)] += dIds * (c1 - c0) - dIdt * tval[p] * c0;
                }
            }
            out[out_lanes.albedo()] = dl[(L - 1) * K + iAlb];
        }
        if (layers_out) *layers_out = Ly;
    }

    // bvp*Condition (sktran_do_rte.cpp:1898-2294) with every entry tagged by the one layer whose quantities it holds
    void assemble(int m, const Layers<T>& Ly, const T& albedo, const std::vector<LayerSolution<T>>& sol,
                  std::vector<Entry>& ent, std::vector<Rhs>& rhs) const {
        const int N = P.N, L = P.L, n = 2 * N * L;
        ent.clear();
        rhs.clear();
        std::vector<std::vector<T>> theta(L, std::vector<T>(N));
        for (int p = 0; p < L; ++p)
            for (int j = 0; j < N; ++j) theta[p][j] = exp(-Solver<T>::S_abs(sol[p].k[j]) * Ly.od[p]);
        for (int i = 0; i < N; ++i) {
            for (int j = 0; j < N; ++j) {
                ent.push_back({i, j, 0, sol[0].Wp[i + j * N]});
                ent.push_back({i, j + N, 0, sol[0].Wm[i + j * N] * theta[0][j]});
            }
            rhs.push_back({i, 0, -sol[0].Gpt[i]});
        }
        for (int bd = 1; bd < L; ++bd) {
            const int r0 = N + (bd - 1) * 2 * N, c0 = (bd - 1) * 2 * N, u = bd - 1, l = bd;
            const auto& U = sol[u];
            const auto& Lo = sol[l];
            for (int i = 0; i < N; ++i) {
                for (int j = 0; j < N; ++j) {
                    ent.push_back({r0 + i + N, c0 + j, u, U.Wp[i + j * N] * theta[u][j]});
                    ent.push_back({r0 + i + N, c0 + 2 * N + j, l, -Lo.Wp[i + j * N]});
                    ent.push_back({r0 + i, c0 + j, u, U.Wm[i + j * N] * theta[u][j]});
                    ent.push_back({r0 + i, c0 + 2 * N + j, l, -Lo.Wm[i + j * N]});
                    ent.push_back({r0 + i + N, c0 + N + j, u, U.Wm[i + j * N]});
                    ent.push_back({r0 + i + N, c0 + 3 * N + j, l, -(Lo.Wm[i + j * N] * theta[l][j])});
                    ent.push_back({r0 + i, c0 + N + j, u, U.Wp[i + j * N]});
                    ent.push_back({r0 + i, c0 + 3 * N + j, l, -(Lo.Wp[i + j * N] * theta[l][j])});
                }
                rhs.push_back({r0 + i, u, -U.Gmb[i]});
                rhs.push_back({r0 + i, l, Lo.Gmt[i]});
                rhs.push_back({r0 + i + N, u, -U.Gpb[i]});
                rhs.push_back({r0 + i + N, l, Lo.Gpt[i]});
            }
        }
        {
            const int r0 = N + (L - 1) * 2 * N, c0 = n - 2 * N, b = L - 1;
            const auto& B = sol[b];
            const bool refl = (m == 0);
            const double kd = (m == 0) ? 2.0 : 1.0;
            for (int i = 0; i < N; ++i) {
                for (int j = 0; j < N; ++j) {
                    T vm = B.Wm[i + j * N], vp = B.Wp[i + j * N];
                    if (refl)
                        for (int q = 0; q < N; ++q) {
                            vm -= T(kd) * albedo * T(P.wt[q] * P.mu[q]) * B.Wp[q + j * N];
                            vp -= T(kd) * albedo * T(P.wt[q] * P.mu[q]) * B.Wm[q + j * N];
                        }
                    ent.push_back({r0 + i, c0 + j, b, vm * theta[b][j]});
                    ent.push_back({r0 + i, c0 + N + j, b, vp});
                }
                T gds(0.0);
                if (refl) gds = T(P.csz) * albedo / T(PI) * Ly.trans[L];
                T um = B.Gmb[i];
                if (refl)
                    for (int q = 0; q < N; ++q) um -= T(kd) * albedo * T(P.wt[q] * P.mu[q]) * B.Gpb[q];
                rhs.push_back({r0 + i, b, gds - um});
            }
        }
    }
};

// Layer-lane derivatives -> native atmosphere derivatives [nloc*(2+G) + 1], restating
// sktran_do_layerarray.cpp:487-652 (group_and_triangle_fraction), :660-868 (per-wavelength "extinctions"
// factors) and do_source_planeparallel.cpp:160-179.
inline void map_to_native(const Plan& P, const Lanes& lanes, const WavelInputs& in, const std::vector<double>& tot_ext,
                          const std::vector<double>& scat_ext, const std::vector<double>& ssa_value, const double* dlane,
                          double* native) {
    const int L = P.L, nloc = P.nloc, G = lanes.G;
    const int nnative = nloc * (2 + G) + 1;
    std::fill(native, native + nnative, 0.0);
    for (int p = 0; p < L; ++p) {
        double dh = P.ceil_h[p] - P.floor_h[p];
        for (int q = 0; q < nloc; ++q) {
            double w = P.W[size_t(p) * nloc + q];
            if (!(w > 0)) continue;
            for (int g = 0; g < G; ++g)
                native[2 * nloc + nloc * g + q] += w * dlane[lanes.scat(p, g)] * (in.ssa[q] * in.ext[q] / scat_ext[p]);
            native[q] += w * dh * dlane[lanes.od(p)];
            native[q + nloc] += w * dlane[lanes.ssa(p)] * (in.ext[q] / tot_ext[p]);
            native[q] += w * dlane[lanes.ssa(p)] * ((in.ssa[q] - ssa_value[p]) / tot_ext[p]);
        }
    }
    native[nloc * (2 + G)] += dlane[lanes.albedo()];
}

}  // namespace oracle
