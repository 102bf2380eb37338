#include "disco_plan.h"

#include <algorithm>
#include <cmath>
#include <stdexcept>

namespace disco {

// Gauss-Legendre rule of order n on (-1, 1), ascending nodes.  The reference reads the same numbers from
// the gauss-quad crate / 25-digit tables (sktran_do_quadrature.cpp:25-63); Newton's method in extended
// precision lands on the same doubles.
void gauss_rule(int n, std::vector<double>& nodes, std::vector<double>& weights) {
    nodes.resize(n);
    weights.resize(n);
    const long double pi = 3.14159265358979323846264338327950288L;
    for (int k = 0; k < (n + 1) / 2; ++k) {
        long double x = cosl(pi * (k + 0.75L) / (n + 0.5L));
        long double dp = 1.0L;
        for (int it = 0; it < 64; ++it) {
            long double p0 = 1.0L, p1 = x;  // P_0, P_1
            for (int j = 2; j <= n; ++j) {
                long double p2 = ((2 * j - 1) * x * p1 - (j - 1) * p0) / j;
                p0 = p1;
                p1 = p2;
            }
            if (n == 1) { p0 = 1.0L; p1 = x; }
            dp = n * (x * p1 - p0) / (x * x - 1.0L);
            long double step = p1 / dp;
            x -= step;
            if (fabsl(step) < 1e-19L) break;
        }
        // recompute derivative at the converged node
        long double p0 = 1.0L, p1 = x;
        for (int j = 2; j <= n; ++j) {
            long double p2 = ((2 * j - 1) * x * p1 - (j - 1) * p0) / j;
            p0 = p1;
            p1 = p2;
        }
        dp = n * (x * p1 - p0) / (x * x - 1.0L);
        long double w = 2.0L / ((1.0L - x * x) * dp * dp);
        nodes[k] = (double)(-x);
        nodes[n - 1 - k] = (double)x;
        weights[k] = weights[n - 1 - k] = (double)w;
    }
}

// d^l_{m0}(acos x) by the upward recurrence in l (cpp/include/sasktran2/math/wigner.h:56-149, n = 0).
double wigner_dm0(int m, int l, double coszen) {
    if (l < m) return 0.0;
    const double theta = std::acos(coszen);
    const double x = std::cos(theta);
    double fact = 1.0;  // (2m)! / (m! m!)
    for (int i = 2 * m; i > 1; --i) {
        fact *= double(i);
        if (i <= m) fact /= double(i);
        if (i <= m) fact /= double(i);
    }
    const int zeta = (m > 0 && (m % 2 != 0)) ? -1 : 1;
    double cur = zeta * std::pow(2.0, -double(m)) * std::sqrt(fact) * std::pow(1 - x, double(m) / 2.0) *
                 std::pow(1 + x, double(m) / 2.0);
    double prev = 0.0;
    for (int ll = m + 1; ll <= l; ++ll) {
        double mult = 1.0 / (std::sqrt(double(ll * ll - m * m)) * ll);
        double a = (2 * ll - 1) * (ll * x);
        double b = ll * std::sqrt(double((ll - 1) * (ll - 1) - m * m));
        double next = mult * (a * cur - b * prev);
        prev = cur;
        cur = next;
    }
    return cur;
}

namespace {

// Grid::calculate_interpolation_weights (cpp/lib/grids/grid.cpp:43-300), in-bounds cases
void grid_weights(const std::vector<double>& g, int interp, double x, int idx[2], double w[2]) {
    const int n = (int)g.size();
    idx[0] = idx[1] = -1;
    w[0] = w[1] = 0.0;
    if (interp == 2) {
        for (int i = 0; i + 1 < n; ++i)
            if (x + 0.1 >= g[i] && x < g[i + 1]) {
                idx[0] = i;
                w[0] = 1.0;
                return;
            }
        throw std::runtime_error("altitude grid: layer mid-point outside the grid");
    }
    bool uniform = true;
    const double d0 = g[1] - g[0];
    for (int i = 1; i < n; ++i) {
        double di = g[i] - g[i - 1];
        if (std::abs(di - d0) > 1e-12 * std::min(std::abs(di), std::abs(d0))) uniform = false;
    }
    int lo;
    double frac;
    if (uniform) {
        lo = (int)std::floor((x - g[0]) / d0);
        if (lo >= n - 1) throw std::runtime_error("altitude grid: layer mid-point outside the grid");
        frac = (x - g[lo]) / d0;
    } else {
        int hi = (int)(std::lower_bound(g.begin(), g.end(), x) - g.begin());
        if (hi == 0) hi = 1;
        lo = hi - 1;
        frac = (x - g[lo]) / (g[hi] - g[lo]);
    }
    idx[0] = lo;
    idx[1] = lo + 1;
    if (interp == 0) {
        w[0] = w[1] = 0.5;
    } else {
        w[1] = frac;
        w[0] = 1 - w[1];
    }
}

}  // namespace

HostPlan build_plan(int nstr, const GeometrySpec& geo, const std::vector<LineOfSight>& los) {
    if (nstr < 2 || (nstr % 2) != 0) throw std::runtime_error("number of streams must be even and >= 2");
    if (geo.altitudes.size() < 2) throw std::runtime_error("altitude grid needs at least two points");
    if (geo.geotype != 0 && geo.geotype != 1)
        throw std::runtime_error("B200 DO path supports plane-parallel and pseudo-spherical geometry only");
    if (!(geo.cos_sza > 0.0)) throw std::runtime_error("B200 DO path needs the sun above the horizon (cos_sza > 0)");
    HostPlan P;
    P.nstr = nstr;
    P.N = nstr / 2;
    P.nloc = (int)geo.altitudes.size();
    P.L = P.nloc - 1;
    P.nlos = (int)los.size();
    P.csz = geo.cos_sza;
    P.plane_parallel = geo.geotype == 0;
    P.interp = geo.interp;
    const int N = P.N, L = P.L;

    P.mu.assign(nstr, 0.0);
    P.wt.assign(nstr, 0.0);
    if (nstr == 2) {  // sktran_do_quadrature.cpp:16-23
        P.mu[0] = 0.5;
        P.mu[1] = -0.5;
        P.wt[0] = P.wt[1] = 1.0;
    } else {
        std::vector<double> x, w;
        gauss_rule(N, x, w);
        for (int i = 0; i < N; ++i) {
            P.mu[i] = 0.5 * x[i] + 0.5;
            P.mu[i + N] = -0.5 * x[i] - 0.5;
            P.wt[i] = P.wt[i + N] = 0.5 * w[i];
        }
    }
    P.lp_mu.assign((size_t)nstr * N * nstr, 0.0);
    P.lp_csz.assign((size_t)nstr * nstr, 0.0);
    P.lp_los.assign((size_t)P.nlos * nstr * nstr, 0.0);
    P.los_mu.resize(P.nlos);
    P.los_cosmphi.assign((size_t)P.nlos * nstr, 0.0);
    const double top = geo.altitudes.back();
    for (int j = 0; j < P.nlos; ++j) {
        if (los[j].cos_vza <= 0.0)
            throw std::runtime_error(
                "Error, currently only calculation of upwelling radiances is supported in plane parallel mode");
        if (los[j].observer_altitude < top)
            throw std::runtime_error("B200 DO path needs the observer at or above the top of the atmosphere");
        P.los_mu[j] = los[j].cos_vza;
        const double az = -los[j].rel_azimuth;  // do_source_planeparallel.cpp:619
        for (int m = 0; m < nstr; ++m) P.los_cosmphi[(size_t)j * nstr + m] = std::cos(m * az);
    }
    for (int m = 0; m < nstr; ++m)
        for (int l = 0; l < nstr; ++l) {
            for (int i = 0; i < N; ++i) P.lp_mu[((size_t)m * N + i) * nstr + l] = wigner_dm0(m, l, P.mu[i]);
            P.lp_csz[(size_t)m * nstr + l] = wigner_dm0(m, l, geo.cos_sza);
            for (int j = 0; j < P.nlos; ++j)
                P.lp_los[((size_t)j * nstr + m) * nstr + l] = wigner_dm0(m, l, P.los_mu[j]);
        }
    std::vector<double> ceil_h(L), floor_h(L);
    P.layer_dh.resize(L);
    for (int p = 0; p < L; ++p) {
        ceil_h[p] = geo.altitudes[P.nloc - 1 - p];
        floor_h[p] = geo.altitudes[P.nloc - 2 - p];
        P.layer_dh[p] = ceil_h[p] - floor_h[p];
        if (!(P.layer_dh[p] > 0)) throw std::runtime_error("altitude grid must be strictly ascending");
    }
    P.interp_idx.assign((size_t)L * 2, -1);
    P.interp_w.assign((size_t)L * 2, 0.0);
    for (int p = 0; p < L; ++p) {
        int idx[2];
        double w[2];
        grid_weights(geo.altitudes, geo.interp, 0.5 * (ceil_h[p] + floor_h[p]), idx, w);
        for (int c = 0; c < 2; ++c) {
            // the reference only uses strictly positive weights (sktran_do_layerarray.cpp:381-383)
            if (idx[c] >= 0 && w[c] > 0) {
                P.interp_idx[p * 2 + c] = idx[c];
                P.interp_w[p * 2 + c] = w[c];
            }
        }
    }
    P.chapman.assign((size_t)L * L, 0.0);
    if (P.plane_parallel) {
        for (int p = 0; p < L; ++p)
            for (int q = 0; q <= p; ++q) P.chapman[(size_t)p * L + q] = 1.0 / geo.cos_sza;
    } else {
        // Unrefracted straight solar ray from the floor of layer p (sktran_do_geometrylayerarray.cpp:75-118);
        // identical to the shell ray trace of :122-186 for a sun above the horizon.
        const double sin2 = 1 - geo.cos_sza * geo.cos_sza;
        for (int p = 0; p < L; ++p) {
            const double rp = geo.earth_radius + floor_h[p];
            for (int q = 0; q <= p; ++q) {
                const double rf = geo.earth_radius + floor_h[q], rc = geo.earth_radius + ceil_h[q];
                P.chapman[(size_t)p * L + q] =
                    (std::sqrt(rc * rc - rp * rp * sin2) - std::sqrt(rf * rf - rp * rp * sin2)) / (rc - rf);
            }
        }
    }
    return P;
}

}  // namespace disco
