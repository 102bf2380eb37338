// =====================================================================================================
//  ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the shipped product.
//
//  CPU restatement of SASKTRAN2's dedicated two-stream source (multiple_scatter_source = TwoStream, solar, scalar) for
//  ground-viewing lines of sight in plane-parallel / pseudo-spherical geometry: the "explicit" closed-form path of
//  cpp/lib/sktran_disco/cpp_twostream_source.cpp
//      prepare_explicit_column      :1923-1974   level -> layer optics (arithmetic mean of the two bounding levels), beam
//      forward_explicit_layers      :2032-2122   closed-form eigenpair, Green's function particular solution
//      build_and_solve_explicit_bvp :1976-2030   pentadiagonal boundary-value problem, pentadiagonal_solve :1861-1899
//      explicit_plane_view          :2551-2660   source-function integration toward one line of sight
//      exp_difference / integrated_exp_difference / exp_moment :33-131 ("resonant" removable singularities)
//  Inputs as loaded by the adapter (:864-930): level arrays top-down, b1 = leg_coeff[1] - 3 f / (1 - f), albedo.
//  It is a multiple-scatter-only source: no single-scatter term and no direct-beam bounce.
//
//  Pin (tests/test_oracle_twostream.py): the reference asserts this source equal to its two-stream discrete-ordinates
//  source with single scatter off to rtol 2e-8 on the inputs of tests/engine/test_twostream.py:104-160; the same inputs
//  are run through this restatement and through oracle::Solver (nstr = 2, include_ss = false), itself pinned to the
//  reference's DISORT tables.
// =====================================================================================================
#pragma once
#include <array>
#include <cmath>
#include <limits>
#include <vector>

#include "disco_oracle.hpp"

namespace oracle {
namespace twostream {

constexpr double FOUR_PI = 4.0 * PI;

// :33-63
inline double exp_moment(int order, double rate, double thickness) {
    if (thickness == 0.0) return 0.0;
    const double scaled_rate = rate * thickness;
    double unit_moment;
    if (std::abs(scaled_rate) < 0.5) {
        double factorial_term = 1.0;
        unit_moment = 0.0;
        for (int term = 0; term < 40; ++term) {
            const double contribution = factorial_term / double(order + term + 1);
            unit_moment += contribution;
            if (std::abs(contribution) <= std::numeric_limits<double>::epsilon() * std::max(std::abs(unit_moment), 1.0)) break;
            factorial_term *= -scaled_rate / double(term + 1);
        }
    } else {
        const double exponential = std::exp(-scaled_rate);
        unit_moment = -std::expm1(-scaled_rate) / scaled_rate;
        for (int current = 1; current <= order; ++current) unit_moment = (double(current) * unit_moment - exponential) / scaled_rate;
    }
    return std::pow(thickness, order + 1) * unit_moment;
}

// (e^{-a t} - e^{-b t}) / (b - a), :65-98 (value only)
inline double exp_difference(double a, double b, double thickness) {
    const double delta = b - a;
    const double scaled_delta = delta * thickness;
    if (std::abs(scaled_delta) > 1.0e-5) return (std::exp(-a * thickness) - std::exp(-b * thickness)) / delta;
    const double midpoint = 0.5 * (a + b);
    const double u = 0.5 * delta * thickness;
    const double u2 = u * u;
    const double sinhc = 1.0 + u2 * (1.0 / 6.0 + u2 * (1.0 / 120.0 + u2 / 5040.0));
    return thickness * std::exp(-midpoint * thickness) * sinhc;
}

// :100-131 (value only)
inline double integrated_exp_difference(double a, double b, double thickness) {
    const double delta = b - a;
    if (std::abs(delta * thickness) > 1.0e-4) return (exp_difference(0.0, a, thickness) - exp_difference(0.0, b, thickness)) / delta;
    const double midpoint = 0.5 * (a + b);
    const double half_delta = 0.5 * delta;
    return exp_moment(1, midpoint, thickness) + half_delta * half_delta * exp_moment(3, midpoint, thickness) / 6.0;
}

// exp_difference_ratio, :569-594
inline double exp_difference_ratio(double exp_a, double exp_b, double a, double b, double thickness) {
    const double delta = b - a;
    if (std::abs(delta * thickness) <= 1.0e-5) return exp_difference(a, b, thickness);
    return (exp_a - exp_b) / delta;
}

inline double positive_ratio(double num, double den) { return den > 0.0 ? num / den : 0.0; }  // :528-537

// plane_source_nonresonant, :1843-1859
inline bool nonresonant(double rate, double k, double od, double inverse_view) {
    const double d[5] = {rate + inverse_view, inverse_view - k, k + inverse_view, rate + k, rate - k};
    for (double x : d)
        if (std::abs(x * od) <= 1.0e-5) return false;
    return true;
}

struct Homogeneous {
    std::vector<double> k, xp, xm, omega, norm;
};
struct Particular {
    std::vector<double> ap, am, exponential, cp, cm, gpt, gpb, gmt, gmb;
};

// One wavelength.  Level arrays are indexed like the caller's grid (ascending altitude); P carries the geometry
// (layers top-down, chapman factors, lines of sight).  radiance[nlos].
inline void solve_wavelength(const Plan& P, const double* ext, const double* ssa, const double* b1_level, double irradiance,
                             double albedo, double* radiance) {
    const int n = P.L, nlev = P.nloc;
    const double mu = 0.5, csz = P.csz;  // ColumnGeometry::quadrature_cosine, :459
    auto lev = [&](const double* a, int top_down) { return a[nlev - 1 - top_down]; };   // load_* :864-877
    // ---- prepare_explicit_column<true>
    std::vector<double> od(n), w(n), b1(n), secant(n), transmission(n + 1), attenuation(n + 1);
    for (int l = 0; l < n; ++l) {
        const double st = lev(ext, l) * lev(ssa, l), sb = lev(ext, l + 1) * lev(ssa, l + 1);
        const double avg_ext = 0.5 * (lev(ext, l) + lev(ext, l + 1));
        const double avg_scat = 0.5 * (st + sb);
        od[l] = avg_ext * (P.ceil_h[l] - P.floor_h[l]);
        w[l] = std::min(positive_ratio(avg_scat, avg_ext), 1.0 - 1.0e-9);
        b1[l] = positive_ratio(0.5 * (st * lev(b1_level, l) + sb * lev(b1_level, l + 1)), avg_scat);
    }
    attenuation[0] = 0.0;
    transmission[0] = irradiance;
    for (int bd = 0; bd < n; ++bd) {
        double slant = 0.0;
        for (int l = 0; l < n; ++l) {
            const double f = P.chapman[size_t(bd) * n + l];
            if (f != 0.0) slant += od[l] * f;
        }
        attenuation[bd + 1] = -slant;
        transmission[bd + 1] = std::exp(-slant) * irradiance;
    }
    for (int l = 0; l < n; ++l) secant[l] = positive_ratio(attenuation[l] - attenuation[l + 1], od[l]);
    // ---- forward_explicit_layers<true>
    const double angular = std::sqrt(std::max(0.0, (1.0 - mu * mu) * (1.0 - csz * csz)));
    Homogeneous h[2];
    Particular p[2];
    std::vector<double> sol[2];
    for (int az = 0; az < 2; ++az) {
        for (auto* v : {&h[az].k, &h[az].xp, &h[az].xm, &h[az].omega, &h[az].norm}) v->assign(n, 0.0);
        for (auto* v : {&p[az].ap, &p[az].am, &p[az].exponential, &p[az].cp, &p[az].cm, &p[az].gpt, &p[az].gpb, &p[az].gmt, &p[az].gmb})
            v->assign(n, 0.0);
        for (int l = 0; l < n; ++l) {
            double d, s;
            if (az == 0) {
                d = w[l] * b1[l] * mu - 1.0 / mu;
                s = (w[l] - 1.0) / mu;
            } else {
                d = -1.0 / mu;
                s = (w[l] * b1[l] * (1.0 - mu * mu) - 2.0) / (2.0 * mu);
            }
            const double k = std::sqrt(s * d);
            const double s_over_k = s / k;
            h[az].k[l] = k;
            h[az].xp[l] = 0.5 * (1.0 - s_over_k);
            h[az].xm[l] = 0.5 * (1.0 + s_over_k);
            h[az].omega[l] = std::exp(-k * od[l]);
            h[az].norm[l] = mu * (h[az].xp[l] * h[az].xp[l] - h[az].xm[l] * h[az].xm[l]);
            double qp, qm;
            if (az == 0) {
                qp = w[l] * (1.0 + b1[l] * csz * mu) / FOUR_PI;
                qm = w[l] * (1.0 - b1[l] * csz * mu) / FOUR_PI;
            } else {
                qp = qm = w[l] * b1[l] * angular / FOUR_PI;
            }
            p[az].ap[l] = (qp * h[az].xp[l] + qm * h[az].xm[l]) / h[az].norm[l];
            p[az].am[l] = (qm * h[az].xp[l] + qp * h[az].xm[l]) / h[az].norm[l];
            const double rate = secant[l], amplitude = transmission[l];
            p[az].exponential[l] = std::exp(-rate * od[l]);
            const double cp_ratio = exp_difference_ratio(h[az].omega[l], p[az].exponential[l], k, rate, od[l]);
            const double cm_ratio = exp_difference_ratio(1.0, h[az].omega[l] * p[az].exponential[l], 0.0, rate + k, od[l]);
            p[az].cp[l] = amplitude * cp_ratio;
            p[az].cm[l] = amplitude * cm_ratio;
            p[az].gpt[l] = p[az].am[l] * p[az].cm[l] * h[az].xm[l];
            p[az].gpb[l] = p[az].ap[l] * p[az].cp[l] * h[az].xp[l];
            p[az].gmt[l] = p[az].am[l] * p[az].cm[l] * h[az].xp[l];
            p[az].gmb[l] = p[az].ap[l] * p[az].cp[l] * h[az].xm[l];
        }
        // ---- build_and_solve_explicit_bvp<true>
        const int size = 2 * n, last = size - 1;
        std::vector<double> e(size, 0.0), c(size, 0.0), dd(size, 0.0), a(size, 0.0), b(size, 0.0), rhs(size, 0.0);
        rhs[0] = -p[az].gpt[0];
        for (int l = 0; l < n - 1; ++l) {
            rhs[2 * l + 1] = p[az].gmt[l + 1] - p[az].gmb[l];
            rhs[2 * l + 2] = p[az].gpt[l + 1] - p[az].gpb[l];
        }
        const double delta = az == 0 ? 1.0 : 0.0;
        const double direct = delta * csz * albedo / PI * transmission[n];
        rhs[last] = direct - (p[az].gmb[n - 1] - 2.0 * delta * mu * albedo * p[az].gpb[n - 1]);
        dd[0] = h[az].xp[0];
        a[0] = h[az].xm[0] * h[az].omega[0];
        for (int l = 0; l < n - 1; ++l) {
            const int row = 2 * l;
            c[row + 1] = h[az].xm[l] * h[az].omega[l];
            dd[row + 1] = h[az].xp[l];
            a[row + 1] = -h[az].xm[l + 1];
            b[row + 1] = -h[az].xp[l + 1] * h[az].omega[l + 1];
            e[row + 2] = h[az].xp[l] * h[az].omega[l];
            c[row + 2] = h[az].xm[l];
            dd[row + 2] = -h[az].xp[l + 1];
            a[row + 2] = -h[az].xm[l + 1] * h[az].omega[l + 1];
        }
        e[last] = 0.0;
        c[last] = (h[az].xm[n - 1] - 2.0 * mu * albedo * delta * h[az].xp[n - 1]) * h[az].omega[n - 1];
        dd[last] = h[az].xp[n - 1] - 2.0 * mu * albedo * delta * h[az].xm[n - 1];
        a[last] = b[last] = 0.0;
        // pentadiagonal_solve, :1861-1899
        std::vector<double> inverse_mu(size), alpha(size, 0.0), beta(size, 0.0), gamma(size, 0.0), z(size);
        inverse_mu[0] = 1.0 / dd[0];
        alpha[0] = a[0] * inverse_mu[0];
        beta[0] = b[0] * inverse_mu[0];
        z[0] = rhs[0] * inverse_mu[0];
        if (size > 1) {
            gamma[1] = c[1];
            inverse_mu[1] = 1.0 / (dd[1] - alpha[0] * gamma[1]);
            alpha[1] = (a[1] - beta[0] * gamma[1]) * inverse_mu[1];
            beta[1] = b[1] * inverse_mu[1];
            z[1] = (rhs[1] - z[0] * gamma[1]) * inverse_mu[1];
        }
        for (int i = 2; i < size; ++i) {
            gamma[i] = c[i] - alpha[i - 2] * e[i];
            inverse_mu[i] = 1.0 / (dd[i] - beta[i - 2] * e[i] - alpha[i - 1] * gamma[i]);
            if (i + 1 < size) alpha[i] = (a[i] - beta[i - 1] * gamma[i]) * inverse_mu[i];
            if (i + 2 < size) beta[i] = b[i] * inverse_mu[i];
            z[i] = (rhs[i] - z[i - 2] * e[i] - z[i - 1] * gamma[i]) * inverse_mu[i];
        }
        rhs[size - 1] = z[size - 1];
        if (size > 1) rhs[size - 2] = z[size - 2] - alpha[size - 2] * rhs[size - 1];
        for (int i = size - 3; i >= 0; --i) rhs[i] = z[i] - alpha[i] * rhs[i + 1] - beta[i] * rhs[i + 2];
        sol[az] = rhs;
    }
    // ---- explicit_plane_view<true> per line of sight
    for (int j = 0; j < P.nlos; ++j) {
        const double view_cosine = P.los_mu[j], inverse_view = 1.0 / view_cosine;
        const double phase_mu = view_cosine * mu;
        const double phase_sine = 0.25 * std::sqrt(std::max(0.0, (1.0 - view_cosine * view_cosine) * (1.0 - mu * mu)));
        const double azimuth_weight[2] = {1.0, std::cos(P.los_az[j])};
        bool fast = true;
        for (int l = 0; l < n; ++l)
            for (int az = 0; az < 2; ++az) fast = fast && nonresonant(secant[l], h[az].k[l], od[l], inverse_view);
        double att = 1.0, integrated = 0.0;
        for (int l = 0; l < n; ++l) {
            const double beam = std::exp(-od[l] * inverse_view);
            const double exponential = p[0].exponential[l], rate = secant[l];
            const double source_integral = fast ? (1.0 - exponential * beam) / (1.0 + rate * view_cosine)
                                                : inverse_view * exp_difference(0.0, rate + inverse_view, od[l]);
            double source = 0.0;
            for (int az = 0; az < 2; ++az) {
                double lp, lm;   // explicit_lpsum<true>, :2144-2157
                if (az == 0) {
                    lp = 0.5 * w[l] * (1.0 - b1[l] * phase_mu);
                    lm = 0.5 * w[l] * (1.0 + b1[l] * phase_mu);
                } else {
                    lp = lm = w[l] * b1[l] * phase_sine;
                }
                const double xp = h[az].xp[l], xm = h[az].xm[l], k = h[az].k[l], omega = h[az].omega[l];
                const double yp = lp * xp + lm * xm, ym = lp * xm + lm * xp;
                double hm, hp, dp_ratio, dm_ratio;
                if (fast) {
                    hm = (omega - beam) / (1.0 - k * view_cosine);
                    hp = (1.0 - omega * beam) / (1.0 + k * view_cosine);
                    dp_ratio = (source_integral - exponential * hm) / (rate + k);
                    dm_ratio = (hp - source_integral) / (rate - k);
                } else {
                    hm = inverse_view * exp_difference(k, inverse_view, od[l]);
                    hp = inverse_view * exp_difference(0.0, k + inverse_view, od[l]);
                    dp_ratio = inverse_view * integrated_exp_difference(rate + k, rate + inverse_view, od[l]);
                    dm_ratio = inverse_view * integrated_exp_difference(k + inverse_view, rate + inverse_view, od[l]);
                }
                const double particular = p[az].ap[l] * yp * (transmission[l] * dm_ratio) + p[az].am[l] * ym * (transmission[l] * dp_ratio);
                source += azimuth_weight[az] * (sol[az][2 * l] * yp * hp + sol[az][2 * l + 1] * ym * hm + particular);
            }
            integrated += source * att;
            att *= beam;
        }
        const int last = n - 1;
        const double base_surface = p[0].gpb[last] + sol[0][2 * last] * h[0].xp[last] * h[0].omega[last] + sol[0][2 * last + 1] * h[0].xm[last];
        radiance[j] = integrated + att * (base_surface * (2.0 * mu) * albedo);
    }
}

}  // namespace twostream
}  // namespace oracle
