// Per-thread body of the dedicated two-stream kernel (host/device: tests/host_emul.cpp runs it on the CPU).
// Dedicated two-stream source (multiple_scatter_source = TwoStream, solar, scalar) for ground-viewing lines of sight:
// the reference's closed-form "explicit" path, cpp/lib/sktran_disco/cpp_twostream_source.cpp
//     prepare_explicit_column :1923-1974, forward_explicit_layers :2032-2122, build_and_solve_explicit_bvp :1976-2030,
//     pentadiagonal_solve :1861-1899, explicit_plane_view :2551-2660, exp_difference / integrated_exp_difference :33-131
// as ONE kernel, one thread per wavelength, one top -> bottom sweep over the layers, O(1) state per thread.
//
// What makes the single sweep possible: the radiance is linear in the boundary-value solution x (I = w.x + c, w = the
// line of sight's source multipliers).  The reference eliminates the pentadiagonal system top -> bottom (A = L~ U, z =
// L~^-1 b, unit upper-triangular U with super-diagonals alpha, beta) and back-substitutes bottom -> top.  Here
// w.x = w.U^-1 z = (U^-T w).z, and U^T is LOWER triangular: v = U^-T w is a forward recurrence
//     v_i = w_i - alpha_{i-1} v_{i-1} - beta_{i-2} v_{i-2}
// that runs in the same sweep as the elimination (the multipliers w_i of layer l only need the attenuation of the layers
// above it).  No factor is ever stored: per wavelength the kernel reads its 3 x nloc inputs once and writes nlos
// radiances (1.5 kB in, 16 B out at config 3), everything else lives in registers.  The pseudo-spherical
// chapman sums need the optical depths of the layers above: those sit in shared memory ([layer][thread], conflict
// free), next to the [L][L] chapman table, which the TMA engine (cp.async.bulk + mbarrier) brings in once per block.
//
// Roofline: ~700 FP64 instruction slots per (wavelength, layer) against 24 bytes of input: 60 flop/B, ten times the
// machine balance of B200 (37 TFLOP/s FP64 / 6.5 TB/s): FP64-pipe bound, not HBM bound.
#pragma once
#include "disco_bodies.h"

namespace disco {
namespace ts {

constexpr double kFourPi = 4.0 * kPi;

// ---- removable singularities, cpp_twostream_source.cpp:33-131 ------------------------------------------------------
// The reference evaluates these with library exp / pow and a convergence-tested series; the same functions are evaluated
// here from the layer's five exponentials (e^{-k od} per order, e^{-s od}, e^{-od / mu} per line of sight) that the sweep
// has anyway: every product e^{-(a + b) t} is a multiplication, e^{-x t / 2} a square root, the moment series a fixed
// Horner polynomial - optically thin layers (the top half of a line-by-line atmosphere) take these branches all the time.
// g(x) = (1 - e^{-x t}) / x = exp_difference(0, x, t); ex = e^{-x t} supplied by the caller
DISCO_HD double g_fun(double x, double t, double ex) {
    if (fabs(x * t) > 1.0e-5) return (1.0 - ex) / x;
    const double u = 0.5 * x * t, u2 = u * u;
    return t * sqrt(ex) * (1.0 + u2 * (1.0 / 6.0 + u2 * (1.0 / 120.0 + u2 / 5040.0)));
}
// (e^{-a t} - e^{-b t}) / (b - a) with ea, eb supplied
DISCO_HD double exp_difference(double a, double b, double t, double ea, double eb) {
    const double delta = b - a;
    if (fabs(delta * t) > 1.0e-5) return (ea - eb) / delta;
    const double u = 0.5 * delta * t, u2 = u * u;
    return t * sqrt(ea * eb) * (1.0 + u2 * (1.0 / 6.0 + u2 * (1.0 / 120.0 + u2 / 5040.0)));
}
// exp_moment(1, r, t) and exp_moment(3, r, t) (:33-63): M_n = int_0^t s^n e^{-r s} ds = t^{n+1} u_n(r t),
// u_n(x) = sum_j (-x)^j / (j! (n + j + 1)) for |x| < 0.5, else the upward recurrence u_n = (n u_{n-1} - e^{-x}) / x
// from u_0 = (1 - e^{-x}) / x.  er = e^{-r t}.
DISCO_HD void exp_moments_1_3(double r, double t, double er, double& m1, double& m3) {
    const double x = r * t, t2 = t * t;
    double u1, u3;
    if (fabs(x) < 0.5) {
        // coefficients 1 / (j! (j + 2)) and 1 / (j! (j + 4)), j = 0 .. 15 (0.5^16 / 16! < 1e-18)
        constexpr double c1[16] = {1.0 / 2, 1.0 / 3, 1.0 / 8, 1.0 / 30, 1.0 / 144, 1.0 / 840, 1.0 / 5760, 1.0 / 45360,
                                   1.0 / 403200, 1.0 / 3991680, 1.0 / 43545600, 1.0 / 518918400, 1.0 / 6706022400.0,
                                   1.0 / 93405312000.0, 1.0 / 1394852659200.0, 1.0 / 22230464256000.0};
        constexpr double c3[16] = {1.0 / 4, 1.0 / 5, 1.0 / 12, 1.0 / 42, 1.0 / 192, 1.0 / 1080, 1.0 / 7200, 1.0 / 55440,
                                   1.0 / 483840, 1.0 / 4717440, 1.0 / 50803200, 1.0 / 598752000, 1.0 / 7664025600.0,
                                   1.0 / 105859353600.0, 1.0 / 1569209241600.0, 1.0 / 24845812992000.0};
        const double y = -x;
        u1 = c1[15];
        u3 = c3[15];
#pragma unroll
        for (int j = 14; j >= 0; --j) {
            u1 = u1 * y + c1[j];
            u3 = u3 * y + c3[j];
        }
    } else {
        const double rx = 1.0 / x;
        const double u0 = (1.0 - er) * rx;
        u1 = (u0 - er) * rx;
        const double u2 = (2.0 * u1 - er) * rx;
        u3 = (3.0 * u2 - er) * rx;
    }
    m1 = t2 * u1;
    m3 = t2 * t2 * u3;
}
// [g(a) - g(b)] / (b - a), integrated_exp_difference :100-131; ga = g(a), gb = g(b) are values the sweep has already
DISCO_HD double integrated_exp_difference(double a, double b, double t, double ea, double eb, double ga, double gb) {
    const double delta = b - a;
    if (fabs(delta * t) > 1.0e-4) return (ga - gb) / delta;
    const double hd = 0.5 * delta;
    double m1, m3;
    exp_moments_1_3(0.5 * (a + b), t, sqrt(ea * eb), m1, m3);
    return m1 + hd * hd * m3 / 6.0;
}

DISCO_HD double positive_ratio(double num, double den) { return den > 0.0 ? num / den : 0.0; }

struct AzState {
    // elimination state of the last two rows (unit upper-triangular U: alpha, beta; transformed right-hand side z)
    double a1, a2, b1, b2, z1, z2;   // alpha_{i-1}, alpha_{i-2}, beta_{i-1}, beta_{i-2}, z_{i-1}, z_{i-2}
    // previous layer's homogeneous / particular quantities (continuity rows couple two layers)
    double xp, xm, om, gpb, gmb;
};

// one elimination row of pentadiagonal_solve (:1861-1899); i >= 2 form, also valid for the first rows with zeroed state
DISCO_HD void eliminate_row(AzState& S, double e, double c, double d, double a, double b, double rhs,
                                              double& alpha, double& beta, double& z) {
    const double gamma = c - S.a2 * e;
    const double inv = 1.0 / (d - S.b2 * e - S.a1 * gamma);
    alpha = (a - S.b1 * gamma) * inv;
    beta = b * inv;
    z = (rhs - S.z2 * e - S.z1 * gamma) * inv;
    S.a2 = S.a1;
    S.a1 = alpha;
    S.b2 = S.b1;
    S.b1 = beta;
    S.z2 = S.z1;
    S.z1 = z;
}


// One wavelength `w`, lines of sight los0 .. los0 + NLOS - 1.  chap: [L][L] chapman table; od_col / od_stride: scratch for
// the optical depths of the layers above (pseudo-spherical only; shared memory [layer][thread] on the device).
template <int NLOS>
DISCO_HD void twostream_body(const ChunkView& V, int w, int los0, const double* chap, double* od_col, int od_stride) {
    const int n = V.T.L, nloc = V.T.nloc, nlos_total = V.T.nlos;
    const bool pp = V.plane_parallel != 0;
    const double mu = 0.5, csz = V.T.csz, inv_csz = 1.0 / csz;
    const double* ext = V.ext + (size_t)nloc * w;
    const double* ssa = V.ssa + (size_t)nloc * w;
    const double* leg = V.leg + (size_t)V.nleg * nloc * w;
    const double* fdm = V.fdm ? V.fdm + (size_t)nloc * w : nullptr;
    const double irradiance = V.solar[w], albedo = V.albedo[w];
    const double angular = sqrt(fmax(0.0, (1.0 - mu * mu) * (1.0 - csz * csz)));

    // per line of sight
    double view_cos[NLOS], inv_view[NLOS], phase_mu[NLOS], phase_sine[NLOS], azw1[NLOS];
    double att[NLOS], integrated[NLOS];
    double v1[NLOS][2], v2[NLOS][2], dot[NLOS][2];   // U^-T w recurrence per azimuth order, (U^-T w).z
    bool live[NLOS];
#pragma unroll
    for (int j = 0; j < NLOS; ++j) {
        live[j] = los0 + j < nlos_total;
        const int jj = live[j] ? los0 + j : los0;
        view_cos[j] = V.T.los_mu[jj];
        inv_view[j] = 1.0 / view_cos[j];
        phase_mu[j] = view_cos[j] * mu;
        phase_sine[j] = 0.25 * sqrt(fmax(0.0, (1.0 - view_cos[j] * view_cos[j]) * (1.0 - mu * mu)));
        azw1[j] = V.T.los_cosmphi[(size_t)jj * V.T.nstr + 1];
        att[j] = 1.0;
        integrated[j] = 0.0;
#pragma unroll
        for (int az = 0; az < 2; ++az) v1[j][az] = v2[j][az] = dot[j][az] = 0.0;
    }
    AzState S[2];
#pragma unroll
    for (int az = 0; az < 2; ++az) S[az] = AzState{0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};

    auto b1_of = [&](int q) {   // load_b1, :879-897
        const double f = fdm ? fdm[q] : 0.0;
        return leg[1 + (size_t)V.nleg * q] - 3.0 * f / (1.0 - f);
    };
    // level above the current layer (top of the atmosphere first)
    double k_top = ext[nloc - 1], w_top = ssa[nloc - 1], b_top = b1_of(nloc - 1);
    double slant_top = 0.0;              // slant optical depth above the current layer's top

    for (int l = 0; l < n; ++l) {
        const int q = nloc - 2 - l;
        const double k_bot = ext[q], w_bot = ssa[q], b_bot = b1_of(q);
        // ---- prepare_explicit_column
        const double st = k_top * w_top, sb = k_bot * w_bot;
        const double avg_ext = 0.5 * (k_top + k_bot), avg_scat = 0.5 * (st + sb);
        const double od = avg_ext * V.layer_dh[l];
        const double om = fmin(positive_ratio(avg_scat, avg_ext), 1.0 - 1.0e-9);
        const double b1 = positive_ratio(0.5 * (st * b_top + sb * b_bot), avg_scat);
        double slant_bot;
        if (pp) {
            slant_bot = slant_top + od * inv_csz;   // chapman = 1 / mu0 on and below the diagonal: the same running sum
        } else {
            od_col[(size_t)l * od_stride] = od;
            slant_bot = 0.0;
            const double* row = chap + (size_t)l * n;
            for (int p = 0; p <= l; ++p) {
                const double f = row[p];
                if (f != 0.0) slant_bot += od_col[(size_t)p * od_stride] * f;
            }
        }
        const double rate = positive_ratio(slant_bot - slant_top, od);      // average secant
        const double trans = exp(-slant_top) * irradiance;                   // beam at the layer top
        const double expo = exp(-rate * od);
        // ---- line-of-sight exponentials of this layer
        double beam[NLOS], g_si[NLOS];
#pragma unroll
        for (int j = 0; j < NLOS; ++j) {
            beam[j] = exp(-od * inv_view[j]);
            g_si[j] = g_fun(rate + inv_view[j], od, expo * beam[j]);   // source_integral = g / mu = (1 - e^{-(s + 1/mu) od}) / (1 + s mu)
        }
        const bool last_layer = l == n - 1;
#pragma unroll
        for (int az = 0; az < 2; ++az) {
            // ---- forward_explicit_layers: closed-form eigenpair and Green's function coefficients
            double d, s;
            if (az == 0) {
                d = om * b1 * mu - 1.0 / mu;
                s = (om - 1.0) / mu;
            } else {
                d = -1.0 / mu;
                s = (om * b1 * (1.0 - mu * mu) - 2.0) / (2.0 * mu);
            }
            const double k = sqrt(s * d);
            const double s_over_k = s / k;
            const double xp = 0.5 * (1.0 - s_over_k), xm = 0.5 * (1.0 + s_over_k);
            const double omega = exp(-k * od);
            const double inv_norm = 1.0 / (mu * (xp * xp - xm * xm));
            double qp, qm;
            if (az == 0) {
                qp = om * (1.0 + b1 * csz * mu) / kFourPi;
                qm = om * (1.0 - b1 * csz * mu) / kFourPi;
            } else {
                qp = qm = om * b1 * angular / kFourPi;
            }
            const double ap = (qp * xp + qm * xm) * inv_norm, am = (qm * xp + qp * xm) * inv_norm;
            const double cp = trans * exp_difference(k, rate, od, omega, expo);
            const double g_sk = g_fun(rate + k, od, omega * expo);
            const double cm = trans * g_sk;
            const double gpt = am * cm * xm, gpb = ap * cp * xp, gmt = am * cm * xp, gmb = ap * cp * xm;
            // ---- build_and_solve_explicit_bvp: the rows that become complete with this layer
            AzState& A = S[az];
            double alpha_e = 0.0, beta_e = 0.0, z_e = 0.0;   // row 2l   (unknown L_l)
            double alpha_o = 0.0, z_o = 0.0, dummy;          // row 2l-1 (unknown M_{l-1}); z of row 2l+1 comes next layer
            if (l == 0) {
                eliminate_row(A, 0.0, 0.0, xp, xm * omega, 0.0, -gpt, alpha_e, beta_e, z_e);
            } else {
                eliminate_row(A, 0.0, A.xm * A.om, A.xp, -xm, -xp * omega, gmt - A.gmb, alpha_o, dummy, z_o);
                // the previous layer's M unknown: its v was formed last layer, its z is known now
#pragma unroll
                for (int j = 0; j < NLOS; ++j) dot[j][az] += v1[j][az] * z_o;
                eliminate_row(A, A.xp * A.om, A.xm, -xp, -xm * omega, 0.0, gpt - A.gpb, alpha_e, beta_e, z_e);
            }
            // ---- explicit_plane_view: multipliers of (L_l, M_l) and the particular source of this layer
            const double delta = az == 0 ? 1.0 : 0.0;
#pragma unroll
            for (int j = 0; j < NLOS; ++j) {
                double lp, lm;
                if (az == 0) {
                    lp = 0.5 * om * (1.0 - b1 * phase_mu[j]);
                    lm = 0.5 * om * (1.0 + b1 * phase_mu[j]);
                } else {
                    lp = lm = om * b1 * phase_sine[j];
                }
                const double yp = lp * xp + lm * xm, ym = lp * xm + lm * xp;
                const double hm = inv_view[j] * exp_difference(k, inv_view[j], od, omega, beam[j]);
                const double g_ki = g_fun(k + inv_view[j], od, omega * beam[j]);
                const double hp = inv_view[j] * g_ki;
                const double dp_ratio = inv_view[j] * integrated_exp_difference(rate + k, rate + inv_view[j], od, expo * omega,
                                                                                expo * beam[j], g_sk, g_si[j]);
                const double dm_ratio = inv_view[j] * integrated_exp_difference(k + inv_view[j], rate + inv_view[j], od, omega * beam[j],
                                                                                expo * beam[j], g_ki, g_si[j]);
                const double azw = az == 0 ? 1.0 : azw1[j];
                const double particular = ap * yp * (trans * dm_ratio) + am * ym * (trans * dp_ratio);
                integrated[j] += azw * particular * att[j];
                double wl = azw * yp * hp * att[j], wm = azw * ym * hm * att[j];
                if (last_layer && az == 0) {   // ground-leaving radiance 2 mu albedo (G+bot + L X+ omega + M X-), attenuated by every layer
                    const double att_ground = att[j] * beam[j] * (2.0 * mu) * albedo;
                    wl += att_ground * xp * omega;
                    wm += att_ground * xm;
                    integrated[j] += att_ground * gpb;
                }
                // v_{2l} = w_{2l} - alpha_{2l-1} v_{2l-1} - beta_{2l-2} v_{2l-2};  alpha_{2l-1} = alpha_o, beta_{2l-2}: the even row
                // before, whose b = 0 (beta = 0 for every even row)
                const double v_e = wl - alpha_o * v1[j][az];
                dot[j][az] += v_e * z_e;
                // v_{2l+1} = w_{2l+1} - alpha_{2l} v_{2l} - beta_{2l-1} v_{2l-1}
                const double v_o = wm - alpha_e * v_e - A.b2 * v1[j][az];
                v2[j][az] = v_e;
                v1[j][az] = v_o;
            }
            (void)beta_e;
            A.xp = xp;
            A.xm = xm;
            A.om = omega;
            A.gpb = gpb;
            A.gmb = gmb;
            if (last_layer) {
                // ground row (:2008-2028)
                const double direct = delta * csz * albedo / kPi * (exp(-slant_bot) * irradiance);
                const double rhs = direct - (gmb - 2.0 * delta * mu * albedo * gpb);
                double al, be, zl;
                eliminate_row(A, 0.0, (xm - 2.0 * mu * albedo * delta * xp) * omega, xp - 2.0 * mu * albedo * delta * xm, 0.0, 0.0, rhs,
                              al, be, zl);
#pragma unroll
                for (int j = 0; j < NLOS; ++j) dot[j][az] += v1[j][az] * zl;
            }
        }
#pragma unroll
        for (int j = 0; j < NLOS; ++j) att[j] *= beam[j];
        k_top = k_bot;
        w_top = w_bot;
        b_top = b_bot;
        slant_top = slant_bot;
    }
#pragma unroll
    for (int j = 0; j < NLOS; ++j)
        if (live[j]) V.radiance[(size_t)w * nlos_total + los0 + j] = integrated[j] + dot[j][0] + dot[j][1];
}

}  // namespace ts
}  // namespace disco
