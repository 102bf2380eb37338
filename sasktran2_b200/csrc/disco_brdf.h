// Kernel-based surface BRDFs of the discrete-ordinates solve: a BRDF that is a linear combination of
// geometry-only kernels, brdf = sum_k args_k(wavelength) K_k(mu_in, mu_out, dphi) / pi, has Fourier coefficients that are
// the same combination of the kernels' coefficients.  The reference expands the full BRDF per wavelength and order by a
// 512-point quadrature (SurfaceStorage::compute_expansion, cpp/include/sktran_disco/sktran_do_surface.h:49-91; models
// cpp/include/sasktran2/atmosphere/surface.h:112-362); here the kernels are expanded once per engine on the host and the
// per-wavelength work is a 3-term combination on the device: MODIS (isotropic + Ross-thick + Li-sparse-R).
// The snow model of Kokhanovsky, r0 exp(-alpha(wavelength) K0 K0 / r0) / pi, is not linear in its argument: its
// geometry-only parts r0 and K0 K0 / r0 are tabulated per (angle pair, azimuth sample) on the host and the 512-point
// azimuth quadrature runs on the device per wavelength (k_brdf_expand_snow, 1 exp per sample).
#pragma once
#include <vector>

#include "disco_plan.h"

namespace disco {

constexpr int kBrdfLambertian = 0, kBrdfKokhanovsky = 1, kBrdfModis = 2;

struct BrdfTables {
    int kind = 0, nk = 0, nstr = 0, N = 0, nlos = 0;
    // per kernel k and azimuth order m, quadrature factors folded in:
    std::vector<double> Rss;    // [k][m][i][q]    (1 + delta_m0) rho^k_m(mu_i, mu_q) w_q mu_q
    std::vector<double> rsun;   // [k][m][i]       rho^k_m(mu_i, mu_0)
    std::vector<double> Rls;    // [k][m][los][q]  (1 + delta_m0) rho^k_m(mu_los, mu_q) w_q mu_q
    std::vector<double> rlsun;  // [k][m][los]     rho^k_m(mu_los, mu_0)
};

// Snow model: azimuth samples of the reference's quadrature folded to 0 <= phi <= pi (the model is even in phi)
struct SnowTables {
    int nsamples = 0, npairs = 0, N = 0, nlos = 0;
    std::vector<double> cosphi, weight;   // [nsamples]  cos(phi_s), quadrature weight (both mirror images)
    std::vector<double> r0, g;            // [pair][nsamples]  R0(mu_s, mu_v, theta), K0(mu_s) K0(mu_v) / R0
    std::vector<double> scale;            // [pair]  w_q mu_q for the stream-incidence pairs, 1 for the solar ones
    // pair order: stream-stream (i, q) -> i * N + q | stream-sun i -> N^2 + i | los-stream (los, q) -> N^2 + N + los * N + q | los-sun
};
SnowTables build_snow_tables(const HostPlan& plan);

int brdf_num_args(int kind);   // 1, 1, 3
// value of kernel k (already divided by pi) - exposed for tests
double brdf_kernel_value(int kind, int k, double mu_in, double mu_out, double phi_diff);
BrdfTables build_brdf_tables(int kind, const HostPlan& plan);

}  // namespace disco
