// Kernel-based surface BRDFs of the discrete-ordinates solve: a BRDF that is a linear combination of
// geometry-only kernels, brdf = sum_k args_k(wavelength) K_k(mu_in, mu_out, dphi) / pi, has Fourier coefficients that are
// the same combination of the kernels' coefficients.  The reference expands the full BRDF per wavelength and order by a
// 512-point quadrature (SurfaceStorage::compute_expansion, cpp/include/sktran_disco/sktran_do_surface.h:49-91; models
// cpp/include/sasktran2/atmosphere/surface.h:112-362); here the kernels are expanded once per engine on the host and the
// per-wavelength work is a 3-term combination on the device.  Supported: MODIS (isotropic + Ross-thick + Li-sparse-R).
// The snow model of Kokhanovsky is not linear in its argument and is refused.
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

int brdf_num_args(int kind);   // 1, 1, 3
// value of kernel k (already divided by pi) - exposed for tests
double brdf_kernel_value(int kind, int k, double mu_in, double mu_out, double phi_diff);
BrdfTables build_brdf_tables(int kind, const HostPlan& plan);

}  // namespace disco
