#include "disco_brdf.h"

#include <algorithm>
#include <cmath>
#include <stdexcept>

namespace disco {
namespace {
constexpr double kPiB = 3.14159265358979323846;
}

int brdf_num_args(int kind) { return kind == kBrdfModis ? 3 : 1; }

double brdf_kernel_value(int kind, int k, double mu_in, double mu_out, double phi_diff) {
    if (kind != kBrdfModis) throw std::runtime_error("brdf_kernel_value: only the MODIS model is kernel based");
    if (k == 0) return 1.0 / kPiB;   // isotropic
    // Ross-thick (volumetric) and Li-sparse-R (geometric) kernels, cpp/include/sasktran2/atmosphere/surface.h:252-290
    const double cs = mu_in, cv = mu_out;
    const double ss = std::sqrt(1 - cs * cs), sv = std::sqrt(1 - cv * cv);
    const double ts = ss / cs, tv = sv / cv;
    const double craa = -std::cos(phi_diff), sraa = std::sin(phi_diff);   // raa = 0 is backscatter in the kernels' convention
    const double csa = std::max(-1.0, std::min(1.0, cs * cv + ss * sv * craa));
    const double sa = std::acos(csa);
    if (k == 1) return (((0.5 * kPiB - sa) * csa + std::sin(sa)) / (cs + cv) - 0.25 * kPiB) / kPiB;
    const double d2 = ts * ts + tv * tv - 2 * ts * tv * craa;
    const double ct = std::max(-1.0, std::min(1.0, 2 * std::sqrt(d2 + ts * ts * tv * tv * sraa * sraa) * cs * cv / (cs + cv)));
    const double t = std::acos(ct);
    const double o = (t - std::sin(t) * ct) * (cs + cv) / (kPiB * cs * cv);
    return (o - (cs + cv - 0.5 * (1 + csa)) / (cs * cv)) / kPiB;
}

SnowTables build_snow_tables(const HostPlan& plan) {
    SnowTables T;
    const int N = plan.N, nlos = plan.nlos;
    T.N = N;
    T.nlos = nlos;
    T.npairs = N * N + N + nlos * N + nlos;
    std::vector<double> qx, qw;
    gauss_rule(512, qx, qw);
    // the reference visits phi = pi (+-0.5 x +- 0.5) for the first 256 nodes with weight w / 2 each: two |phi| per node
    for (int i = 0; i < 256; ++i)
        for (double a : {0.5 * qx[i] + 0.5, -0.5 * qx[i] + 0.5}) {
            T.cosphi.push_back(std::cos(kPiB * a));
            T.weight.push_back(2.0 * 0.5 * qw[i]);
        }
    T.nsamples = (int)T.cosphi.size();
    T.r0.assign((size_t)T.npairs * T.nsamples, 0.0);
    T.g.assign((size_t)T.npairs * T.nsamples, 0.0);
    T.scale.assign(T.npairs, 1.0);
    auto fill = [&](int pair, double mu_out, double mu_in, double scale) {
        // SnowKokhanovsky::brdf, cpp/include/sasktran2/atmosphere/surface.h:151-199
        const double mus = mu_in, muv = mu_out;
        const double ss = std::sqrt(1 - mus * mus), sv = std::sqrt(1 - muv * muv);
        const double k0k0 = (3.0 / 7.0) * (1.0 + 2.0 * mus) * (3.0 / 7.0) * (1.0 + 2.0 * muv);
        for (int s = 0; s < T.nsamples; ++s) {
            const double cost = std::max(-1.0, std::min(1.0, -mus * muv + ss * sv * T.cosphi[s]));
            const double theta = std::acos(cost) * 180.0 / kPiB;
            const double p = 11.1 * std::exp(-0.087 * theta) + 1.1 * std::exp(-0.014 * theta);
            const double r0 = (1.247 + 1.186 * (mus + muv) + 5.157 * mus * muv + p) / (4.0 * (mus + muv));
            T.r0[(size_t)pair * T.nsamples + s] = r0;
            T.g[(size_t)pair * T.nsamples + s] = k0k0 / r0;
        }
        T.scale[pair] = scale;
    };
    for (int i = 0; i < N; ++i) {
        for (int q = 0; q < N; ++q) fill(i * N + q, plan.mu[i], plan.mu[q], plan.wt[q] * plan.mu[q]);
        fill(N * N + i, plan.mu[i], plan.csz, 1.0);
    }
    for (int j = 0; j < nlos; ++j) {
        for (int q = 0; q < N; ++q) fill(N * N + N + j * N + q, plan.los_mu[j], plan.mu[q], plan.wt[q] * plan.mu[q]);
        fill(N * N + N + nlos * N + j, plan.los_mu[j], plan.csz, 1.0);
    }
    return T;
}

BrdfTables build_brdf_tables(int kind, const HostPlan& plan) {
    if (kind != kBrdfModis) throw std::runtime_error("build_brdf_tables: only the MODIS model is kernel based");
    BrdfTables T;
    T.kind = kind;
    T.nk = 3;
    T.nstr = plan.nstr;
    T.N = plan.N;
    T.nlos = plan.nlos;
    const int N = plan.N, M = plan.nstr, nlos = plan.nlos, nk = T.nk;
    // azimuth samples of the reference's rule: the first half of the 512 Gauss-Legendre nodes, mirrored four ways
    std::vector<double> qx, qw;
    gauss_rule(512, qx, qw);
    std::vector<double> phi, wphi;
    for (int i = 0; i < 256; ++i)
        for (double a : {0.5 * qx[i] + 0.5, -0.5 * qx[i] + 0.5, 0.5 * qx[i] - 0.5, -0.5 * qx[i] - 0.5}) {
            phi.push_back(kPiB * a);
            wphi.push_back(0.5 * qw[i]);
        }
    std::vector<double> cosm((size_t)M * phi.size());
    for (int m = 0; m < M; ++m)
        for (size_t s = 0; s < phi.size(); ++s) cosm[(size_t)m * phi.size() + s] = std::cos(m * phi[s]);
    // rho^k_m(mu_out, mu_in) for every order
    auto expand = [&](int k, double mu_out, double mu_in, double* out_m) {
        std::vector<double> f(phi.size());
        for (size_t s = 0; s < phi.size(); ++s) f[s] = wphi[s] * brdf_kernel_value(kind, k, mu_in, mu_out, phi[s]);
        for (int m = 0; m < M; ++m) {
            double acc = 0.0;
            const double* c = &cosm[(size_t)m * phi.size()];
            for (size_t s = 0; s < phi.size(); ++s) acc += f[s] * c[s];
            out_m[m] = acc * 0.5 * kPiB * (m == 0 ? 1.0 : 2.0);
        }
    };
    T.Rss.assign((size_t)nk * M * N * N, 0.0);
    T.rsun.assign((size_t)nk * M * N, 0.0);
    T.Rls.assign((size_t)nk * M * nlos * N, 0.0);
    T.rlsun.assign((size_t)nk * M * nlos, 0.0);
    std::vector<double> e(M);
    for (int k = 0; k < nk; ++k) {
        for (int i = 0; i < N; ++i) {
            for (int q = 0; q < N; ++q) {
                expand(k, plan.mu[i], plan.mu[q], e.data());
                for (int m = 0; m < M; ++m)
                    T.Rss[(((size_t)k * M + m) * N + i) * N + q] = (m == 0 ? 2.0 : 1.0) * e[m] * plan.wt[q] * plan.mu[q];
            }
            expand(k, plan.mu[i], plan.csz, e.data());
            for (int m = 0; m < M; ++m) T.rsun[((size_t)k * M + m) * N + i] = e[m];
        }
        for (int j = 0; j < nlos; ++j) {
            for (int q = 0; q < N; ++q) {
                expand(k, plan.los_mu[j], plan.mu[q], e.data());
                for (int m = 0; m < M; ++m)
                    T.Rls[(((size_t)k * M + m) * nlos + j) * N + q] = (m == 0 ? 2.0 : 1.0) * e[m] * plan.wt[q] * plan.mu[q];
            }
            expand(k, plan.los_mu[j], plan.csz, e.data());
            for (int m = 0; m < M; ++m) T.rlsun[((size_t)k * M + m) * nlos + j] = e[m];
        }
    }
    return T;
}

}  // namespace disco
