// ORACLE — TEST INFRASTRUCTURE ONLY (see disco_oracle.hpp).  extern "C" surface for ctypes.
#include "disco_oracle.hpp"
#include "twostream_oracle.hpp"
#include "limb_oracle.hpp"

#include <string>
#ifdef _OPENMP
#include <omp.h>
#endif

static thread_local std::string g_err;
static std::string g_last_err;

template <int G>
static void reverse_wavelength(const oracle::Plan& P, oracle::dgeev_fn dgeev, const oracle::WavelInputs& in, int nlos, int nd,
                               int nnative, double* radiance, double* native, double* lanes_out) {
    using namespace oracle;
    ReverseSolver<G> R(P, dgeev);
    std::vector<double> dlane(size_t(nlos) * nd);
    Layers<typename ReverseSolver<G>::T> Ly;
    R.solve_wavelength(in, radiance, dlane.data(), &Ly);
    Lanes lanes{P.L, G, false};
    for (int j = 0; j < nlos; ++j) {
        if (native) map_to_native(P, lanes, in, Ly.tot_ext, Ly.scat_ext, Ly.ssa_value, &dlane[size_t(j) * nd], native + size_t(j) * nnative);
        if (lanes_out) std::memcpy(lanes_out + size_t(j) * nd, &dlane[size_t(j) * nd], sizeof(double) * nd);
    }
}


extern "C" {

const char* oracle_last_error() { return g_last_err.c_str(); }

// 0: the reference's multiplier formulas (default); 1: removable singularities evaluated with phi/psi
void oracle_set_stable_multipliers(int on) { oracle::stable_multipliers_ref() = on; }

// delta-M state of the next oracle_do_radiance call: f [nloc, nwavel], d_f [nloc, nwavel, ngroups] (null: unscaled)
static const double* g_f = nullptr;
static const double* g_df = nullptr;
void oracle_set_delta_m(const double* f, const double* d_f) {
    g_f = f;
    g_df = d_f;
}

// surface BRDF of the next oracle_do_radiance call: kind 0 Lambertian (the albedo argument), 1 snow (Kokhanovsky),
// 2 MODIS; args [nargs, nwavel] column-major (Surface::brdf_args)
static int g_brdf_kind = 0, g_brdf_nargs = 1;
static const double* g_brdf_args = nullptr;
void oracle_set_brdf(int kind, int nargs, const double* args) {
    g_brdf_kind = kind;
    g_brdf_nargs = nargs;
    g_brdf_args = args;
}
// thermal emission of the next oracle_do_radiance call (config.emission_source = discrete_ordinates): emission_source
// [nloc, nwavel] column-major (nullptr: none) and the surface emission [nwavel] (nullptr: zero)
static const double* g_emission = nullptr;
static const double* g_surface_emission = nullptr;
void oracle_set_emission(const double* emission, const double* surface_emission) {
    g_emission = emission;
    g_surface_emission = surface_emission;
}
// Fourier coefficient of a BRDF model (SurfaceStorage::compute_expansion) and the model itself, for the tests
double oracle_brdf_expansion(int m, int kind, const double* args, double mu_out, double mu_in) {
    return oracle::compute_expansion(m, kind, args, mu_out, mu_in);
}
double oracle_brdf_value(int kind, const double* args, double mu_in, double mu_out, double phi_diff) {
    return oracle::brdf_value(kind, args, mu_in, mu_out, phi_diff);
}

// 0: forward-mode dense duals (the reference's default, do_backprop = false); 1: reverse mode (do_backprop = true,
// RTESolver::backprop): layer-local duals + transposed band solves per line of sight
static int g_reverse = 0;
void oracle_set_reverse_mode(int on) { g_reverse = on; }

// 1: pseudo-spherical chapman factors from the closed straight-line formula instead of the reference's ray tracer
void oracle_set_chapman_straight_line(int on) { oracle::chapman_straight_line_ref() = on; }

int oracle_num_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// Arrays follow the reference's C-ABI layouts: ssa/ext [nloc, nwavel] column-major (q + nloc*w),
// leg [nleg, nloc, nwavel], d_leg [nleg, nloc, nwavel, ngroups], radiance [nwavel, nlos] C-order,
// native [nwavel, nlos, nloc*(2+ngroups)+1], lanes [nwavel, nlos, L*(ngroups+2)+1].
int oracle_do_radiance(int nstr, int nloc, int nwavel, int nleg, int nlos, const double* alt, int interp, int geotype,
                       double cos_sza, double earth_radius, const double* los_cos_vza, const double* los_rel_az,
                       const double* ssa, const double* ext, const double* leg, const double* solar,
                       const double* albedo, const double* d_leg, int ngroups, int include_ss, int num_azimuth,
                       int calc_derivs, int nthreads, void* dgeev_ptr, double* radiance, double* native,
                       double* lanes_out) {
    using namespace oracle;
    try {
        std::vector<double> a(alt, alt + nloc), cz(los_cos_vza, los_cos_vza + nlos), az(los_rel_az, los_rel_az + nlos);
        Plan P = make_plan(nstr, a, interp, geotype, cos_sza, earth_radius, cz, az);
        dgeev_fn dgeev = (dgeev_fn)dgeev_ptr;
        const int G = (calc_derivs && d_leg) ? ngroups : 0;
        const int L = nloc - 1;
        const int nd = calc_derivs ? L * (G + 2) + 1 : 0;
        const int nnative = nloc * (2 + G) + 1;
        int failed = 0;
#ifdef _OPENMP
        if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 1)
        for (int w = 0; w < nwavel; ++w) {
            if (failed) continue;
            try {
                WavelInputs in;
                in.ext = ext + size_t(nloc) * w;
                in.ssa = ssa + size_t(nloc) * w;
                in.leg = leg + size_t(nleg) * nloc * w;
                in.nleg = nleg;
                in.f = g_f ? g_f + size_t(nloc) * w : nullptr;
                std::vector<double> dfw;
                if (g_f && g_df && calc_derivs && d_leg && ngroups > 0) {
                    dfw.resize(size_t(nloc) * ngroups);
                    for (int g = 0; g < ngroups; ++g)
                        std::memcpy(&dfw[size_t(nloc) * g], g_df + size_t(nloc) * (w + size_t(nwavel) * g), sizeof(double) * nloc);
                    in.d_f = dfw.data();
                }
                in.solar = solar[w];
                in.albedo = albedo[w];
                if (g_brdf_kind != 0 && g_brdf_args) {
                    in.brdf_kind = g_brdf_kind;
                    in.brdf_args = g_brdf_args + size_t(g_brdf_nargs) * w;
                }
                if (g_emission) in.emission = g_emission + size_t(nloc) * w;
                if (g_surface_emission) in.surface_emission = g_surface_emission[w];
                if ((g_emission || g_surface_emission) && calc_derivs)
                    throw std::runtime_error("oracle: weighting functions with thermal emission are not restated");
                in.ngroups = G;
                in.include_ss = include_ss != 0;
                in.num_azimuth = num_azimuth > 0 ? num_azimuth : nstr;
                if (!calc_derivs) {
                    in.d_leg = nullptr;
                    Solver<double> S(P, dgeev);
                    S.solve_wavelength(in, radiance + size_t(w) * nlos, nullptr);
                } else {
                    nd_ref() = nd;
                    // per-wavelength view of d_leg [nleg, nloc, nwavel, ngroups] -> [nleg, nloc, ngroups]
                    std::vector<double> dl;
                    if (G > 0) {
                        dl.resize(size_t(nleg) * nloc * G);
                        for (int g = 0; g < G; ++g)
                            std::memcpy(&dl[size_t(nleg) * nloc * g],
                                        d_leg + size_t(nleg) * nloc * (w + size_t(nwavel) * g),
                                        sizeof(double) * nleg * nloc);
                        in.d_leg = dl.data();
                    } else {
                        in.d_leg = nullptr;
                    }
                    if (g_reverse && in.brdf_kind != 0)
                        throw std::runtime_error("reverse-mode oracle: Lambertian surfaces only (use the forward-mode lanes)");
                    if (g_reverse) {
                        double* nat = native ? native + size_t(w) * nlos * nnative : nullptr;
                        double* lo = lanes_out ? lanes_out + size_t(w) * nlos * nd : nullptr;
                        double* rw = radiance + size_t(w) * nlos;
                        switch (G) {
                            case 0: reverse_wavelength<0>(P, dgeev, in, nlos, nd, nnative, rw, nat, lo); break;
                            case 1: reverse_wavelength<1>(P, dgeev, in, nlos, nd, nnative, rw, nat, lo); break;
                            case 2: reverse_wavelength<2>(P, dgeev, in, nlos, nd, nnative, rw, nat, lo); break;
                            case 3: reverse_wavelength<3>(P, dgeev, in, nlos, nd, nnative, rw, nat, lo); break;
                            default: throw std::runtime_error("reverse-mode oracle: at most 3 scattering groups");
                        }
                        nd_ref() = 0;
                        continue;
                    }
                    Solver<Dual> S(P, dgeev);
                    S.lanes.L = L;
                    S.lanes.G = G;
                    std::vector<double> dlane(size_t(nlos) * nd);
                    Layers<Dual> Ly;
                    S.solve_wavelength(in, radiance + size_t(w) * nlos, dlane.data(), &Ly);
                    for (int j = 0; j < nlos; ++j) {
                        if (native)
                            map_to_native(P, S.lanes, in, Ly.tot_ext, Ly.scat_ext, Ly.ssa_value, &dlane[size_t(j) * nd],
                                          native + (size_t(w) * nlos + j) * nnative);
                        if (lanes_out)
                            std::memcpy(lanes_out + (size_t(w) * nlos + j) * nd, &dlane[size_t(j) * nd], sizeof(double) * nd);
                    }
                    nd_ref() = 0;
                }
            } catch (const std::exception& e) {
#pragma omp critical
                {
                    failed = 1;
                    g_last_err = e.what();
                }
            }
        }
        return failed ? -3 : 0;
    } catch (const std::exception& e) {
        g_last_err = e.what();
        return -3;
    }
}

// Distance of every (wavelength, order, layer) cell to the removable singularities of the reference's multiplier
// formulas: d_sec = min_j |secant_p - k_j| (C+, D-: sktran_do_rte.cpp:1203-1226, sktran_do_opticallayer.cpp:897-938) and
// d_los = min_{j, los} |1 - mu_los k_j| (h-: sktran_do_opticallayer.cpp:339-344).  out: [nwavel][nstr][L][2].
int oracle_degeneracy(int nstr, int nloc, int nwavel, int nleg, int nlos, const double* alt, int interp, int geotype,
                      double cos_sza, double earth_radius, const double* los_cos_vza, const double* los_rel_az,
                      const double* ssa, const double* ext, const double* leg, void* dgeev_ptr, double* out) {
    using namespace oracle;
    try {
        std::vector<double> a(alt, alt + nloc), cz(los_cos_vza, los_cos_vza + nlos), az(los_rel_az, los_rel_az + nlos);
        Plan P = make_plan(nstr, a, interp, geotype, cos_sza, earth_radius, cz, az);
        const int L = nloc - 1, N = nstr / 2;
#pragma omp parallel for schedule(dynamic, 1)
        for (int w = 0; w < nwavel; ++w) {
            WavelInputs in{};
            in.ext = ext + size_t(nloc) * w;
            in.ssa = ssa + size_t(nloc) * w;
            in.leg = leg + size_t(nleg) * nloc * w;
            in.nleg = nleg;
            in.f = g_f ? g_f + size_t(nloc) * w : nullptr;
            in.solar = 1.0;
            in.d_leg = nullptr;
            in.ngroups = 0;
            Solver<double> S(P, (dgeev_fn)dgeev_ptr);
            Layers<double> Ly;
            S.layer_optics(in, Ly);
            LayerSolution<double> sol;
            for (int m = 0; m < nstr; ++m)
                for (int p = 0; p < L; ++p) {
                    S.homogeneous(m, Ly.ssa[p], Ly.beta[p], sol);
                    double d1 = 1e300, d2 = 1e300;
                    for (int j = 0; j < N; ++j) {
                        d1 = std::min(d1, std::abs(Ly.secant[p] - sol.k[j]));
                        for (int l = 0; l < nlos; ++l) d2 = std::min(d2, std::abs(1.0 - cz[l] * sol.k[j]));
                    }
                    double* o = out + ((size_t(w) * nstr + m) * L + p) * 2;
                    o[0] = d1;
                    o[1] = d2;
                }
        }
        return 0;
    } catch (const std::exception& e) {
        g_last_err = e.what();
        return -3;
    }
}

// Dedicated two-stream source (twostream_oracle.hpp).  Same array layouts as oracle_do_radiance; leg needs >= 2 moments;
// the delta-M fraction set by oracle_set_delta_m enters b1 = leg[1] - 3 f / (1 - f) (cpp_twostream_source.cpp:879-897).
int oracle_twostream_radiance(int nloc, int nwavel, int nleg, int nlos, const double* alt, int interp, int geotype,
                              double cos_sza, double earth_radius, const double* los_cos_vza, const double* los_rel_az,
                              const double* ssa, const double* ext, const double* leg, const double* solar,
                              const double* albedo, int nthreads, double* radiance) {
    using namespace oracle;
    try {
        if (nleg < 2) throw std::runtime_error("two-stream source needs at least two phase moments");
        std::vector<double> a(alt, alt + nloc), cz(los_cos_vza, los_cos_vza + nlos), az(los_rel_az, los_rel_az + nlos);
        Plan P = make_plan(2, a, interp, geotype, cos_sza, earth_radius, cz, az);
#ifdef _OPENMP
        if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(static)
        for (int w = 0; w < nwavel; ++w) {
            std::vector<double> b1(nloc);
            for (int q = 0; q < nloc; ++q) {
                const double f = g_f ? g_f[q + size_t(nloc) * w] : 0.0;
                b1[q] = leg[1 + size_t(nleg) * (q + size_t(nloc) * w)] - 3.0 * f / (1.0 - f);
            }
            twostream::solve_wavelength(P, ext + size_t(nloc) * w, ssa + size_t(nloc) * w, b1.data(), solar[w], albedo[w],
                                        radiance + size_t(w) * nlos);
        }
        return 0;
    } catch (const std::exception& e) {
        g_last_err = e.what();
        return -3;
    }
}

// 1: tangent-layer lengths without the reference's rounding-dependent 0.1 m error (limb_oracle.hpp, exact_tangent_ref)
void oracle_set_exact_tangent(int on) { oracle::limb::exact_tangent_ref() = on; }

// Spherical line-of-sight path (limb_oracle.hpp).  rays: [nrays][5] = kind (0 GroundViewingSolar(cos_sza, rel_az, cos_vza,
// observer_altitude), 1 TangentAltitudeSolar(tangent_altitude, rel_az, observer_altitude, cos_sza)) followed by the four
// constructor arguments.  ms_do: multiple_scatter_source = DiscreteOrdinates (interpolated DO source table);
// ss_exact: single_scatter_source = Exact.  radiance, los_od: [nwavel, nrays] C-order.
int oracle_limb_radiance(int nstr, int nloc, int nwavel, int nleg, int nrays, const double* alt, int interp, double cos_sza,
                         double saa, double earth_radius, const double* rays, int num_sza, int ms_do, int ss_exact,
                         int num_ss_moments, const double* ssa, const double* ext, const double* leg, const double* solar,
                         const double* albedo, oracle::dgeev_fn dgeev, int nthreads, double* radiance, double* los_od) {
    using namespace oracle;
    try {
        std::vector<double> a(alt, alt + nloc);
        std::vector<limb::RaySpec> specs(nrays);
        for (int i = 0; i < nrays; ++i) {
            specs[i].kind = (int)rays[i * 5];
            for (int k = 0; k < 4; ++k) specs[i].p[k] = rays[i * 5 + 1 + k];
        }
        limb::LimbGeometry G(nstr, a, interp, cos_sza, saa, earth_radius, specs, num_sza);
        limb::LimbConfig cfg;
        cfg.ms_do = ms_do != 0;
        cfg.ss_exact = ss_exact != 0;
        cfg.num_ss_moments = num_ss_moments;
        limb::LimbSolver S(G, cfg, dgeev);
        std::string err;
#ifdef _OPENMP
        if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic)
        for (int w = 0; w < nwavel; ++w) {
            try {
                WavelInputs in{};
                in.ext = ext + size_t(nloc) * w;
                in.ssa = ssa + size_t(nloc) * w;
                in.leg = leg + size_t(nleg) * nloc * w;
                in.nleg = nleg;
                in.f = nullptr;
                in.solar = solar[w];
                in.albedo = albedo[w];
                in.d_leg = nullptr;
                in.ngroups = 0;
                in.include_ss = false;
                in.num_azimuth = nstr;
                S.solve_wavelength(in, radiance + size_t(w) * nrays, los_od ? los_od + size_t(w) * nrays : nullptr);
            } catch (const std::exception& e) {
#pragma omp critical
                err = e.what();
            }
        }
        if (!err.empty()) throw std::runtime_error(err);
        return 0;
    } catch (const std::exception& e) {
        g_last_err = e.what();
        return -3;
    }
}

// Geometry export of the spherical path for checking the product's host-side ray tracer: per ray the number of layers,
// ground flag, and per layer [layer_distance, od_quad_start, od_quad_end, cos_sza_entrance, cos_sza_exit, saz_entrance,
// saz_exit, r_entrance, r_exit] (9 doubles, at most max_layers per ray); cos_scatter [nrays]: cosine of the
// single-scattering angle of each (straight) ray.
int oracle_limb_geometry(int nloc, int nrays, const double* alt, int interp, double cos_sza, double saa, double earth_radius,
                         const double* rays, int max_layers, int* nlayers, int* ground_hit, double* layer_data,
                         double* cos_scatter) {
    using namespace oracle;
    try {
        std::vector<double> a(alt, alt + nloc);
        std::vector<limb::RaySpec> specs(nrays);
        for (int i = 0; i < nrays; ++i) {
            specs[i].kind = (int)rays[i * 5];
            for (int k = 0; k < 4; ++k) specs[i].p[k] = rays[i * 5 + 1 + k];
        }
        limb::LimbGeometry G(2, a, interp, cos_sza, saa, earth_radius, specs, 1);
        for (int i = 0; i < nrays; ++i) {
            const auto& r = G.rays[i];
            nlayers[i] = (int)r.layers.size();
            ground_hit[i] = r.ground_is_hit ? 1 : 0;
            cos_scatter[i] = G.cos_scatter[i];
            for (int j = 0; j < (int)r.layers.size() && j < max_layers; ++j) {
                const auto& l = r.layers[j];
                double* o = layer_data + (size_t(i) * max_layers + j) * 9;
                o[0] = l.layer_distance; o[1] = l.od_quad_start; o[2] = l.od_quad_end;
                o[3] = l.cos_sza_entrance; o[4] = l.cos_sza_exit; o[5] = l.saz_entrance; o[6] = l.saz_exit;
                o[7] = l.r_entrance; o[8] = l.r_exit;
            }
        }
        return 0;
    } catch (const std::exception& e) {
        g_last_err = e.what();
        return -3;
    }
}

// Geometry plan export, for checking the product's host-side tables against the oracle's.
int oracle_plan(int nstr, int nloc, int nlos, const double* alt, int interp, int geotype, double cos_sza,
                double earth_radius, const double* los_cos_vza, const double* los_rel_az, double* mu, double* wt,
                double* lp_mu, double* lp_csz, double* lp_los, double* W, double* chapman) {
    using namespace oracle;
    try {
        std::vector<double> a(alt, alt + nloc), cz(los_cos_vza, los_cos_vza + nlos), az(los_rel_az, los_rel_az + nlos);
        Plan P = make_plan(nstr, a, interp, geotype, cos_sza, earth_radius, cz, az);
        std::copy(P.mu.begin(), P.mu.end(), mu);
        std::copy(P.wt.begin(), P.wt.end(), wt);
        std::copy(P.lp_mu.begin(), P.lp_mu.end(), lp_mu);
        std::copy(P.lp_csz.begin(), P.lp_csz.end(), lp_csz);
        std::copy(P.lp_los.begin(), P.lp_los.end(), lp_los);
        std::copy(P.W.begin(), P.W.end(), W);
        std::copy(P.chapman.begin(), P.chapman.end(), chapman);
        return 0;
    } catch (const std::exception& e) {
        g_last_err = e.what();
        return -3;
    }
}

// Band solver check (restates the generator idea of the reference's test_band_factorization.cpp:64-227):
// solves A x = b (trans = 0) or A^T x = b (trans = 1) for a dense row-major A with bandwidths kl = ku.
int oracle_band_solve(int n, int kl, const double* dense_a, double* b, int trans) {
    oracle::BandLU lu;
    lu.init(n, kl, kl);
    for (int i = 0; i < n; ++i)
        for (int j = std::max(0, i - kl); j <= std::min(n - 1, i + kl); ++j) lu.at(i, j) = dense_a[size_t(i) * n + j];
    int info = lu.factor();
    if (info != 0) return info;
    if (trans)
        lu.solve_transposed(b);
    else
        lu.solve(b);
    return 0;
}
}
