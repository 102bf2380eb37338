// Host-side geometry plan of the DO solve: everything that does not depend on wavelength.
// Replaces, for the scalar plane-parallel / pseudo-spherical path, what the reference keeps in
//   SKTRAN_DO_UserSpec::cacheLPOfStreamAngles   cpp/lib/sktran_disco/sktran_do_specs.cpp:58-103
//   PersistentConfiguration::configure          cpp/lib/sktran_disco/sktran_do_pconfig.cpp:8-47
//   GeometryLayerArray                          cpp/lib/sktran_disco/sktran_do_geometrylayerarray.cpp:8-119
//   DOSourcePlaneParallelPostProcessing::initialize_geometry   source_term/do_source_planeparallel.cpp:599-661
#pragma once
#include <string>
#include <vector>

namespace disco {

struct LineOfSight {
    double cos_vza;
    double rel_azimuth;
    double observer_altitude;
};

struct GeometrySpec {
    std::vector<double> altitudes;  // ascending grid, metres
    int interp = 1;                 // 0 shell, 1 linear, 2 lower   (cpp/include/c_api/geometry.h:10-13)
    int geotype = 0;                // 0 plane-parallel, 1 pseudo-spherical, 2 spherical, 3 ellipsoidal
    double cos_sza = 1.0;
    double saa = 0.0;
    double earth_radius = 6372000.0;
};

struct HostPlan {
    int nstr = 0, N = 0, L = 0, nloc = 0, nlos = 0;
    int interp = 1;                    // the geometry's interpolation method (GeometrySpec::interp)
    double csz = 0.0;
    std::vector<double> mu, wt;        // [nstr]; first N are mu > 0
    std::vector<double> lp_mu;         // [m][i<N][l]
    std::vector<double> lp_csz;        // [m][l]
    std::vector<double> lp_los;        // [los][m][l]
    std::vector<double> los_mu;        // [nlos]
    std::vector<double> los_cosmphi;   // [nlos][m]
    std::vector<double> layer_dh;      // [L] ceiling - floor
    std::vector<int> interp_idx;       // [L][2] contributing grid points (-1 = none)
    std::vector<double> interp_w;      // [L][2]
    std::vector<double> chapman;       // [L][L] row p, col q (0 above the diagonal)
    bool plane_parallel = true;
};

// Gauss-Legendre rule of order n on (-1, 1), ascending nodes
void gauss_rule(int n, std::vector<double>& nodes, std::vector<double>& weights);
// Wigner function d^l_{m0}(acos coszen) (cpp/include/sasktran2/math/wigner.h:56-149)
double wigner_dm0(int m, int l, double coszen);

// Throws std::runtime_error with a reference-style message on unsupported / invalid input.
HostPlan build_plan(int nstr, const GeometrySpec& geo, const std::vector<LineOfSight>& los);

}  // namespace disco
