// Spherical ("limb") line-of-sight path: host-side geometry plan (everything that does not depend on wavelength) for the
// CUDA kernels of disco_limb.cu.  What it replaces in the reference (paths relative to cpp/):
//   ray construction          lib/viewinggeometry/{groundviewing,tangentaltitudesolar}.cpp, lib/geometry/geometry.cpp:155-232
//   spherical shell tracer    lib/raytracing/spherical_shell.cpp:6-262 (straight rays, observer outside the atmosphere;
//                             solar rays: observer inside looking up), include/sasktran2/raytracing.h:319-560
//   optical-depth stencils    include/sasktran2/raytracing.h:590-625 (construct_od_matrix), lib/sourceintegrator/sourceintegrator.cpp:11-33
//   DO source table layout    lib/sktran_disco/source_term/do_source.cpp:61-159, do_source_diffuse_storage.cpp:8-267, 415-433
//   exact solar geometry      lib/solar/solartransmissionexact.cpp:36-96, lib/phasefunction/phasehandler.cpp:241-262
// How it is organised is its own: one flat, device-ready table per quantity (CSR where ragged), segments of all rays
// concatenated, the (cos zenith, altitude, SZA) source points that any ray touches compacted into one list.
#pragma once
#include <vector>

#include "disco_plan.h"

namespace disco {

struct LimbRay {
    int kind;      // 0 GroundViewingSolar(cos_sza, rel_az, cos_vza, observer_altitude); 1 TangentAltitudeSolar(tangent_altitude, rel_az, observer_altitude, cos_sza)
    double p[4];   // in the reference constructors' argument order
};

struct LimbOptions {
    int num_sza = 1;          // config.num_do_sza
    bool ms_do = true;        // multiple_scatter_source == discrete_ordinates
    bool ss_exact = false;    // single_scatter_source == exact
    int num_ss_moments = 16;  // config.num_singlescatter_moments
};

constexpr int kLimbAngles = 100;       // cos-zenith grid of the source table, linspace(-1, 1, 100) (do_source_diffuse_storage.cpp:29)
constexpr int kLimbSrcEntries = 8;     // 2 (SZA) x 2 (altitude) x 2 (cos zenith) interpolation corners per segment
constexpr int kLimbStencil = 4;        // grid points a traced layer can touch (raytracing.h:412)

struct LimbPlan {
    int nstr = 0, nrays = 0, nsza = 0, nalt = 0, nseg = 0, npts = 0, nbnd = 0, nss = 0;
    bool ms_do = true, ss_exact = false;
    std::vector<HostPlan> sza_plans;      // per SZA: lp_csz, chapman, csz (pseudo-spherical construction, no lines of sight)
    std::vector<double> sza_grid;         // [nsza] cos(SZA)
    std::vector<double> layer_fraction;   // [L] by layer index p (0 = top): relative optical-depth position of the sampled altitude
    std::vector<double> lp_ang;           // [angle][m][l] d^l_m0 at the table's cos-zenith grid (LegendrePhaseStorage::fill)
    // needed source points (compacted m_need_to_calculate_map): angle index, altitude (layer mid-point) index, SZA index
    std::vector<int> pt_angle, pt_alt, pt_sza;
    // ---- segments ("layers" of the traced rays); seg_start[r] .. seg_start[r+1]: ray r, first = farthest from the observer
    std::vector<int> seg_start;           // [nrays + 1]
    std::vector<int> od_idx;              // [nseg][4] shared stencil of the segment (-1 padded -> weight 0, index 0)
    std::vector<double> od_w, ent_w, exit_w;   // [nseg][4] optical-depth / entrance / exit weights on that stencil
    std::vector<int> mid_idx;             // [nseg][2] SSA interpolation at the segment mid-point
    std::vector<double> mid_w;            // [nseg][2]
    std::vector<double> seg_len;          // [nseg] layer_distance
    std::vector<double> seg_qfrac;        // [nseg][2] od_quad_start_fraction, od_quad_end_fraction
    std::vector<int> seg_lower;           // [nseg] lower-interpolation endpoint swap: 0 none, 1 both ends use the entrance weights, 2 both the exit weights
    std::vector<int> src_pt;              // [nseg][8] index into the point list (-1: unused)
    std::vector<double> src_w;            // [nseg][8] interpolation weight (altitude x angle x SZA)
    std::vector<double> src_cos;          // [nseg][nstr] cos(m * azimuth) of the segment
    // ---- per ray
    std::vector<int> gnd_hit;             // [nrays]
    std::vector<int> gnd_sza_idx;         // [nrays][2]
    std::vector<double> gnd_sza_w;        // [nrays][2] (order-0 ground source: the cos-zenith weights sum to one)
    std::vector<double> gnd_mu_in;        // [nrays] cos(SZA) at the ground point (exact single scatter), <= 0: no direct term
    std::vector<double> wig_ss;           // [nrays][nss] d^l_00 at the ray's single-scattering angle
    // ---- solar rays of the exact single-scatter source: boundary b of ray r is row seg_start[r] + r + b
    std::vector<int> sol_start;           // [nbnd + 1] CSR
    std::vector<int> sol_idx;
    std::vector<double> sol_w;
    std::vector<int> sol_blocked;         // [nbnd] the solar ray hits the ground
    std::vector<int> ray_order;           // [nrays] rays by decreasing number of segments
    // diagnostics (tests): per-ray geometric summaries
    std::vector<double> ray_cos_scatter;  // [nrays]
};

// Throws std::runtime_error on unsupported input (observer inside the atmosphere, sun below the horizon of a source point, ...).
LimbPlan build_limb_plan(int nstr, const GeometrySpec& geo, const std::vector<LimbRay>& rays, const LimbOptions& opt);

}  // namespace disco
