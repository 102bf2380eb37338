// Device-side view of the spherical ("limb") line-of-sight path: geometry tables of disco_limb.h uploaded once per
// engine, plus the per-chunk arrays of the source table.  See disco_limb.cu for the kernels.
#pragma once
#include <cuda_runtime.h>

#include "disco_bodies.h"
#include "disco_limb.h"

namespace disco {

struct LimbView {
    int nrays, nsza, npts, nseg, nss;
    int ms_do, ss_exact;
    // geometry (device pointers, layouts as in LimbPlan)
    const double* layer_fraction;
    const double* lp_ang;
    const int *pt_angle, *pt_alt, *pt_sza;
    const int* seg_start;
    const int* od_idx;
    const double *od_w, *ent_w, *exit_w;
    const int* mid_idx;
    const double* mid_w;
    const double *seg_len, *seg_qfrac;
    const int* seg_lower;
    const int* src_pt;
    const double *src_w, *src_cos;
    const int *gnd_hit, *gnd_sza_idx;
    const double *gnd_sza_w, *gnd_mu_in, *wig_ss;
    const int *sol_start, *sol_idx;
    const double* sol_w;
    const int* sol_blocked;
    const int* ray_order;   // rays sorted by decreasing number of segments (block scheduling order of the integrator)
    // per-chunk arrays, wavelength fastest
    double* coef;      // [nsza][L][M][nw][nstr]  Legendre projection of the diffuse field at the sampled altitude
    double* ground;    // [nsza][nw]              order-0 Lambertian ground source
    double* table;     // [npts][M][nw]           source table at the needed (cos zenith, altitude, SZA) points
    double* phase;     // [nrays][nw][nloc]       single-scatter phase function of every grid point at the ray's angle
    double* radiance;  // [nw][nrays]
    double* los_od;    // [nw][nrays] or null
};

void launch_limb_coef(const ChunkView& V, const LimbView& Lv, int sza_index, cudaStream_t st);
void launch_limb_table(const ChunkView& V, const LimbView& Lv, cudaStream_t st);
void launch_limb_phase(const ChunkView& V, const LimbView& Lv, cudaStream_t st);
void launch_limb_integrate(const ChunkView& V, const LimbView& Lv, cudaStream_t st);

}  // namespace disco
