// CUDA kernels of the batched discrete-ordinates radiance solve (sm_100a, fp64).
// Batched over wavelength x azimuth order x layer; see DESIGN.md for the data layout and rooflines.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "disco_bodies.h"

namespace disco {

// cudaFuncSetAttribute is per device: one flag per (call site, device) instead of a process-wide bool
struct DeviceOnce {
    unsigned long long seen = 0;
    bool first() {
        int d = 0;
        cudaGetDevice(&d);
        const unsigned long long bit = 1ull << (d & 63);
        if (seen & bit) return false;
        seen |= bit;
        return true;
    }
};

// kernel-based BRDF (disco_brdf.h): device tables of the kernels' Fourier coefficients + this chunk's arguments
void launch_surface_general(const ChunkView& V, const BrdfView& B, cudaStream_t s);
void launch_brdf_expand_snow(const ChunkView& V, const BrdfView& B, cudaStream_t s);
void launch_layer_optics(const ChunkView& V, cudaStream_t s);
void launch_beam(const ChunkView& V, cudaStream_t s, bool scan_od = true);
void launch_validate_inputs(const ChunkView& V, cudaStream_t s);
void launch_layer_solve(const ChunkView& V, cudaStream_t s);
// register-resident path for N = 2, 4, 8 (disco_fast*.cuh); needs the eig*/los_* planes and vsrc_w = N
bool fast_path_supported(int N);
void launch_layer_solve_fast(const ChunkView& V, cudaStream_t s);
// the two halves of the above: homogeneous solutions (independent of the solar geometry), particular solutions + LOS terms
void launch_layer_eig_fast(const ChunkView& V, cudaStream_t s);
void launch_layer_post_fast(const ChunkView& V, cudaStream_t s);
// forward BVP of V.nsza solar geometries with one factorisation (2 <= nsza <= 4, N <= 8)
bool bvp_multi_supported(int N, int nsza);
void launch_bvp_multi(const ChunkView& V, cudaStream_t s);
void launch_wf_layer_fast(const ChunkView& V, cudaStream_t s);
int wf_layer_fast_tile(int N, int G, int nlos);   // lines of sight per tile of k_wf_layer_fast, 0: not applicable
void launch_bvp(const ChunkView& V, cudaStream_t s);
void launch_bvp_adjoint(const ChunkView& V, cudaStream_t s);
struct MappingView;
void launch_wf_layer(const ChunkView& V, cudaStream_t s);
void launch_wf_chain(const ChunkView& V, cudaStream_t s);
void launch_wf_map(const ChunkView& V, const MappingView& Mp, int w0, int nw_total, bool log_space, cudaStream_t s);
void launch_wf_ground_reduce(const ChunkView& V, cudaStream_t s);   // per-order ground pieces -> wf_gnd / wf_gndk
void launch_wf_surface_args(const ChunkView& V, const double* d_brdf, size_t arg_stride, double* out, int w0, cudaStream_t s);
void launch_wf_surface(const ChunkView& V, const double* d_brdf, double* out, int w0, cudaStream_t s);
int adjoint_groups_per_problem(int nlos);
int adjoint_max_rhs(int nlos);
size_t bvp_fac_stride(int N, int nrhs, int L);
// Transposed solves from the forward factors (k_bvp_tsolve, disco_bvp.cuh): lanes per problem, problems per warp
// that fit the shared-memory budget, whether the engine uses it for (N, nlos), doubles of multiplier storage per problem
inline int tsolve_lanes(int nlos) { return nlos < 32 ? nlos : 32; }
inline int tsolve_groups_per_warp(int N, int glt) {
    const int fs = ((4 * N + 1 + 1) & ~1) + 2;
    const int per_group = (2 * (2 * N) * fs + 4) * (int)sizeof(double);  // two ring slots of one factor block + skew
    const int fit = (56 * 1024) / (2 * per_group);                 // two warps per block, four blocks per SM
    const int g = 32 / glt;
    return g < fit ? g : (fit < 1 ? 1 : fit);
}
bool adjoint_reuses_factors(int N, int nlos);
size_t bvp_lfac_stride(int N, int L);
void launch_radiance(const ChunkView& V, cudaStream_t s);
// dedicated two-stream source (disco_twostream.cu): radiances of a chunk straight from the staged inputs
bool twostream_supported(int L, bool plane_parallel);
void launch_twostream(const ChunkView& V, cudaStream_t s);
bool nstr_supported(int nstr);
double measure_fp64_tflops();

}  // namespace disco
