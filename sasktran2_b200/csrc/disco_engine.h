// Host orchestration of the CUDA discrete-ordinates solve: device memory, wavelength chunking, streams,
// CUDA-event timing.  Plain CUDA runtime, no framework dependencies.
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "disco_comm.h"
#include "disco_kernels.cuh"
#include "disco_brdf.h"
#include "disco_limb.cuh"
#include "disco_plan.h"

namespace disco {

struct EngineOptions {
    int nstr = 16;
    bool include_ss = true;     // single_scatter_source == discrete_ordinates
    int forced_azimuth = -1;    // num_do_forced_azimuth (<= 0: all nstr orders)
    bool twostream = false;     // multiple_scatter_source == TwoStream: dedicated closed-form kernel when no weighting functions are asked
    double workspace_gb = -1.0; // per-chunk workspace budget; < 0: min(32 GB, a quarter of the free device memory)
    int device = -1;            // -1: current device
    bool validate_inputs = true; // config.input_validation_mode != disabled: device pre-pass over extinction / SSA
};

// Caller-owned host arrays in the reference's C-ABI layouts (cpp/include/c_api/atmosphere.h:86-91)
struct AtmosphereArrays {
    int nloc = 0, nwavel = 0, nleg = 0;
    const double* ssa = nullptr;    // [nloc, nwavel]
    const double* ext = nullptr;    // [nloc, nwavel]
    const double* leg = nullptr;    // [nleg, nloc, nwavel]
    const double* solar = nullptr;  // [nwavel]
    const double* albedo = nullptr; // [nwavel]
    const double* f = nullptr;      // [nloc, nwavel] delta-M truncation fraction, null when no scaling was applied
    // surface BRDF (cpp/include/c_api/brdf.h): 0 Lambertian (`albedo`), 2 MODIS with brdf_args [3, nwavel] column-major
    int brdf_kind = 0, brdf_nargs = 1;
    const double* brdf_args = nullptr;
    // thermal emission solved by the discrete-ordinates path (config.emission_source = discrete_ordinates):
    // emission_source at the grid points [nloc, nwavel] (null: none) and the surface emission [nwavel] (null: none)
    const double* emission = nullptr;
    const double* surface_emission = nullptr;
};

// Weighting-function request: which derivative mappings to evaluate and where the results go
// (OutputC::assign_lane, cpp/lib/output/outputc.cpp:37-160).  All pointers are caller-owned host memory.
struct WfMapping {
    const double* d_ssa = nullptr;        // [nloc, nwavel]
    const double* d_extinction = nullptr; // [nloc, nwavel]
    const double* scat_factor = nullptr;  // [nloc, nwavel] or null (absorber)
    int scat_index = -1;                  // scattering group of this mapping
    const double* interpolator = nullptr; // [nloc, nout] column-major or null
    int nout = 0;
    bool log_radiance_space = false;
    double* out = nullptr;                // [nout][nwavel][nlos]
};
struct WfSurface {
    const double* d_brdf = nullptr;       // [nwavel, nargs] column-major (Lambertian: nargs = 1)
    int nargs = 1;
    size_t nwavel = 0;                    // column stride of d_brdf
    double* out = nullptr;                // [nwavel][nlos]
};
struct WfRequest {
    std::vector<const double*> d_legendre;  // per scattering group: [nleg, nloc, nwavel]
    std::vector<const double*> d_f;         // per scattering group: [nloc, nwavel] d(delta-M fraction), empty: unscaled
    std::vector<WfMapping> mappings;
    std::vector<WfSurface> surfaces;
    bool enabled() const { return !mappings.empty() || !surfaces.empty(); }
};

enum TimingSlot { T_H2D = 0, T_OPTICS, T_LAYER, T_BVP, T_RADIANCE, T_D2H, T_TOTAL_KERNELS, T_WF, T_WF_ADJOINT, T_WF_LAYER, T_WF_CHAIN, T_WF_MAP, T_LIMB_SOURCE, T_LIMB_INTEGRATE, T_NSLOTS };

class DeviceEngine {
  public:
    DeviceEngine(const EngineOptions& opt, const HostPlan& plan);
    // Spherical line-of-sight path: `plan` carries the stream tables and no lines of sight, `limb` the traced rays, the
    // DO source-table layout and one (cos SZA, chapman) set per SZA of the DO grid.  Radiance columns = limb.nrays.
    DeviceEngine(const EngineOptions& opt, const HostPlan& plan, const LimbPlan& limb);
    ~DeviceEngine();
    DeviceEngine(const DeviceEngine&) = delete;
    DeviceEngine& operator=(const DeviceEngine&) = delete;

    // Copy wavelengths [w0, w0+nw) of the atmosphere to the device and keep them resident.
    void stage(const AtmosphereArrays& atm, int w0, int nw, const WfRequest* wf = nullptr);
    // Run the kernels on the staged wavelengths; results stay on the device.
    void solve_staged();
    // Copy radiance [nw, nlos] of the staged range back to the host.
    void fetch(double* radiance_host);
    // limb path: line-of-sight optical depths [nw, nrays] of the staged range (after solve_staged)
    void fetch_los_optical_depth(double* host);
    bool limb() const { return m_is_limb; }
    int radiance_columns() const { return m_nrad; }
    // stage + solve + fetch
    void calculate(const AtmosphereArrays& atm, int w0, int nw, double* radiance_host, const WfRequest* wf = nullptr);
    // Wavelength-sharded solve: after solve_staged() on every rank, collect the results of all ranks on `root` over
    // NCCL and write them into the root's full-spectrum host arrays (radiance [nw_total][nlos]; one
    // [nout][nw_total][nlos] array per mapping in the order of the staged request; one [nw_total][nlos] per surface
    // mapping).  block_start / block_count: the wavelength block of every rank.  Host pointers are read on `root` only.
    // Returns the milliseconds spent in (NCCL exchange, device -> host copies) on this rank.
    void gather_to_root(Comm& comm, int root, const int* block_start, const int* block_count, int nw_total,
                        double* radiance_host, double* const* mapping_host, double* const* surface_host, double ms_out[2]);
    bool wf_active() const { return m_wf_on; }
    bool fast_path() const { return m_fast; }
    bool twostream_direct() const { return m_opt.twostream && !m_wf_on && twostream_supported(m_plan.L, m_plan.plane_parallel); }
    bool twostream_direct_for(bool wf_requested) const { return m_opt.twostream && !wf_requested && twostream_supported(m_plan.L, m_plan.plane_parallel); }
    // test/debug: copy a workspace array of the LAST chunk to the host; returns the number of doubles copied
    size_t debug_copy(const char* name, double* host, size_t max_n);
    std::vector<std::pair<std::string, std::pair<double*, size_t>>> m_dbg;

    const double* timings_ms() const { return m_ms; }   // accumulated over the last solve / calculate
    long long kernel_launches() const { return m_launches; }
    int staged_wavelengths() const { return m_nw; }
    int num_azimuth_solved() const { return (int)m_mlist.size(); }
    size_t workspace_bytes_per_wavelength() const;
    int chunk_wavelengths() const;
    const HostPlan& plan() const { return m_plan; }
    void set_workspace_gb(double gb) { m_opt.workspace_gb = gb; }

  private:
    void init(const EngineOptions& opt);
    void init_limb();
    // spherical path: one eigen-solve and one BVP factorisation for all SZAs of the source table (SK_B200_LIMB_SHARED=0: per SZA)
    bool limb_shared_factorisation() const;
    void free_inputs();
    void free_workspace();
    void ensure_workspace(int chunk);
    template <class T> T* dalloc(size_t n);

    EngineOptions m_opt;
    HostPlan m_plan;
    cudaStream_t m_stream = nullptr;
    // copy/compute overlap inside calculate(): inputs beyond the first chunk arrive on m_copy while the first chunk is
    // solved, outputs of all chunks but the last leave on m_copy while the last chunk is solved
    cudaStream_t m_copy = nullptr;
    cudaEvent_t m_ev_h2d_tail = nullptr, m_ev_out_ready = nullptr;
    bool m_overlap = false, m_tail_pending = false;
    int m_h2d_head = 0;             // staged wavelengths that the compute stream may touch without waiting for m_copy
    int m_early_done = 0;           // wavelengths whose outputs are already on their way to the host
    double* m_early_radiance = nullptr;
    int planned_chunk(int nw, bool wf_on, int ngroups) const;
    size_t ws_bytes(bool wf_on, int ngroups) const;
    void copy_outputs(int w_begin, int w_end, double* radiance_host, cudaStream_t s);
    cudaEvent_t m_ev[8] = {};
    std::vector<cudaEvent_t> m_marks;   // per-kernel timing events, created once and reused by every solve
    int m_device = 0;                   // every public method selects it first (engines on several GPUs per process)
    // geometry tables on device
    double *d_mu = nullptr, *d_wt = nullptr, *d_lp_mu = nullptr, *d_lp_csz = nullptr, *d_lp_los = nullptr;
    double* d_wf_tab = nullptr;  // per-order tables of the fast weighting-function kernel (Tables::wf_tab)
    double *d_los_mu = nullptr, *d_los_cosmphi = nullptr, *d_layer_dh = nullptr, *d_interp_w = nullptr, *d_chapman = nullptr;
    int *d_interp_idx = nullptr, *d_mlist = nullptr;
    std::vector<int> m_mlist;
    // staged inputs
    int m_nw = 0, m_nleg = 0, m_cap_nw = 0, m_cap_nleg = 0;
    double *d_ext = nullptr, *d_ssa = nullptr, *d_leg = nullptr, *d_solar = nullptr, *d_albedo = nullptr;
    unsigned char* d_los_zero = nullptr;   // Tables::los_zero
    double *d_emission = nullptr, *d_semis = nullptr;   // thermal sources of the staged range (allocated on first use)
    int m_cap_emission = 0, m_cap_semis = 0;
    bool m_emission_on = false, m_semis_on = false;
    double* d_radiance = nullptr;
    unsigned int* d_status = nullptr;
    // weighting functions
    struct DevMapping {
        double *d_ssa = nullptr, *d_ext = nullptr, *scat = nullptr, *interp = nullptr, *out = nullptr;
        WfMapping host;
    };
    struct DevSurface {
        double *d_brdf = nullptr, *out = nullptr;
        WfSurface host;
    };
    void free_wf_inputs();
    bool m_wf_on = false;
    int m_ngroups = 0, m_w0 = 0, m_nw_total = 0, m_wf_nw = 0, m_wf_nleg = 0;
    double* d_dleg = nullptr;
    double* d_fdm = nullptr;     // delta-M inputs of the staged range: f | d_f per group, [1 + G][nloc, nw]
    size_t m_cap_fdm = 0;
    bool m_has_f = false;
    std::vector<DevMapping> m_maps;
    std::vector<DevSurface> m_surfs;
    bool m_ws_wf = false;
    int m_ws_ngroups = 0;     // scattering groups the workspace was sized for (lay_dbeta, wf_loc, wf_native)
    bool m_fast = false;      // register-resident layer solve (disco_fast*.cuh)
    // staging of the other ranks' results on the gather root
    double* d_gather = nullptr;
    size_t m_cap_gather = 0;
    // chunk workspace
    int m_ws_chunk = 0;
    std::vector<void*> m_ws_ptrs;
    ChunkView m_view{};
    double m_ms[T_NSLOTS] = {};
    long long m_launches = 0;
    // ---- spherical line-of-sight path
    bool m_is_limb = false;
    int m_nrad = 0;                 // radiance columns: DO lines of sight, or traced rays of the limb path
    LimbPlan m_limb;
    LimbView m_lview{};
    std::vector<void*> m_limb_ptrs; // geometry tables on the device
    std::vector<double*> d_sza_lp_csz, d_sza_chapman;   // per SZA of the DO grid
    double* d_los_od = nullptr;     // [nw][nrays] of the staged range
    // ---- kernel-based surface BRDF
    int m_brdf_kind = 0, m_brdf_tab_kind = 0, m_brdf_nargs = 1;
    double *d_brdf_args = nullptr, *d_zero_albedo = nullptr;
    size_t m_cap_brdf = 0;
    double *d_brdf_Rss = nullptr, *d_brdf_rsun = nullptr, *d_brdf_Rls = nullptr, *d_brdf_rlsun = nullptr;
    int m_brdf_nk = 0;
    bool m_ws_brdf = false;
    int m_ws_brdf_kind = 0;
    // snow model: sample tables (device) and the per-chunk coefficient array
    double *d_snow_r0 = nullptr, *d_snow_g = nullptr, *d_snow_cos = nullptr, *d_snow_w = nullptr, *d_snow_scale = nullptr;
    int m_snow_npairs = 0, m_snow_nsamples = 0;
    double* d_brdf_pw = nullptr;   // workspace: [chunk][M][npairs]
};

}  // namespace disco
