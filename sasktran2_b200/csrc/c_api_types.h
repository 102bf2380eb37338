// Handle types behind the opaque pointers of the C ABI (include/sasktran2_b200.h), shared by c_api.cpp (the
// discrete-ordinates path) and c_api_extra.cpp (the rest of the reference's cpp/include/c_api surface).
#pragma once
#include "../../include/sasktran2_b200.h"

#include <map>
#include <memory>
#include <mutex>
#include <set>
#include <string>
#include <vector>

#include "disco_engine.h"

namespace skapi {
// records the message for sk_b200_last_error(), logs it according to the log level, returns `code`
int fail(int code, const std::string& msg);
void* host_alloc(size_t nbytes);
void host_free(void* p);
}  // namespace skapi

// ---------------------------------------------------------------------------------------------------
// handle types
// ---------------------------------------------------------------------------------------------------
struct Config {
    int num_stokes = 1;
    int multiple_scatter_source = 3;  // none
    int single_scatter_source = 0;    // exact
    int num_streams = 16;
    int num_threads = 1;
    int threading_model = 0;
    int wavelength_batch_size = 1;
    int num_singlescatter_moments = 16;
    int apply_delta_scaling = 0;
    int num_do_sza = 1;
    int num_do_forced_azimuth = -1;
    int do_backprop = 0;
    int emission_source = 1;     // none
    int occultation_source = 1;  // none
    int solar_refraction = 0;
    int wf_enabled = 1;
    int wf_precision = 0;
    int input_validation_mode = 0;
    int log_level = 3;
    // the rest of cpp/include/c_api/config.h (defaults cpp/lib/config/config.cpp:5-33): stored and returned; only
    // values the CUDA path cannot honour are refused at sk_engine_create
    int singlescatter_phasemode = 0;     // from_legendre
    int num_do_spherical_iterations = 0;
    int num_hr_spherical_iterations = 50;
    int num_hr_incoming = 110, num_hr_outgoing = 110, num_hr_full_incoming_points = -1;
    int initialize_hr_with_do = 0;
    double successive_orders_relative_tolerance = 1e-6, successive_orders_absolute_tolerance = 0.0;
    int successive_orders_anderson_depth = 0;
    double successive_orders_damping = 1.0;
    std::vector<double> successive_orders_altitude_grid_m;
    int los_refraction = 0, multiple_scatter_refraction = 0;
    int stokes_basis = 0;                // standard
    int output_los_optical_depth = 0;
    std::vector<int> flux_types{0, 1};   // upwelling, downwelling
};

struct Geometry1D {
    disco::GeometrySpec spec;
    std::vector<double> refractive_index;  // [nloc], ones (cpp/include/sasktran2/geometry.h: Geometry1D::refractive_index)
};

// Ray kinds of cpp/include/c_api/viewing_geometry.h.  Ground-viewing rays feed the plane-parallel / pseudo-spherical
// post-processing; the others are recorded and refused by sk_engine_create unless the path for them exists.
struct OtherRay {
    int kind;  // 1 tangent_altitude_solar, 2 tangent_altitude (2-D), 3 solar_angles_observer_location
    double a, b, c, d;
};
struct ViewingGeometry {
    std::vector<disco::LimbRay> ordered;   // every solar-angle ray in the order it was added (spherical path)
    std::vector<disco::LineOfSight> rays;
    std::vector<double> ray_cos_sza;
    std::vector<OtherRay> other_rays;
    int num_tangent_rays = 0;
    int num_flux_observers = 0;
};

// Host memory the library owns and copies to / from the device every call (derivative mappings) is page-locked
// when a CUDA device is present, so that the H2D copies run at PCIe speed; without a device it is plain malloc.
template <class T>
struct PinnedAlloc {
    using value_type = T;
    PinnedAlloc() = default;
    template <class U>
    PinnedAlloc(const PinnedAlloc<U>&) {}
    T* allocate(size_t n) { return static_cast<T*>(skapi::host_alloc(n * sizeof(T))); }
    void deallocate(T* p, size_t) { skapi::host_free(p); }
    template <class U>
    bool operator==(const PinnedAlloc<U>&) const { return true; }
    template <class U>
    bool operator!=(const PinnedAlloc<U>&) const { return false; }
};
using PinnedVec = std::vector<double, PinnedAlloc<double>>;

struct MappingImpl {
    int nwavel = 0, nloc = 0, nleg = 0;
    PinnedVec d_ssa, d_extinction, scat_factor, d_legendre;
    bool has_d_ssa = false, has_d_extinction = false, has_legendre = false;
    int scat_deriv_index = -1;
    std::string interp_dim = "altitude", assign_name;
    bool log_radiance_space = false;
    std::vector<double> interpolator;  // column-major [dim1 = nloc, dim2 = nout]
    int interp_d1 = 0, interp_d2 = 0;
    std::vector<double> d_emission;    // [nloc, nwavel]; emission sources are outside the CUDA path
    bool is_scattering() const { return has_legendre; }
    int num_output() const { return interp_d2 > 0 ? interp_d2 : nloc; }
};
struct DerivativeMapping {
    MappingImpl* impl;
};

struct SurfaceMappingImpl {
    int nwavel = 0, nargs = 1;
    std::vector<double> d_brdf;  // [nwavel, nargs] column-major
    std::vector<double> d_emission;  // [nwavel]
    bool has_d_brdf = false;
    std::string interp_dim = "dummy";
    std::vector<double> interpolator;
    int interp_d1 = 0, interp_d2 = 0;
};
struct SurfaceDerivativeMapping {
    SurfaceMappingImpl* impl;
};

struct AtmosphereStorage {
    int nloc = 0, nwavel = 0, nleg = 0, nstokes = 1;
    double *ssa = nullptr, *ext = nullptr, *emission = nullptr, *leg = nullptr, *solar = nullptr;
    std::map<std::string, MappingImpl> mappings;  // name order == the reference's std::map order
    int num_scat_groups = 0;
    // delta-M scaling state (AtmosphereGridStorageFull::f, d_f, applied_f_order, grid_storage.h:40-60)
    int applied_f_order = 0;
    PinnedVec f;                  // [nloc, nwavel] truncation fraction, empty until the scaling is applied
    std::vector<PinnedVec> d_f;   // per scattering group: [nloc, nwavel]
};

struct BRDF {
    int kind = 0;  // 0 lambertian
    int nstokes = 1;
};

struct Surface {
    int nwavel = 0, nstokes = 1;
    double* emission = nullptr;
    BRDF* brdf = nullptr;
    double* brdf_args = nullptr;  // [nargs, nwavel]; Lambertian: albedo[nwavel]
    std::vector<double> default_albedo;
    std::map<std::string, SurfaceMappingImpl> mappings;
};

struct Atmosphere {
    AtmosphereStorage* storage = nullptr;
    Surface* surface = nullptr;
    bool calc_derivs = false;
    bool calc_emission_derivs = false;
    unsigned long long revision = 0;  // Atmosphere::m_revision (cpp/include/sasktran2/atmosphere/atmosphere.h:43, 183-186)
};

struct DerivMem {
    double* ptr;
    int nrad, nstokes, nderiv;
};
struct OutputC {
    std::vector<double> los_optical_depth;  // [nwavel, nlos] column-major as upstream (Output::m_los_optical_depth), filled when the config asks
    double* radiance = nullptr;
    int nrad = 0, nstokes = 1;
    double* flux = nullptr;
    int nflux = 0;
    std::map<std::string, DerivMem> derivs;
    std::map<std::string, DerivMem> surface_derivs;
};

struct Engine {
    Config cfg;
    Geometry1D* geometry = nullptr;
    ViewingGeometry* viewing = nullptr;
    std::unique_ptr<disco::DeviceEngine> dev;
    int ncols() const { return dev ? dev->radiance_columns() : 0; }  // radiance columns: lines of sight
    Atmosphere* atmosphere = nullptr;  // set by calculate_radiance(only_initialize) for block calls
    int staged_start = 0, staged_count = 0;
    std::mutex mtx;
};

