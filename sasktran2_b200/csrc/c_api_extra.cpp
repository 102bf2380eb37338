// The remainder of the reference's C ABI (cpp/include/c_api/*.h): every `sk_*` symbol the Rust layer
// (rust/sasktran2-rs/src/bindings/*.rs) references, so that libsasktran2_b200.so links in place of libcsasktran2.
// Entry points on or next to the discrete-ordinates path are real (config accessors, atmosphere revision counter,
// refractive index, viewing-geometry constructors, surface mapping interpolators, sk_lapack_dgesv).  Entry points of
// subsystems this library does not contain (geodetic, 2-D geometry, JVP / VJP drivers, flux outputs) return the
// reference's failure codes (-3 / NULL) and leave a message in sk_b200_last_error(): there is no CPU fallback.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>

#include "c_api_types.h"

using skapi::fail;

typedef struct Geometry2D Geometry2D;
typedef struct Geodetic Geodetic;
typedef struct OutputJVP OutputJVP;
typedef struct OutputVJP OutputVJP;

namespace {
int unsupported(const char* what) {
    return fail(-3, std::string("sasktran2_b200: ") + what + " is outside the B200 discrete-ordinates path (no CPU fallback)");
}
}  // namespace

extern "C" {

// ---- Config: cpp/include/c_api/config.h:31-137 -----------------------------------------------------------------
#define CFG_INT(name, field)                       \
    int sk_config_get_##name(Config* c, int* v) {  \
        if (!c || !v) return -1;                   \
        *v = c->field;                             \
        return 0;                                  \
    }                                              \
    int sk_config_set_##name(Config* c, int v) {   \
        if (!c) return -1;                         \
        c->field = v;                              \
        return 0;                                  \
    }
#define CFG_DBL(name, field)                         \
    int sk_config_get_##name(Config* c, double* v) { \
        if (!c || !v) return -1;                     \
        *v = c->field;                               \
        return 0;                                    \
    }                                                \
    int sk_config_set_##name(Config* c, double v) {  \
        if (!c) return -1;                           \
        c->field = v;                                \
        return 0;                                    \
    }
CFG_INT(singlescatter_phasemode, singlescatter_phasemode)
CFG_INT(num_do_spherical_iterations, num_do_spherical_iterations)
CFG_INT(num_hr_spherical_iterations, num_hr_spherical_iterations)
CFG_INT(num_hr_incoming, num_hr_incoming)
CFG_INT(num_hr_outgoing, num_hr_outgoing)
CFG_INT(num_hr_full_incoming_points, num_hr_full_incoming_points)
CFG_INT(initialize_hr_with_do, initialize_hr_with_do)
CFG_DBL(successive_orders_relative_tolerance, successive_orders_relative_tolerance)
CFG_DBL(successive_orders_absolute_tolerance, successive_orders_absolute_tolerance)
CFG_INT(successive_orders_anderson_depth, successive_orders_anderson_depth)
CFG_DBL(successive_orders_damping, successive_orders_damping)
CFG_INT(los_refraction, los_refraction)
CFG_INT(multiple_scatter_refraction, multiple_scatter_refraction)
CFG_INT(stokes_basis, stokes_basis)
CFG_INT(output_los_optical_depth, output_los_optical_depth)

int sk_config_get_num_successive_orders_altitudes(Config* c, int* n) {
    if (!c || !n) return -1;
    *n = (int)c->successive_orders_altitude_grid_m.size();
    return 0;
}
int sk_config_get_successive_orders_altitude_grid_m(Config* c, double* grid) {
    if (!c || !grid) return -1;
    std::copy(c->successive_orders_altitude_grid_m.begin(), c->successive_orders_altitude_grid_m.end(), grid);
    return 0;
}
int sk_config_set_successive_orders_altitude_grid_m(Config* c, const double* grid, int n) {
    if (!c || (n > 0 && !grid)) return -1;
    if (n < 0) return -2;
    c->successive_orders_altitude_grid_m.assign(grid, grid + n);
    return 0;
}
int sk_config_get_num_flux_types(Config* c, int* n) {
    if (!c || !n) return -1;
    *n = (int)c->flux_types.size();
    return 0;
}
int sk_config_get_flux_types(Config* c, int* types) {
    if (!c || !types) return -1;
    std::copy(c->flux_types.begin(), c->flux_types.end(), types);
    return 0;
}
int sk_config_set_flux_types(Config* c, const int* types, int n) {
    if (!c || (n > 0 && !types)) return -1;
    if (n < 0) return -2;
    c->flux_types.assign(types, types + n);
    return 0;
}

// ---- Atmosphere revision counter: cpp/include/c_api/atmosphere.h:152-155 (cpp/c_api/atmosphere.cpp:473-507) ----
int sk_atmosphere_mark_changed(Atmosphere* a) {
    if (!a) return -1;
    ++a->revision;
    return 0;
}
int sk_atmosphere_get_revision(Atmosphere* a, unsigned long long* revision) {
    if (!a || !revision) return -1;
    *revision = a->revision;
    return 0;
}

// ---- Geometry: cpp/include/c_api/geometry.h:22-61 ----------------------------------------------------------------
int sk_geometry1d_get_refractive_index_ptr(const Geometry1D* g, double** refractive_index) {
    if (!g || !refractive_index) return -1;
    auto* gm = const_cast<Geometry1D*>(g);
    if (gm->refractive_index.size() != gm->spec.altitudes.size()) gm->refractive_index.assign(gm->spec.altitudes.size(), 1.0);
    *refractive_index = gm->refractive_index.data();
    return 0;
}
Geometry2D* sk_geometry2d_create(double, double, double, const double*, int, const double*, int, int) {
    unsupported("Geometry2D");
    return nullptr;
}
void sk_geometry2d_destroy(Geometry2D*) {}
int sk_geometry2d_get_location_shape(const Geometry2D*, int*, int*) { return unsupported("Geometry2D"); }
int sk_geometry2d_get_altitudes(const Geometry2D*, double*) { return unsupported("Geometry2D"); }
int sk_geometry2d_get_horizontal_angles(const Geometry2D*, double*) { return unsupported("Geometry2D"); }
int sk_geometry2d_get_refractive_index_ptr(const Geometry2D*, const double**) { return unsupported("Geometry2D"); }
int sk_geometry2d_get_refractive_index_mut_ptr(Geometry2D*, double**) { return unsupported("Geometry2D"); }
int sk_geometry2d_get_location_index(const Geometry2D*, int, int, int*) { return unsupported("Geometry2D"); }

// ---- Viewing geometry: cpp/include/c_api/viewing_geometry.h:18-37 ------------------------------------------------
int sk_viewing_geometry_add_tangent_altitude_solar(ViewingGeometry* v, double tangent_altitude_m, double relative_azimuth_angle,
                                                   double observeraltitude, double cos_sza) {
    if (!v) return -1;
    // spherical path (row a14): TangentAltitudeSolar(tangent_altitude, rel_az, observer_altitude, cos_sza)
    v->ordered.push_back({1, {tangent_altitude_m, relative_azimuth_angle, observeraltitude, cos_sza}});
    v->num_tangent_rays += 1;
    return 0;
}
int sk_viewing_geometry_add_tangent_altitude(ViewingGeometry* v, double tangent_altitude_m, double observer_altitude_m,
                                             double horizontal_angle_radians, double viewing_azimuth_radians) {
    if (!v) return -1;
    v->other_rays.push_back({2, tangent_altitude_m, observer_altitude_m, horizontal_angle_radians, viewing_azimuth_radians});
    return 0;
}
int sk_viewing_geometry_add_solar_angles_observer_location(ViewingGeometry* v, double cos_sza, double relative_azimuth_angle,
                                                           double cos_viewing_zenith, double observeraltitude) {
    if (!v) return -1;
    v->other_rays.push_back({3, cos_sza, relative_azimuth_angle, cos_viewing_zenith, observeraltitude});
    return 0;
}
int sk_viewing_geometry_add_flux_observer_solar(ViewingGeometry* v, double, double) {
    if (!v) return -1;
    v->num_flux_observers += 1;  // refused by sk_engine_create: flux outputs are not computed on the CUDA path
    return 0;
}

// ---- Derivative mappings: cpp/include/c_api/deriv_mapping.h:19-20, 58-83 -----------------------------------------
int sk_deriv_mapping_get_d_emission(DerivativeMapping* m, double** d_emission) {
    if (!m || !d_emission) return -1;
    if (m->impl->d_emission.empty()) m->impl->d_emission.assign((size_t)m->impl->nloc * m->impl->nwavel, 0.0);
    *d_emission = m->impl->d_emission.data();
    return 0;
}
int sk_surface_deriv_mapping_get_d_emission(SurfaceDerivativeMapping* m, double** emission) {
    if (!m || !emission) return -1;
    if (m->impl->d_emission.empty()) m->impl->d_emission.assign((size_t)std::max(m->impl->nwavel, 1), 0.0);
    *emission = m->impl->d_emission.data();
    return 0;
}
int sk_surface_deriv_mapping_get_interpolator(SurfaceDerivativeMapping* m, double** interpolator, int* dim1, int* dim2) {
    if (!m || !interpolator || !dim1 || !dim2) return -1;
    *interpolator = m->impl->interpolator.empty() ? nullptr : m->impl->interpolator.data();
    *dim1 = m->impl->interp_d1;
    *dim2 = m->impl->interp_d2;
    return 0;
}
int sk_surface_deriv_mapping_set_interpolator(SurfaceDerivativeMapping* m, double* interpolator, int dim1, int dim2) {
    if (!m || !interpolator) return -1;
    if (dim1 < 0 || dim2 < 0) return -2;
    m->impl->interpolator.assign(interpolator, interpolator + (size_t)dim1 * dim2);
    m->impl->interp_d1 = dim1;
    m->impl->interp_d2 = dim2;
    return 0;
}
int sk_surface_deriv_mapping_get_interp_dim(SurfaceDerivativeMapping* m, const char** name) {
    if (!m || !name) return -1;
    *name = m->impl->interp_dim.c_str();
    return 0;
}
int sk_surface_deriv_mapping_set_interp_dim(SurfaceDerivativeMapping* m, const char* name) {
    if (!m || !name) return -1;
    m->impl->interp_dim = name;
    return 0;
}

// ---- BRDF: cpp/include/c_api/brdf.h:11-12 (num_args / num_deriv: cpp/include/sasktran2/atmosphere/surface.h:232-234,
//      352-354).  The handles exist so that callers can build their surfaces; sk_engine_calculate_radiance refuses a
//      non-Lambertian BRDF (row a10: only the closed-form Lambertian expansion is on the CUDA path). ----
BRDF* sk_brdf_create_kokhanovsky(int nstokes) {
    auto* b = new BRDF();
    b->kind = 1;
    b->nstokes = nstokes;
    return b;
}
BRDF* sk_brdf_create_modis(int nstokes) {
    auto* b = new BRDF();
    b->kind = 2;
    b->nstokes = nstokes;
    return b;
}

// ---- Output: cpp/include/c_api/output.h:22-53 ---------------------------------------------------------------------
int sk_output_assign_flux_derivative_memory(OutputC* o, const char*, double*, int, int) {
    if (!o) return -1;
    return unsupported("flux output");
}
int sk_output_assign_surface_flux_derivative_memory(OutputC* o, const char*, double*, int) {
    if (!o) return -1;
    return unsupported("flux output");
}
int sk_output_get_los_optical_depth(OutputC* o, double** od) {
    if (!o || !od) return -1;
    // cpp/c_api/output.cpp:311-326: pointer to the [nwavel, nlos] matrix the engine filled (config.output_los_optical_depth)
    if (o->los_optical_depth.empty()) {
        *od = nullptr;
        return skapi::fail(-2, "line-of-sight optical depths were not computed (spherical path with config.output_los_optical_depth only)");
    }
    *od = o->los_optical_depth.data();
    return 0;
}
OutputJVP* sk_output_jvp_create(double*, double*, int, int) {
    unsupported("the JVP driver (sk_engine_linearization_backend reports Jacobian-only)");
    return nullptr;
}
void sk_output_jvp_destroy(OutputJVP*) {}
int sk_output_jvp_assign_derivative_tangent(OutputJVP*, const char*, const double*, int) { return unsupported("the JVP driver"); }
int sk_output_jvp_assign_surface_tangent(OutputJVP*, const char*, const double*, int) { return unsupported("the JVP driver"); }
OutputVJP* sk_output_vjp_create(double*, const double*, int, int) {
    unsupported("the VJP driver (sk_engine_linearization_backend reports Jacobian-only)");
    return nullptr;
}
void sk_output_vjp_destroy(OutputVJP*) {}
int sk_output_vjp_assign_derivative_gradient(OutputVJP*, const char*, double*, int) { return unsupported("the VJP driver"); }
int sk_output_vjp_assign_surface_gradient(OutputVJP*, const char*, double*, int) { return unsupported("the VJP driver"); }
int sk_output_vjp_finalize(OutputVJP*) { return unsupported("the VJP driver"); }

// ---- Engine: cpp/include/c_api/engine.h:19-41 ---------------------------------------------------------------------
// The Rust layer asks sk_engine_linearization_backend first; this engine answers "Jacobian" for both modes, so the
// JVP / VJP products are formed upstream from the streamed weighting functions and these are never reached.
Engine* sk_engine_create_2d(Config*, Geometry2D*, ViewingGeometry*) {
    unsupported("Geometry2D");
    return nullptr;
}
int sk_engine_calculate_jvp(Engine*, Atmosphere*, OutputJVP*) { return unsupported("the JVP driver"); }
int sk_engine_initialize_jvp(Engine*, Atmosphere*, OutputJVP*) { return unsupported("the JVP driver"); }
int sk_engine_calculate_jvp_wavelength_thread(Engine*, OutputJVP*, int, int) { return unsupported("the JVP driver"); }
int sk_engine_calculate_vjp(Engine*, Atmosphere*, OutputVJP*) { return unsupported("the VJP driver"); }
int sk_engine_initialize_vjp(Engine*, Atmosphere*, OutputVJP*) { return unsupported("the VJP driver"); }
int sk_engine_calculate_vjp_block_thread(Engine*, OutputVJP*, int, int, int) { return unsupported("the VJP driver"); }

// ---- Geodetic: cpp/include/c_api/geodetic.h (coordinate helper of the Python layer, not on the path) --------------
Geodetic* sk_geodetic_create(double, double) {
    unsupported("Geodetic");
    return nullptr;
}
void sk_geodetic_destroy(Geodetic*) {}
int sk_geodetic_get_altitude(const Geodetic*, double*) { return unsupported("Geodetic"); }
int sk_geodetic_get_latitude(const Geodetic*, double*) { return unsupported("Geodetic"); }
int sk_geodetic_get_longitude(const Geodetic*, double*) { return unsupported("Geodetic"); }
int sk_geodetic_get_location(const Geodetic*, double*, double*, double*) { return unsupported("Geodetic"); }
int sk_geodetic_get_local_south(const Geodetic*, double*, double*, double*) { return unsupported("Geodetic"); }
int sk_geodetic_get_local_up(const Geodetic*, double*, double*, double*) { return unsupported("Geodetic"); }
int sk_geodetic_get_local_west(const Geodetic*, double*, double*, double*) { return unsupported("Geodetic"); }
int sk_geodetic_get_altitude_intercepts(const Geodetic*, double, double, double, double, double, double, double, double*,
                                        double*, double*, double*, double*, double*) {
    return unsupported("Geodetic");
}
int sk_geodetic_from_lat_lon_altitude(const Geodetic*, double, double, double) { return unsupported("Geodetic"); }
int sk_geodetic_from_tangent_altitude(const Geodetic*, double, double, double, double, double, double, double, double*,
                                      double*, double*) {
    return unsupported("Geodetic");
}
int sk_geodetic_from_tangent_point(const Geodetic*, double, double, double, double, double, double) {
    return unsupported("Geodetic");
}
int sk_geodetic_from_xyz(const Geodetic*, double, double, double) { return unsupported("Geodetic"); }
int sk_geodetic_is_valid(const Geodetic*, int*) { return unsupported("Geodetic"); }
int sk_geodetic_get_osculating_spheroid(const Geodetic*, double*, double*, double*, double*) { return unsupported("Geodetic"); }

// ---- sk_lapack_dgesv: cpp/include/c_api/sk_lapack.h (LAPACK dgesv semantics: column-major A (lda x n) overwritten by
//      its LU factors with partial pivoting, 1-based ipiv, B (ldb x nrhs) overwritten by the solution; returns info) ----
long long sk_lapack_dgesv(long long n_, long long nrhs_, double* a, long long lda_, long long* ipiv_, double* b, long long ldb_) {
    const int64_t n = n_, nrhs = nrhs_, lda = lda_, ldb = ldb_;
    long long* ipiv = ipiv_;
    if (n < 0) return -1;
    if (nrhs < 0) return -2;
    if (!a || lda < std::max<int64_t>(1, n)) return -4;
    if (!ipiv) return -5;
    if (!b || ldb < std::max<int64_t>(1, n)) return -7;
    int64_t info = 0;
    for (int64_t k = 0; k < n; ++k) {
        int64_t p = k;
        double amax = std::abs(a[k + k * lda]);
        for (int64_t i = k + 1; i < n; ++i)
            if (std::abs(a[i + k * lda]) > amax) {
                amax = std::abs(a[i + k * lda]);
                p = i;
            }
        ipiv[k] = p + 1;
        if (a[p + k * lda] == 0.0) {
            if (info == 0) info = k + 1;
            continue;
        }
        if (p != k)
            for (int64_t c = 0; c < n; ++c) std::swap(a[p + c * lda], a[k + c * lda]);
        const double r = 1.0 / a[k + k * lda];
        for (int64_t i = k + 1; i < n; ++i) a[i + k * lda] *= r;
        for (int64_t c = k + 1; c < n; ++c) {
            const double t = a[k + c * lda];
            if (t != 0.0)
                for (int64_t i = k + 1; i < n; ++i) a[i + c * lda] -= a[i + k * lda] * t;
        }
    }
    if (info != 0) return info;
    for (int64_t r = 0; r < nrhs; ++r) {
        double* x = b + r * ldb;
        for (int64_t k = 0; k < n; ++k)
            if (ipiv[k] - 1 != k) std::swap(x[k], x[ipiv[k] - 1]);
        for (int64_t k = 0; k < n; ++k)
            for (int64_t i = k + 1; i < n; ++i) x[i] -= a[i + k * lda] * x[k];
        for (int64_t k = n - 1; k >= 0; --k) {
            x[k] /= a[k + k * lda];
            for (int64_t i = 0; i < k; ++i) x[i] -= a[i + k * lda] * x[k];
        }
    }
    return 0;
}

}  // extern "C"
