/*
 * sasktran2_b200 — C ABI of the B200-native discrete-ordinates radiance engine.
 *
 * Drop-in boundary: every `sk_*` symbol below has exactly the signature of the reference's C ABI
 * (usask-arg/sasktran2, cpp/include/c_api/*.h; file:line cited per group), so the reference's Rust layer
 * (rust/sasktran2-sys bindgen output -> rust/sasktran2-rs/src/bindings/*.rs) can link against
 * libsasktran2_b200.so instead of libcsasktran2 for the discrete-ordinates path.  Handles are opaque heap
 * objects; every numeric array is owned by the caller and must outlive the handle that maps it
 * (rust/sasktran2-rs/src/bindings/atmosphere_storage.rs:21-33, output.rs:240-281).
 *
 * Error convention (cpp/c_api/engine.cpp:52-82, 468-499): 0 OK; -1 null / uninitialised handle; -2 bad
 * argument; -3 failure inside the solve (message via sk_b200_last_error()).  `*_create` returns NULL on failure.
 *
 * Scope of the CUDA path: num_stokes = 1, multiple_scatter_source = DiscreteOrdinates (0),
 * single_scatter_source = DiscreteOrdinates (2) or None (3), plane-parallel / pseudo-spherical Geometry1D,
 * ground-viewing lines of sight with the observer at or above the top of the atmosphere, Lambertian surface.
 * Anything else is refused at sk_engine_create / sk_engine_calculate_radiance — there is no CPU fallback.
 *
 * `sk_b200_*` symbols are extensions (device-resident staging, timing, device selection).
 */
#ifndef SASKTRAN2_B200_H
#define SASKTRAN2_B200_H

#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct Config Config;
typedef struct Geometry1D Geometry1D;
typedef struct ViewingGeometry ViewingGeometry;
typedef struct AtmosphereStorage AtmosphereStorage;
typedef struct DerivativeMapping DerivativeMapping;
typedef struct SurfaceDerivativeMapping SurfaceDerivativeMapping;
typedef struct Atmosphere Atmosphere;
typedef struct Surface Surface;
typedef struct BRDF BRDF;
typedef struct OutputC OutputC;
typedef struct Engine Engine;

/* ---- Config: cpp/include/c_api/config.h:1-138 (enum values cpp/include/sasktran2/config.h:41-71,
 *      defaults cpp/lib/config/config.cpp:5-33) ---- */
Config* sk_config_create();
void sk_config_destroy(Config* config);
int sk_config_get_num_stokes(Config* config, int* num_stokes);
int sk_config_set_num_stokes(Config* config, int num_stokes);
int sk_config_get_multiple_scatter_source(Config* config, int* multiple_scatter_source);
int sk_config_set_multiple_scatter_source(Config* config, int multiple_scatter_source);
int sk_config_get_single_scatter_source(Config* config, int* single_scatter_source);
int sk_config_set_single_scatter_source(Config* config, int single_scatter_source);
int sk_config_get_num_streams(Config* config, int* num_streams);
int sk_config_set_num_streams(Config* config, int num_streams);
int sk_config_get_num_threads(Config* config, int* num_threads);
int sk_config_set_num_threads(Config* config, int num_threads);
int sk_config_get_threading_model(Config* config, int* threading_model);
int sk_config_set_threading_model(Config* config, int threading_model);
int sk_config_get_wavelength_batch_size(Config* config, int* batch_size);
int sk_config_set_wavelength_batch_size(Config* config, int batch_size);
int sk_config_get_num_singlescatter_moments(Config* config, int* num_moments);
int sk_config_set_num_singlescatter_moments(Config* config, int num_moments);
int sk_config_get_apply_delta_scaling(Config* config, int* apply_delta_scaling);
int sk_config_set_apply_delta_scaling(Config* config, int apply_delta_scaling);
int sk_config_get_num_do_sza(Config* config, int* num_sza);
int sk_config_set_num_do_sza(Config* config, int num_sza);
int sk_config_get_num_do_forced_azimuth(Config* config, int* num_forced_azimuth);
int sk_config_set_num_do_forced_azimuth(Config* config, int num_forced_azimuth);
int sk_config_get_do_backprop(Config* config, int* do_backprop);
int sk_config_set_do_backprop(Config* config, int do_backprop);
int sk_config_get_emission_source(Config* config, int* emission_source);
int sk_config_set_emission_source(Config* config, int emission_source);
int sk_config_get_occultation_source(Config* config, int* occultation_source);
int sk_config_set_occultation_source(Config* config, int occultation_source);
int sk_config_get_solar_refraction(Config* config, int* refraction);
int sk_config_set_solar_refraction(Config* config, int refraction);
int sk_config_get_wf_enabled(Config* config, int* enabled);
int sk_config_set_wf_enabled(Config* config, int enabled);
int sk_config_get_wf_precision(Config* config, int* precision);
int sk_config_set_wf_precision(Config* config, int precision);
int sk_config_get_input_validation_mode(Config* config, int* mode);
int sk_config_set_input_validation_mode(Config* config, int mode);
int sk_config_get_log_level(Config* config, int* log_level);
int sk_config_set_log_level(Config* config, int log_level);

/* ---- Geometry1D: cpp/include/c_api/geometry.h:10-21 ---- */
Geometry1D* sk_geometry1d_create(double cos_sza, double saa, double earth_radius, double* grid_values,
                                 int ngrid_values, int interp_method, int geotype);
void sk_geometry1d_destroy(Geometry1D* geometry);
int sk_geometry1d_get_num_altitudes(const Geometry1D* geometry);
int sk_geometry1d_get_altitudes(const Geometry1D* geometry, double* altitudes);

/* ---- ViewingGeometry: cpp/include/c_api/viewing_geometry.h:9-39 ---- */
ViewingGeometry* sk_viewing_geometry_create();
void sk_viewing_geometry_destroy(ViewingGeometry* geometry);
void sk_viewing_geometry_add_ground_viewing_solar(ViewingGeometry* geometry, double cos_sza,
                                                  double relative_azimuth_angle, double observeraltitude,
                                                  double cos_viewing_zenith);
int sk_viewing_geometry_num_rays(ViewingGeometry* geometry, int* num_rays);
int sk_viewing_geometry_num_flux_observers(ViewingGeometry* geometry, int* num_observers);

/* ---- AtmosphereStorage / Atmosphere / Surface: cpp/include/c_api/atmosphere.h:86-190 ---- */
AtmosphereStorage* sk_atmosphere_storage_create(int nlocation, int nwavel, int nphase_moments, int nstokes,
                                                double* ssa, double* total_extinction, double* emission_source,
                                                double* leg_coeff, double* solar_irradiance);
void sk_atmosphere_storage_destroy(AtmosphereStorage* storage);
int sk_atmosphere_storage_get_derivative_mapping(AtmosphereStorage* storage, const char* name,
                                                 DerivativeMapping** mapping);
int sk_atmosphere_storage_get_derivative_mapping_by_index(AtmosphereStorage* storage, int index,
                                                          DerivativeMapping** mapping);
int sk_atmosphere_storage_get_num_derivative_mappings(AtmosphereStorage* storage, int* num_mappings);
int sk_atmosphere_storage_get_derivative_mapping_name(AtmosphereStorage* storage, int index, const char** name);
int sk_atmosphere_storage_finalize_scattering_derivatives(AtmosphereStorage* storage);
int sk_atmosphere_storage_set_zero(AtmosphereStorage* storage);
Atmosphere* sk_atmosphere_create(AtmosphereStorage* storage, Surface* surface, int calculate_derivatives,
                                 int calculate_emission_derivatives);
void sk_atmosphere_destroy(Atmosphere* atmosphere);
int sk_atmosphere_apply_delta_m_scaling(Atmosphere* atmosphere, int order);
Surface* sk_surface_create(int nwavel, int nstokes, double* emission);
void sk_surface_destroy(Surface* surface);
int sk_surface_set_brdf(Surface* surface, BRDF* brdf, double* brdf_args);
int sk_surface_get_derivative_mapping(Surface* storage, const char* name, SurfaceDerivativeMapping** mapping);
int sk_surface_get_num_derivative_mappings(Surface* storage, int* num_mappings);
int sk_surface_get_derivative_mapping_name(Surface* storage, int index, const char** name);
int sk_surface_set_zero(Surface* storage);

/* ---- BRDF: cpp/include/c_api/brdf.h:9-16 (Lambertian only on the CUDA path) ---- */
BRDF* sk_brdf_create_lambertian(int nstokes);
int sk_brdf_get_num_deriv(BRDF* config, int* num_deriv);
int sk_brdf_get_num_args(BRDF* config, int* num_args);
void sk_brdf_destroy(BRDF* config);

/* ---- DerivativeMapping: cpp/include/c_api/deriv_mapping.h:10-88 ---- */
int sk_deriv_mapping_destroy(DerivativeMapping* mapping);
int sk_deriv_mapping_set_zero(DerivativeMapping* mapping);
int sk_deriv_mapping_get_d_ssa(DerivativeMapping* mapping, double** ssa);
int sk_deriv_mapping_get_d_extinction(DerivativeMapping* mapping, double** extinction);
int sk_deriv_mapping_get_scat_factor(DerivativeMapping* mapping, double** scat_factor);
int sk_deriv_mapping_get_d_legendre(DerivativeMapping* mapping, double** d_legendre);
int sk_deriv_mapping_get_scat_deriv_index(DerivativeMapping* mapping, int* scat_deriv_index);
int sk_deriv_mapping_set_scat_deriv_index(DerivativeMapping* mapping, int scat_deriv_index);
int sk_deriv_mapping_get_num_location(DerivativeMapping* mapping, int* num_location);
int sk_deriv_mapping_get_num_wavel(DerivativeMapping* mapping, int* num_wavel);
int sk_deriv_mapping_get_num_legendre(DerivativeMapping* mapping, int* num_legendre);
int sk_deriv_mapping_set_interp_dim(DerivativeMapping* mapping, const char* name);
int sk_deriv_mapping_set_assign_name(DerivativeMapping* mapping, const char* name);
int sk_deriv_mapping_set_log_radiance_space(DerivativeMapping* mapping, int log_radiance_space);
int sk_deriv_mapping_get_log_radiance_space(DerivativeMapping* mapping, int* log_radiance_space);
int sk_deriv_mapping_is_scattering_derivative(DerivativeMapping* mapping, int* is_scattering_derivative);
int sk_deriv_mapping_get_num_output(DerivativeMapping* mapping, int* num_output);
int sk_deriv_mapping_get_assign_name(DerivativeMapping* mapping, const char** name);
int sk_deriv_mapping_get_interp_dim(DerivativeMapping* mapping, const char** name);
int sk_deriv_mapping_set_interpolator(DerivativeMapping* mapping, double* interpolator, int dim1, int dim2);
int sk_deriv_mapping_clear_interpolator(DerivativeMapping* mapping);
int sk_deriv_mapping_get_interpolator(DerivativeMapping* mapping, double** interpolator, int* dim1, int* dim2);
int sk_surface_deriv_mapping_get_num_wavel(SurfaceDerivativeMapping* mapping, int* num_wavel);
int sk_surface_deriv_mapping_get_num_brdf_args(SurfaceDerivativeMapping* mapping, int* num_brdf_args);
int sk_surface_deriv_mapping_get_d_brdf(SurfaceDerivativeMapping* mapping, double** brdf);
int sk_surface_deriv_mapping_set_zero(SurfaceDerivativeMapping* mapping);
int sk_surface_deriv_mapping_destroy(SurfaceDerivativeMapping* mapping);

/* ---- Output: cpp/include/c_api/output.h:11-33 ---- */
OutputC* sk_output_create(double* radiance, int nrad, int nstokes, double* flux, int nflux);
void sk_output_destroy(OutputC* config);
int sk_output_assign_derivative_memory(OutputC* output, const char* name, double* derivative_mapping, int nrad,
                                       int nstokes, int nderiv);
int sk_output_assign_surface_derivative_memory(OutputC* output, const char* name, double* derivative_mapping,
                                               int nrad, int nstokes);

/* ---- Engine: cpp/include/c_api/engine.h:17-50 ---- */
Engine* sk_engine_create(Config* engine, Geometry1D* geometry, ViewingGeometry* viewing_geometry);
int sk_engine_calculate_radiance(Engine* engine, Atmosphere* atmosphere, OutputC* output, int only_initialize);
int sk_engine_effective_wavelength_batch_size(Engine* engine, int num_wavelengths);
int sk_engine_supports_linearization(Engine* engine, int mode, int* supported);
int sk_engine_linearization_backend(Engine* engine, int mode, int* backend);
int sk_engine_calculate_radiance_block_thread(Engine* engine, OutputC* output, int wavelength_start,
                                              int wavelength_count, int thread_idx);
void sk_engine_destroy(Engine* engine);
int sk_openmp_support_enabled();

/* ---- extensions ---- */
const char* sk_b200_last_error();
int sk_b200_device_count();
int sk_b200_set_device(int device);
/* Copy wavelengths [start, start+count) of the atmosphere (and, when `output` has derivative memory assigned,
 * the derivative mappings) to the device and keep them resident.  `output` may be NULL (radiances only). */
int sk_b200_engine_stage_atmosphere(Engine* engine, Atmosphere* atmosphere, OutputC* output, int wavelength_start,
                                    int wavelength_count);
/* Run the kernels on the staged wavelengths; nothing crosses PCIe. */
int sk_b200_engine_solve_staged(Engine* engine);
/* Copy the staged range's results into the output buffers (same offsets as the full-spectrum call). */
int sk_b200_engine_fetch_output(Engine* engine, OutputC* output);
/* ms of the last staged solve / full call: [h2d, optics, layer, bvp, radiance, d2h, kernels_total, wf,
   wf_adjoint, wf_layer, wf_chain, wf_map] (wf = sum of the last four) */
int sk_b200_engine_get_timings(Engine* engine, double* out_ms, int n);
long long sk_b200_engine_kernel_launches(Engine* engine);
/* number of azimuth orders solved and wavelengths per workspace chunk (diagnostics) */
int sk_b200_engine_info(Engine* engine, int* num_azimuth, int* chunk_wavelengths, double* workspace_mb_per_wavelength);
int sk_b200_engine_set_workspace_gb(Engine* engine, double gb);
/* DFMA micro-benchmark on the current device: the FP64 roofline denominator (TFLOP/s) */
double sk_b200_measure_fp64_tflops();
/* 1 when the adjoint boundary-value solve of an (N = num_streams / 2, nlos) problem reuses the forward LU factors
 * (transposed solves, the reference's dgbtrs('T')), 0 when it factorises A^T a second time. */
int sk_b200_adjoint_reuses_factors(int n_half_streams, int nlos);
/* page-locked host memory for caller-side buffers (falls back to malloc without a CUDA device) */
void* sk_b200_host_alloc(size_t nbytes);
void sk_b200_host_free(void* p);
/* test/debug: copy a named workspace array of the last solved chunk (names: see disco_engine.cu) */
long long sk_b200_engine_debug_copy(Engine* engine, const char* name, double* host, long long max_n);

#ifdef __cplusplus
}
#endif
#endif
