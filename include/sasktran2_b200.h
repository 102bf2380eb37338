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
typedef struct Geometry2D Geometry2D;
typedef struct Geodetic Geodetic;
typedef struct OutputJVP OutputJVP;
typedef struct OutputVJP OutputVJP;

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
/* emission source (cpp/include/sasktran2/config.h:122-128): 1 none, 2 discrete_ordinates - the storage's emission_source
   array goes through the DO solve (plane-parallel / pseudo-spherical, radiances only; needs both scattering sources on the
   DO path as upstream's Config::validate_config); 0 standard, 3 volume_emission_rate, 4 twostream: engine creation fails */
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
/* stored and returned; options of sources outside the CUDA path (HR, successive orders, refraction, flux) */
int sk_config_get_singlescatter_phasemode(Config* config, int* phasemode);
int sk_config_set_singlescatter_phasemode(Config* config, int phasemode);
int sk_config_get_num_do_spherical_iterations(Config* config, int* num_iterations);
int sk_config_set_num_do_spherical_iterations(Config* config, int num_iterations);
int sk_config_get_num_hr_spherical_iterations(Config* config, int* num_iterations);
int sk_config_set_num_hr_spherical_iterations(Config* config, int num_iterations);
int sk_config_get_num_hr_incoming(Config* config, int* num_incoming);
int sk_config_set_num_hr_incoming(Config* config, int num_incoming);
int sk_config_get_num_hr_outgoing(Config* config, int* num_outgoing);
int sk_config_set_num_hr_outgoing(Config* config, int num_outgoing);
int sk_config_get_num_hr_full_incoming_points(Config* config, int* num_points);
int sk_config_set_num_hr_full_incoming_points(Config* config, int num_points);
int sk_config_get_initialize_hr_with_do(Config* config, int* initialize);
int sk_config_set_initialize_hr_with_do(Config* config, int initialize);
int sk_config_get_successive_orders_relative_tolerance(Config* config, double* tolerance);
int sk_config_set_successive_orders_relative_tolerance(Config* config, double tolerance);
int sk_config_get_successive_orders_absolute_tolerance(Config* config, double* tolerance);
int sk_config_set_successive_orders_absolute_tolerance(Config* config, double tolerance);
int sk_config_get_successive_orders_anderson_depth(Config* config, int* depth);
int sk_config_set_successive_orders_anderson_depth(Config* config, int depth);
int sk_config_get_successive_orders_damping(Config* config, double* damping);
int sk_config_set_successive_orders_damping(Config* config, double damping);
int sk_config_get_num_successive_orders_altitudes(Config* config, int* num_altitudes);
int sk_config_get_successive_orders_altitude_grid_m(Config* config, double* altitude_grid_m);
int sk_config_set_successive_orders_altitude_grid_m(Config* config, const double* altitude_grid_m, int num_altitudes);
int sk_config_get_los_refraction(Config* config, int* refraction);
int sk_config_set_los_refraction(Config* config, int refraction);
int sk_config_get_multiple_scatter_refraction(Config* config, int* refraction);
int sk_config_set_multiple_scatter_refraction(Config* config, int refraction);
int sk_config_get_stokes_basis(Config* config, int* basis);
int sk_config_set_stokes_basis(Config* config, int basis);
int sk_config_get_output_los_optical_depth(Config* config, int* output);
int sk_config_set_output_los_optical_depth(Config* config, int output);
int sk_config_get_num_flux_types(Config* config, int* num_flux_types);
int sk_config_get_flux_types(Config* config, int* flux_types);
int sk_config_set_flux_types(Config* config, const int* flux_types, int num_flux_types);

/* ---- Geometry1D: cpp/include/c_api/geometry.h:10-21 ---- */
Geometry1D* sk_geometry1d_create(double cos_sza, double saa, double earth_radius, double* grid_values,
                                 int ngrid_values, int interp_method, int geotype);
void sk_geometry1d_destroy(Geometry1D* geometry);
int sk_geometry1d_get_num_altitudes(const Geometry1D* geometry);
int sk_geometry1d_get_altitudes(const Geometry1D* geometry, double* altitudes);
int sk_geometry1d_get_refractive_index_ptr(const Geometry1D* geometry, double** refractive_index);
/* Geometry2D (cpp/include/c_api/geometry.h:24-61): not on the path; create returns NULL, accessors -3 */
Geometry2D* sk_geometry2d_create(double cos_sza, double saa, double earth_radius, const double* altitude_grid_values,
                                 int num_altitudes, const double* horizontal_angle_grid_values,
                                 int num_horizontal_locations, int altitude_interp_method);
void sk_geometry2d_destroy(Geometry2D* geometry);
int sk_geometry2d_get_location_shape(const Geometry2D* geometry, int* num_horizontal_locations, int* num_altitudes);
int sk_geometry2d_get_altitudes(const Geometry2D* geometry, double* altitudes);
int sk_geometry2d_get_horizontal_angles(const Geometry2D* geometry, double* horizontal_angles);
int sk_geometry2d_get_refractive_index_ptr(const Geometry2D* geometry, const double** refractive_index);
int sk_geometry2d_get_refractive_index_mut_ptr(Geometry2D* geometry, double** refractive_index);
int sk_geometry2d_get_location_index(const Geometry2D* geometry, int altitude_index, int horizontal_index,
                                     int* location_index);

/* ---- ViewingGeometry: cpp/include/c_api/viewing_geometry.h:9-39 ---- */
ViewingGeometry* sk_viewing_geometry_create();
void sk_viewing_geometry_destroy(ViewingGeometry* geometry);
void sk_viewing_geometry_add_ground_viewing_solar(ViewingGeometry* geometry, double cos_sza,
                                                  double relative_azimuth_angle, double observeraltitude,
                                                  double cos_viewing_zenith);
int sk_viewing_geometry_add_tangent_altitude_solar(ViewingGeometry* geometry, double tangent_altitude_m,
                                                   double relative_azimuth_angle, double observeraltitude,
                                                   double cos_sza);
int sk_viewing_geometry_add_tangent_altitude(ViewingGeometry* geometry, double tangent_altitude_m,
                                             double observer_altitude_m, double horizontal_angle_radians,
                                             double viewing_azimuth_radians);
int sk_viewing_geometry_add_solar_angles_observer_location(ViewingGeometry* geometry, double cos_sza,
                                                           double relative_azimuth_angle, double cos_viewing_zenith,
                                                           double observeraltitude);
int sk_viewing_geometry_add_flux_observer_solar(ViewingGeometry* geometry, double cos_sza, double observeraltitude);
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
int sk_atmosphere_mark_changed(Atmosphere* atmosphere);
int sk_atmosphere_get_revision(Atmosphere* atmosphere, unsigned long long* revision);
Surface* sk_surface_create(int nwavel, int nstokes, double* emission);
void sk_surface_destroy(Surface* surface);
int sk_surface_set_brdf(Surface* surface, BRDF* brdf, double* brdf_args);
int sk_surface_get_derivative_mapping(Surface* storage, const char* name, SurfaceDerivativeMapping** mapping);
int sk_surface_get_num_derivative_mappings(Surface* storage, int* num_mappings);
int sk_surface_get_derivative_mapping_name(Surface* storage, int index, const char** name);
int sk_surface_set_zero(Surface* storage);

/* ---- BRDF: cpp/include/c_api/brdf.h:9-16 (Lambertian only on the CUDA path) ---- */
BRDF* sk_brdf_create_lambertian(int nstokes);
BRDF* sk_brdf_create_kokhanovsky(int nstokes);
BRDF* sk_brdf_create_modis(int nstokes);
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
int sk_deriv_mapping_get_d_emission(DerivativeMapping* mapping, double** d_emission);
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
/* d_brdf is [num_wavel, num_brdf_args] column-major (the surface's BRDF at the time the mapping is first requested):
   Lambertian 1 column (albedo), MODIS 3 (kernel weights); a mapping on the snow model's argument is refused at solve time */
int sk_surface_deriv_mapping_get_d_brdf(SurfaceDerivativeMapping* mapping, double** brdf);
int sk_surface_deriv_mapping_get_d_emission(SurfaceDerivativeMapping* mapping, double** emission);
int sk_surface_deriv_mapping_get_interpolator(SurfaceDerivativeMapping* mapping, double** interpolator, int* dim1,
                                              int* dim2);
int sk_surface_deriv_mapping_set_interpolator(SurfaceDerivativeMapping* mapping, double* interpolator, int dim1,
                                              int dim2);
int sk_surface_deriv_mapping_get_interp_dim(SurfaceDerivativeMapping* mapping, const char** name);
int sk_surface_deriv_mapping_set_interp_dim(SurfaceDerivativeMapping* mapping, const char* name);
int sk_surface_deriv_mapping_set_zero(SurfaceDerivativeMapping* mapping);
int sk_surface_deriv_mapping_destroy(SurfaceDerivativeMapping* mapping);

/* ---- Output: cpp/include/c_api/output.h:11-33 ---- */
OutputC* sk_output_create(double* radiance, int nrad, int nstokes, double* flux, int nflux);
void sk_output_destroy(OutputC* config);
int sk_output_assign_derivative_memory(OutputC* output, const char* name, double* derivative_mapping, int nrad,
                                       int nstokes, int nderiv);
int sk_output_assign_surface_derivative_memory(OutputC* output, const char* name, double* derivative_mapping,
                                               int nrad, int nstokes);

/* flux outputs, LOS optical depth, JVP / VJP containers (cpp/include/c_api/output.h:22-53): outside the path, -3 / NULL */
int sk_output_assign_flux_derivative_memory(OutputC* output, const char* name, double* derivative_mapping, int nrad,
                                            int nderiv);
int sk_output_assign_surface_flux_derivative_memory(OutputC* output, const char* name, double* derivative_mapping,
                                                    int nrad);
int sk_output_get_los_optical_depth(OutputC* output, double** od);
OutputJVP* sk_output_jvp_create(double* radiance, double* jvp, int nrad, int nstokes);
void sk_output_jvp_destroy(OutputJVP* output);
int sk_output_jvp_assign_derivative_tangent(OutputJVP* output, const char* name, const double* tangent, int nparam);
int sk_output_jvp_assign_surface_tangent(OutputJVP* output, const char* name, const double* tangent, int nparam);
OutputVJP* sk_output_vjp_create(double* radiance, const double* cotangent, int nrad, int nstokes);
void sk_output_vjp_destroy(OutputVJP* output);
int sk_output_vjp_assign_derivative_gradient(OutputVJP* output, const char* name, double* gradient, int nparam);
int sk_output_vjp_assign_surface_gradient(OutputVJP* output, const char* name, double* gradient, int nparam);
int sk_output_vjp_finalize(OutputVJP* output);

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
/* 2-D engine and the JVP / VJP drivers (engine.h:19-41): sk_engine_linearization_backend reports Jacobian-only, so the
 * Rust layer forms the products from the streamed weighting functions and never calls these; they return -3 / NULL */
Engine* sk_engine_create_2d(Config* engine, Geometry2D* geometry, ViewingGeometry* viewing_geometry);
int sk_engine_calculate_jvp(Engine* engine, Atmosphere* atmosphere, OutputJVP* output);
int sk_engine_initialize_jvp(Engine* engine, Atmosphere* atmosphere, OutputJVP* output);
int sk_engine_calculate_jvp_wavelength_thread(Engine* engine, OutputJVP* output, int wavelength, int thread_idx);
int sk_engine_calculate_vjp(Engine* engine, Atmosphere* atmosphere, OutputVJP* output);
int sk_engine_initialize_vjp(Engine* engine, Atmosphere* atmosphere, OutputVJP* output);
int sk_engine_calculate_vjp_block_thread(Engine* engine, OutputVJP* output, int wavelength_start, int wavelength_count,
                                         int thread_idx);

/* ---- Geodetic: cpp/include/c_api/geodetic.h:9-59 (coordinate helper of the Python layer; not on the path: NULL / -3) ---- */
Geodetic* sk_geodetic_create(double equatorial_radius, double flattening_factor);
void sk_geodetic_destroy(Geodetic* geodetic);
int sk_geodetic_get_altitude(const Geodetic* geodetic, double* altitude);
int sk_geodetic_get_latitude(const Geodetic* geodetic, double* latitude);
int sk_geodetic_get_longitude(const Geodetic* geodetic, double* longitude);
int sk_geodetic_get_location(const Geodetic* geodetic, double* x, double* y, double* z);
int sk_geodetic_get_local_south(const Geodetic* geodetic, double* x, double* y, double* z);
int sk_geodetic_get_local_up(const Geodetic* geodetic, double* x, double* y, double* z);
int sk_geodetic_get_local_west(const Geodetic* geodetic, double* x, double* y, double* z);
int sk_geodetic_get_altitude_intercepts(const Geodetic* geodetic, double altitude, double observer_x, double observer_y,
                                        double observer_z, double look_vector_x, double look_vector_y,
                                        double look_vector_z, double* x1, double* y1, double* z1, double* x2,
                                        double* y2, double* z2);
int sk_geodetic_from_lat_lon_altitude(const Geodetic* geodetic, double latitude, double longitude, double altitude);
int sk_geodetic_from_tangent_altitude(const Geodetic* geodetic, double altitude, double observer_x, double observer_y,
                                      double observer_z, double boresight_x, double boresight_y, double boresight_z,
                                      double* look_vector_x, double* look_vector_y, double* look_vector_z);
int sk_geodetic_from_tangent_point(const Geodetic* geodetic, double observer_x, double observer_y, double observer_z,
                                   double look_vector_x, double look_vector_y, double look_vector_z);
int sk_geodetic_from_xyz(const Geodetic* geodetic, double x, double y, double z);
int sk_geodetic_is_valid(const Geodetic* geodetic, int* is_valid);
int sk_geodetic_get_osculating_spheroid(const Geodetic* geodetic, double* radius, double* offset_x, double* offset_y,
                                        double* offset_z);

/* ---- cpp/include/c_api/sk_lapack.h: LAPACK dgesv semantics (the Rust Mie / optical code solves small systems with it) ---- */
long long sk_lapack_dgesv(long long n, long long nrhs, double* a, long long lda, long long* ipiv, double* b, long long ldb);

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
/* Wavelength-sharded runs (one process per GPU, contiguous wavelength blocks, cpp/lib/engine/engine.cpp:610-622 upstream
 * shards the same way over host threads): the only exchange is the final gather of radiances and weighting functions
 * onto one rank over NCCL (NVLink / NVSwitch).  Rank 0 obtains a 128-byte id, the host program hands it to every rank
 * (any side channel), every rank calls sk_b200_comm_init on its device.  After sk_b200_engine_solve_staged on every
 * rank, sk_b200_engine_gather_output (collective) leaves the full-spectrum results in `root_output` on rank `root`
 * (radiance [nw_total * nlos], derivative memory assigned for the same mapping names as the staged outputs);
 * block_start / block_count [world] give every rank's wavelength block; ms_out[2] = {NCCL exchange, D2H} on this rank. */
int sk_b200_comm_unique_id(char* id, int nbytes);
int sk_b200_comm_init(const char* id, int rank, int world);
int sk_b200_comm_destroy();
int sk_b200_engine_gather_output(Engine* engine, OutputC* root_output, int root, const int* block_start,
                                 const int* block_count, int nw_total, double* ms_out);
/* Host-only check of the spherical ("limb") geometry plan that sk_engine_create builds for geometrytype::spherical
 * (replaces cpp/lib/raytracing/spherical_shell.cpp + the geometry_interpolator of
 * cpp/lib/sktran_disco/source_term/do_source_diffuse_storage.cpp:84-209 + cpp/lib/solar/solartransmissionexact.cpp:36-96):
 * per ray 6 doubles [line-of-sight optical depth for `ext` [nloc], segments, sum of DO interpolation weights,
 * cos(single-scattering angle), solar optical depth at the far end, at the near end (-1: blocked by the ground)]. */
int sk_b200_limb_plan_check(Geometry1D* geometry, ViewingGeometry* viewing, int nstr, int num_sza, const double* ext,
                            double* per_ray, int* num_points, double* sza_grid);
/* DFMA micro-benchmark on the current device: the FP64 roofline denominator (TFLOP/s) */
double sk_b200_measure_fp64_tflops();
/* 1 when the adjoint boundary-value solve of an (N = num_streams / 2, nlos) problem reuses the forward LU factors
 * (transposed solves, the reference's dgbtrs('T')), 0 when it factorises A^T a second time. */
int sk_b200_adjoint_reuses_factors(int n_half_streams, int nlos);
/* page-locked host memory for caller-side buffers (falls back to malloc without a CUDA device) */
void* sk_b200_host_alloc(size_t nbytes);
void sk_b200_host_free(void* p);
/* page-lock / release caller-owned host memory in place (cudaHostRegister, portable) */
int sk_b200_host_register(void* p, size_t nbytes);
int sk_b200_host_unregister(void* p);
/* test/debug: copy a named workspace array of the last solved chunk (names: see disco_engine.cu) */
long long sk_b200_engine_debug_copy(Engine* engine, const char* name, double* host, long long max_n);

#ifdef __cplusplus
}
#endif
#endif
