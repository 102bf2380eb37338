// C ABI of the B200 discrete-ordinates engine: the reference's `sk_*` entry points for this path
// (cpp/include/c_api/*.h, implemented upstream in cpp/c_api/*.cpp) re-pointed at the CUDA solver.
#include "../../include/sasktran2_b200.h"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <thread>
#include <set>
#include <cstdlib>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "c_api_types.h"

namespace {
thread_local std::string g_last_error;
int g_log_level = 3;  // spdlog::level::warn (cpp/lib/config/config.cpp:21)
std::mutex g_pin_mu;
std::set<void*> g_pinned_ptrs;
}  // namespace

namespace skapi {
int fail(int code, const std::string& msg) {
    g_last_error = msg;
    if (g_log_level <= 4) std::fprintf(stderr, "[sasktran2_b200] %s\n", msg.c_str());
    return code;
}
void* host_alloc(size_t nbytes) {
    void* p = nullptr;
    if (nbytes == 0) nbytes = 8;
    if (cudaMallocHost(&p, nbytes) == cudaSuccess && p) {
        std::lock_guard<std::mutex> lk(g_pin_mu);
        g_pinned_ptrs.insert(p);
        return p;
    }
    (void)cudaGetLastError();
    p = std::malloc(nbytes);
    if (!p) throw std::bad_alloc();
    return p;
}
void host_free(void* p) {
    if (!p) return;
    bool pinned = false;
    {
        std::lock_guard<std::mutex> lk(g_pin_mu);
        pinned = g_pinned_ptrs.erase(p) > 0;
    }
    if (pinned)
        cudaFreeHost(p);
    else
        std::free(p);
}
}  // namespace skapi
using skapi::fail;
using skapi::host_alloc;
using skapi::host_free;


namespace {

disco::AtmosphereArrays arrays_of(const Engine* e, const Atmosphere* atm) {
    disco::AtmosphereArrays a;
    const AtmosphereStorage* s = atm->storage;
    a.nloc = s->nloc;
    a.nwavel = s->nwavel;
    a.nleg = s->nleg;
    a.ssa = s->ssa;
    a.ext = s->ext;
    a.leg = s->leg;
    a.solar = s->solar;
    a.f = s->applied_f_order > 0 ? s->f.data() : nullptr;
    const Surface* sf = atm->surface;
    a.albedo = sf->brdf_args ? sf->brdf_args : sf->default_albedo.data();
    if (sf->brdf && sf->brdf->kind != 0) {   // kernel-based model: brdf_args is [nargs, nwavel] (Surface::brdf_args upstream)
        a.brdf_kind = sf->brdf->kind;
        a.brdf_nargs = disco::brdf_num_args(sf->brdf->kind);
        a.brdf_args = sf->brdf_args;
        a.albedo = nullptr;
    }
    // thermal sources.  emission_source = discrete_ordinates (2) puts the storage's emission_source array into the DO
    // solve (sktran_do_layerarray.cpp:309-311); the surface emission enters the ground boundary whatever the emission
    // source is (sktran_do_rte.h:229-235) - it is passed on only when some wavelength has a non-zero value
    if (e->cfg.emission_source == 2) a.emission = s->emission;
    if (sf->emission)
        for (int w = 0; w < sf->nwavel; ++w)
            if (sf->emission[w] != 0.0) {
                a.surface_emission = sf->emission;
                break;
            }
    return a;
}

// sasktran2::Sasktran2::validate_input_atmosphere (cpp/lib/engine/engine.cpp:481-540), the parts that
// concern this path
int validate(const Engine* e, const Atmosphere* atm, const OutputC* out, bool check_output) {
    if (!atm || !atm->storage || !atm->surface) return fail(-1, "atmosphere handle is null or incomplete");
    const AtmosphereStorage* s = atm->storage;
    if (s->nstokes != 1) return fail(-2, "B200 DO path supports num_stokes = 1 only");
    if (s->nloc != (int)e->geometry->spec.altitudes.size())
        return fail(-2, "atmosphere storage and geometry have a different number of grid points");
    if (!s->ssa || !s->ext || !s->leg || !s->solar) return fail(-1, "atmosphere storage arrays are null");
    if (s->nleg < 1) return fail(-2, "atmosphere storage needs at least one phase moment");
    if (atm->surface->nwavel != s->nwavel) return fail(-2, "surface and storage have a different number of wavelengths");
    if (atm->surface->brdf && atm->surface->brdf->kind != 0 && !atm->surface->brdf_args)
        return fail(-1, "surface BRDF arguments are null");
    if (atm->surface->brdf && atm->surface->brdf->kind == 1 && atm->calc_derivs && e->cfg.wf_enabled && check_output && out &&
        !out->surface_derivs.empty())
        return fail(-2, "B200 DO path: weighting functions w.r.t. the argument of the snow BRDF are not supported");
    if (e->cfg.emission_source == 2 && !s->emission)
        return fail(-1, "emission_source is DiscreteOrdinates but the atmosphere storage has no emission_source array");
    if (e->cfg.emission_source == 2 && atm->calc_derivs && e->cfg.wf_enabled && check_output && out &&
        !(out->derivs.empty() && out->surface_derivs.empty()))
        return fail(-2, "B200 DO path: weighting functions with thermal emission are not supported");
    if (check_output) {
        if (!out || !out->radiance) return fail(-1, "output handle is null");
        if (out->nstokes != 1) return fail(-2, "output num_stokes must be 1");
        const long long need = (long long)s->nwavel * (long long)e->ncols();
        if (out->nrad != need) return fail(-2, "output radiance has the wrong size (expected nwavel * nlos)");
    }
    return 0;
}

// Collect the derivative mappings that have output memory assigned (OutputC::m_derivatives upstream)
int build_wf_request(Engine* e, Atmosphere* atm, OutputC* out, disco::WfRequest& req) {
    if (!out || !(atm->calc_derivs && e->cfg.wf_enabled && e->cfg.wf_precision == 0)) return 0;
    if (out->derivs.empty() && out->surface_derivs.empty()) return 0;
    AtmosphereStorage* s = atm->storage;
    const int nlos = e->ncols();
    const long long nrad = (long long)s->nwavel * nlos;
    sk_atmosphere_storage_finalize_scattering_derivatives(s);
    if (s->num_scat_groups > 2)  // the weighting-function kernels are instantiated for 0, 1 and 2 groups
        return fail(-2, "B200 DO path: at most 2 scattering derivative groups per call (got " + std::to_string(s->num_scat_groups) + ")");
    req.d_legendre.assign(s->num_scat_groups, nullptr);
    for (auto& kv : s->mappings)
        if (kv.second.is_scattering()) req.d_legendre[kv.second.scat_deriv_index] = kv.second.d_legendre.data();
    if (s->applied_f_order > 0) {
        req.d_f.assign(s->num_scat_groups, nullptr);
        for (int g = 0; g < s->num_scat_groups && g < (int)s->d_f.size(); ++g) req.d_f[g] = s->d_f[g].data();
    }
    for (auto& kv : out->derivs) {
        auto it = s->mappings.find(kv.first);
        if (it == s->mappings.end()) return fail(-2, "derivative memory assigned for unknown mapping '" + kv.first + "'");
        MappingImpl& mi = it->second;
        if (kv.second.nrad != nrad || kv.second.nstokes != 1 || kv.second.nderiv != mi.num_output())
            return fail(-2, "derivative memory of mapping '" + kv.first + "' has the wrong shape");
        DerivativeMapping tmp{&mi};
        double* dummy;
        sk_deriv_mapping_get_d_ssa(&tmp, &dummy);         // allocate (zeros) if the caller never touched them
        sk_deriv_mapping_get_d_extinction(&tmp, &dummy);
        disco::WfMapping m;
        m.d_ssa = mi.d_ssa.data();
        m.d_extinction = mi.d_extinction.data();
        m.scat_factor = mi.is_scattering() ? mi.scat_factor.data() : nullptr;
        m.scat_index = mi.scat_deriv_index;
        m.interpolator = mi.interpolator.empty() ? nullptr : mi.interpolator.data();
        m.nout = mi.num_output();
        m.log_radiance_space = mi.log_radiance_space;
        m.out = kv.second.ptr;
        req.mappings.push_back(m);
    }
    for (auto& kv : out->surface_derivs) {
        auto it = atm->surface->mappings.find(kv.first);
        if (it == atm->surface->mappings.end()) return fail(-2, "surface derivative memory assigned for unknown mapping '" + kv.first + "'");
        if (kv.second.nrad != nrad || kv.second.nstokes != 1) return fail(-2, "surface derivative memory has the wrong shape");
        SurfaceDerivativeMapping tmp{&it->second};
        double* dummy;
        sk_surface_deriv_mapping_get_d_brdf(&tmp, &dummy);
        disco::WfSurface sf;
        sf.d_brdf = it->second.d_brdf.data();
        sf.nargs = it->second.nargs;
        sf.nwavel = (size_t)it->second.nwavel;
        sf.out = kv.second.ptr;
        req.surfaces.push_back(sf);
    }
    return 0;
}

int run_range(Engine* e, Atmosphere* atm, OutputC* out, int start, int count) {
    try {
        disco::WfRequest req;
        int rc = build_wf_request(e, atm, out, req);
        if (rc != 0) return rc;
        const int nlos = e->ncols();
        e->dev->calculate(arrays_of(e, atm), start, count, out->radiance + (size_t)start * nlos, req.enabled() ? &req : nullptr);
        e->staged_start = start;
        e->staged_count = count;
        if (e->dev->limb() && e->cfg.output_los_optical_depth) {
            // Output::m_los_optical_depth is (nwavel, nlos) column-major upstream (cpp/lib/output/output.cpp:57-60)
            const size_t nw_total = (size_t)atm->storage->nwavel;
            out->los_optical_depth.resize(nw_total * nlos, 0.0);
            std::vector<double> tmp((size_t)count * nlos);
            e->dev->fetch_los_optical_depth(tmp.data());
            for (int w = 0; w < count; ++w)
                for (int r = 0; r < nlos; ++r) out->los_optical_depth[(size_t)(start + w) + nw_total * r] = tmp[(size_t)w * nlos + r];
        }
        return 0;
    } catch (const std::exception& ex) {
        return fail(-3, ex.what());
    }
}

}  // namespace

extern "C" {

// ---------------------------------------------------------------------------------------------------
// Config
// ---------------------------------------------------------------------------------------------------
Config* sk_config_create() { return new Config(); }
void sk_config_destroy(Config* c) { delete c; }

#define CFG_GETSET(name, field)                                   \
    int sk_config_get_##name(Config* c, int* v) {                 \
        if (!c || !v) return -1;                                  \
        *v = c->field;                                            \
        return 0;                                                 \
    }                                                             \
    int sk_config_set_##name(Config* c, int v) {                  \
        if (!c) return -1;                                        \
        c->field = v;                                             \
        return 0;                                                 \
    }
CFG_GETSET(num_stokes, num_stokes)
CFG_GETSET(multiple_scatter_source, multiple_scatter_source)
CFG_GETSET(single_scatter_source, single_scatter_source)
CFG_GETSET(num_streams, num_streams)
CFG_GETSET(num_threads, num_threads)
CFG_GETSET(threading_model, threading_model)
CFG_GETSET(wavelength_batch_size, wavelength_batch_size)
CFG_GETSET(num_singlescatter_moments, num_singlescatter_moments)
CFG_GETSET(apply_delta_scaling, apply_delta_scaling)
CFG_GETSET(num_do_sza, num_do_sza)
CFG_GETSET(num_do_forced_azimuth, num_do_forced_azimuth)
CFG_GETSET(do_backprop, do_backprop)
CFG_GETSET(emission_source, emission_source)
CFG_GETSET(occultation_source, occultation_source)
CFG_GETSET(solar_refraction, solar_refraction)
CFG_GETSET(wf_enabled, wf_enabled)
CFG_GETSET(wf_precision, wf_precision)
CFG_GETSET(input_validation_mode, input_validation_mode)
int sk_config_get_log_level(Config* c, int* v) {
    if (!c || !v) return -1;
    *v = c->log_level;
    return 0;
}
int sk_config_set_log_level(Config* c, int v) {
    if (!c) return -1;
    c->log_level = v;
    g_log_level = v;
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// Geometry / viewing geometry
// ---------------------------------------------------------------------------------------------------
Geometry1D* sk_geometry1d_create(double cos_sza, double saa, double earth_radius, double* grid_values, int ngrid_values,
                                 int interp_method, int geotype) {
    if (!grid_values || ngrid_values < 2) {
        fail(-2, "sk_geometry1d_create: need at least two grid values");
        return nullptr;
    }
    auto* g = new Geometry1D();
    g->spec.altitudes.assign(grid_values, grid_values + ngrid_values);
    g->spec.interp = interp_method;
    g->spec.geotype = geotype;
    g->spec.cos_sza = cos_sza;
    g->spec.saa = saa;
    g->spec.earth_radius = earth_radius;
    return g;
}
void sk_geometry1d_destroy(Geometry1D* g) { delete g; }
int sk_geometry1d_get_num_altitudes(const Geometry1D* g) { return g ? (int)g->spec.altitudes.size() : -1; }
int sk_geometry1d_get_altitudes(const Geometry1D* g, double* altitudes) {
    if (!g || !altitudes) return -1;
    std::copy(g->spec.altitudes.begin(), g->spec.altitudes.end(), altitudes);
    return 0;
}

ViewingGeometry* sk_viewing_geometry_create() { return new ViewingGeometry(); }
void sk_viewing_geometry_destroy(ViewingGeometry* v) { delete v; }
void sk_viewing_geometry_add_ground_viewing_solar(ViewingGeometry* v, double cos_sza, double relative_azimuth_angle,
                                                  double observeraltitude, double cos_viewing_zenith) {
    if (!v) return;
    v->rays.push_back({cos_viewing_zenith, relative_azimuth_angle, observeraltitude});
    v->ray_cos_sza.push_back(cos_sza);
    v->ordered.push_back({0, {cos_sza, relative_azimuth_angle, cos_viewing_zenith, observeraltitude}});
}
int sk_viewing_geometry_num_rays(ViewingGeometry* v, int* num_rays) {
    if (!v || !num_rays) return -1;
    *num_rays = (int)(v->rays.size() + v->num_tangent_rays + v->other_rays.size());
    return 0;
}
int sk_viewing_geometry_num_flux_observers(ViewingGeometry* v, int* n) {
    if (!v || !n) return -1;
    *n = v->num_flux_observers;
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// Atmosphere storage, mappings, surface
// ---------------------------------------------------------------------------------------------------
AtmosphereStorage* sk_atmosphere_storage_create(int nlocation, int nwavel, int nphase_moments, int nstokes, double* ssa,
                                                double* total_extinction, double* emission_source, double* leg_coeff,
                                                double* solar_irradiance) {
    if (nstokes != 1 && nstokes != 3) {
        fail(-2, "sk_atmosphere_storage_create: nstokes must be 1 or 3");
        return nullptr;
    }
    auto* s = new AtmosphereStorage();
    s->nloc = nlocation;
    s->nwavel = nwavel;
    s->nleg = nphase_moments;
    s->nstokes = nstokes;
    s->ssa = ssa;
    s->ext = total_extinction;
    s->emission = emission_source;
    s->leg = leg_coeff;
    s->solar = solar_irradiance;
    return s;
}
void sk_atmosphere_storage_destroy(AtmosphereStorage* s) { delete s; }

int sk_atmosphere_storage_get_derivative_mapping(AtmosphereStorage* s, const char* name, DerivativeMapping** mapping) {
    if (!s || !name || !mapping) return -1;
    auto it = s->mappings.find(name);
    if (it == s->mappings.end()) {
        MappingImpl m;
        m.nwavel = s->nwavel;
        m.nloc = s->nloc;
        m.nleg = s->nleg;
        m.assign_name = name;
        it = s->mappings.emplace(name, std::move(m)).first;
    }
    *mapping = new DerivativeMapping{&it->second};
    return 0;
}
int sk_atmosphere_storage_get_num_derivative_mappings(AtmosphereStorage* s, int* n) {
    if (!s || !n) return -1;
    *n = (int)s->mappings.size();
    return 0;
}
int sk_atmosphere_storage_get_derivative_mapping_name(AtmosphereStorage* s, int index, const char** name) {
    if (!s || !name) return -1;
    if (index < 0 || index >= (int)s->mappings.size()) return -2;
    auto it = s->mappings.begin();
    std::advance(it, index);
    *name = it->first.c_str();
    return 0;
}
int sk_atmosphere_storage_get_derivative_mapping_by_index(AtmosphereStorage* s, int index, DerivativeMapping** mapping) {
    if (!s || !mapping) return -1;
    if (index < 0 || index >= (int)s->mappings.size()) return -2;
    auto it = s->mappings.begin();
    std::advance(it, index);
    *mapping = new DerivativeMapping{&it->second};
    return 0;
}
// AtmosphereGridStorageFull::finalize_scattering_derivatives (cpp/include/sasktran2/atmosphere/grid_storage.h:186-209):
// scattering mappings get consecutive group indices in name order
int sk_atmosphere_storage_finalize_scattering_derivatives(AtmosphereStorage* s) {
    if (!s) return -1;
    int idx = 0;
    for (auto& kv : s->mappings)
        if (kv.second.is_scattering()) kv.second.scat_deriv_index = idx++;
    s->num_scat_groups = idx;
    return 0;
}
int sk_atmosphere_storage_set_zero(AtmosphereStorage* s) {
    if (!s) return -1;
    const size_t n = (size_t)s->nloc * s->nwavel;
    if (s->ssa) std::fill(s->ssa, s->ssa + n, 0.0);
    if (s->ext) std::fill(s->ext, s->ext + n, 0.0);
    if (s->emission) std::fill(s->emission, s->emission + n, 0.0);
    if (s->leg) std::fill(s->leg, s->leg + n * s->nleg * (s->nstokes == 3 ? 4 : 1), 0.0);
    // AtmosphereGridStorageFull::set_zero zeroes f (grid_storage.h:319-329): the storage is unscaled again
    s->applied_f_order = 0;
    s->f.clear();
    s->d_f.clear();
    for (auto& kv : s->mappings) {
        DerivativeMapping tmp{&kv.second};
        sk_deriv_mapping_set_zero(&tmp);
    }
    return 0;
}

int sk_deriv_mapping_destroy(DerivativeMapping* m) {
    if (!m) return -1;
    delete m;  // the mapping itself is owned by the storage (cpp/c_api/deriv_mapping.cpp:7-14)
    return 0;
}
int sk_deriv_mapping_set_zero(DerivativeMapping* m) {
    if (!m) return -1;
    for (auto* v : {&m->impl->d_ssa, &m->impl->d_extinction, &m->impl->scat_factor, &m->impl->d_legendre})
        std::fill(v->begin(), v->end(), 0.0);
    return 0;
}
int sk_deriv_mapping_get_d_ssa(DerivativeMapping* m, double** p) {
    if (!m || !p) return -1;
    if (!m->impl->has_d_ssa) {
        m->impl->d_ssa.assign((size_t)m->impl->nloc * m->impl->nwavel, 0.0);
        m->impl->has_d_ssa = true;
    }
    *p = m->impl->d_ssa.data();
    return 0;
}
int sk_deriv_mapping_get_d_extinction(DerivativeMapping* m, double** p) {
    if (!m || !p) return -1;
    if (!m->impl->has_d_extinction) {
        m->impl->d_extinction.assign((size_t)m->impl->nloc * m->impl->nwavel, 0.0);
        m->impl->has_d_extinction = true;
    }
    *p = m->impl->d_extinction.data();
    return 0;
}
static void alloc_legendre(MappingImpl* mi) {
    if (!mi->has_legendre) {
        mi->d_legendre.assign((size_t)mi->nleg * mi->nloc * mi->nwavel, 0.0);
        mi->scat_factor.assign((size_t)mi->nloc * mi->nwavel, 0.0);
        mi->has_legendre = true;
    }
}
int sk_deriv_mapping_get_scat_factor(DerivativeMapping* m, double** p) {
    if (!m || !p) return -1;
    alloc_legendre(m->impl);
    *p = m->impl->scat_factor.data();
    return 0;
}
int sk_deriv_mapping_get_d_legendre(DerivativeMapping* m, double** p) {
    if (!m || !p) return -1;
    alloc_legendre(m->impl);
    *p = m->impl->d_legendre.data();
    return 0;
}
int sk_deriv_mapping_get_scat_deriv_index(DerivativeMapping* m, int* v) {
    if (!m || !v) return -1;
    *v = m->impl->scat_deriv_index;
    return 0;
}
int sk_deriv_mapping_set_scat_deriv_index(DerivativeMapping* m, int v) {
    if (!m) return -1;
    m->impl->scat_deriv_index = v;
    return 0;
}
int sk_deriv_mapping_get_num_location(DerivativeMapping* m, int* v) {
    if (!m || !v) return -1;
    *v = m->impl->nloc;
    return 0;
}
int sk_deriv_mapping_get_num_wavel(DerivativeMapping* m, int* v) {
    if (!m || !v) return -1;
    *v = m->impl->nwavel;
    return 0;
}
int sk_deriv_mapping_get_num_legendre(DerivativeMapping* m, int* v) {
    if (!m || !v) return -1;
    *v = m->impl->nleg;
    return 0;
}
int sk_deriv_mapping_set_interp_dim(DerivativeMapping* m, const char* name) {
    if (!m || !name) return -1;
    m->impl->interp_dim = name;
    return 0;
}
int sk_deriv_mapping_set_assign_name(DerivativeMapping* m, const char* name) {
    if (!m || !name) return -1;
    m->impl->assign_name = name;
    return 0;
}
int sk_deriv_mapping_set_log_radiance_space(DerivativeMapping* m, int v) {
    if (!m) return -1;
    m->impl->log_radiance_space = v != 0;
    return 0;
}
int sk_deriv_mapping_get_log_radiance_space(DerivativeMapping* m, int* v) {
    if (!m || !m->impl || !v) return -1;
    *v = m->impl->log_radiance_space ? 1 : 0;
    return 0;
}
int sk_deriv_mapping_is_scattering_derivative(DerivativeMapping* m, int* v) {
    if (!m || !v) return -1;
    *v = m->impl->is_scattering() ? 1 : 0;
    return 0;
}
int sk_deriv_mapping_get_num_output(DerivativeMapping* m, int* v) {
    if (!m || !v) return -1;
    *v = m->impl->num_output();
    return 0;
}
int sk_deriv_mapping_get_assign_name(DerivativeMapping* m, const char** name) {
    if (!m || !name) return -1;
    *name = m->impl->assign_name.c_str();
    return 0;
}
int sk_deriv_mapping_get_interp_dim(DerivativeMapping* m, const char** name) {
    if (!m || !name) return -1;
    *name = m->impl->interp_dim.c_str();
    return 0;
}
int sk_deriv_mapping_set_interpolator(DerivativeMapping* m, double* interpolator, int dim1, int dim2) {
    if (!m || !interpolator) return -1;
    if (dim1 != m->impl->nloc) return -2;
    m->impl->interpolator.assign(interpolator, interpolator + (size_t)dim1 * dim2);
    m->impl->interp_d1 = dim1;
    m->impl->interp_d2 = dim2;
    return 0;
}
int sk_deriv_mapping_clear_interpolator(DerivativeMapping* m) {
    if (!m) return -1;
    m->impl->interpolator.clear();
    m->impl->interp_d1 = m->impl->interp_d2 = 0;
    return 0;
}
int sk_deriv_mapping_get_interpolator(DerivativeMapping* m, double** interpolator, int* dim1, int* dim2) {
    if (!m || !interpolator || !dim1 || !dim2) return -1;
    *interpolator = m->impl->interpolator.empty() ? nullptr : m->impl->interpolator.data();
    *dim1 = m->impl->interp_d1;
    *dim2 = m->impl->interp_d2;
    return 0;
}

Atmosphere* sk_atmosphere_create(AtmosphereStorage* storage, Surface* surface, int calculate_derivatives,
                                 int calculate_emission_derivatives) {
    if (!storage || !surface) {
        fail(-1, "sk_atmosphere_create: null storage or surface");
        return nullptr;
    }
    auto* a = new Atmosphere();
    a->storage = storage;
    a->surface = surface;
    a->calc_derivs = calculate_derivatives != 0;
    a->calc_emission_derivs = calculate_emission_derivatives != 0;
    return a;
}
void sk_atmosphere_destroy(Atmosphere* a) { delete a; }
// Atmosphere::apply_delta_m_scaling (cpp/lib/atmosphere/atmosphere.cpp:69-203): in-place delta-M scaling of the
// caller's storage arrays and of the derivative mappings, order = number of streams.  It is a host pre-pass over
// caller-owned host memory upstream and stays one here (threaded over wavelengths); the solve reads the truncation
// fraction f and its derivatives d_f (layer optics and Legendre derivative directions,
// sktran_do_layerarray.cpp:396-410, 773-800).
int sk_atmosphere_apply_delta_m_scaling(Atmosphere* a, int order) {
    if (!a || !a->storage) return -1;
    AtmosphereStorage* s = a->storage;
    if (s->nstokes != 1) return fail(-2, "B200 DO path supports num_stokes = 1 only");
    if (order < 0) return fail(-2, "delta-M order must be non-negative");
    if (order >= s->nleg) return 0;  // upstream warns and leaves the atmosphere unscaled
    // like upstream, every call rescales whatever the storage holds now: the caller refills the arrays
    // (sk_atmosphere_storage_set_zero + fill) before each application (src/sasktran2/atmosphere.py:655-662, 846-856)
    if (!s->ssa || !s->ext || !s->leg) return fail(-1, "atmosphere storage arrays are null");
    sk_atmosphere_storage_finalize_scattering_derivatives(s);
    const size_t nloc = s->nloc, nw = s->nwavel, nleg = s->nleg;
    const int G = s->num_scat_groups;
    s->f.assign(nloc * nw, 0.0);
    s->d_f.assign(G, PinnedVec());
    for (auto& v : s->d_f) v.assign(nloc * nw, 0.0);
    std::vector<MappingImpl*> maps, scat_of_group(G, nullptr);
    for (auto& kv : s->mappings) {
        maps.push_back(&kv.second);
        if (kv.second.is_scattering() && kv.second.scat_deriv_index >= 0) scat_of_group[kv.second.scat_deriv_index] = &kv.second;
    }
    const double inv = 1.0 / (2.0 * order + 1.0);
    auto pass = [&](size_t w0, size_t w1) {
        for (size_t w = w0; w < w1; ++w) {
            for (size_t i = 0; i < nloc; ++i) {
                const size_t iw = i + nloc * w;
                double* leg = s->leg + nleg * iw;
                const double f = leg[order] * inv;
                const double ssa0 = s->ssa[iw], ext0 = s->ext[iw];
                s->f[iw] = f;
                const double one_m_wf = 1.0 - ssa0 * f;
                s->ext[iw] = ext0 * one_m_wf;                  // k* = (1 - w f) k
                const double ssa1 = (1.0 - f) / one_m_wf * ssa0;   // w* = (1 - f) / (1 - w f) w
                s->ssa[iw] = ssa1;
                for (int g = 0; g < G; ++g)
                    s->d_f[g][iw] = scat_of_group[g] ? scat_of_group[g]->d_legendre[order + nleg * iw] * inv : 0.0;
                for (size_t j = 0; j < nleg; ++j) {
                    leg[j] /= (1.0 - f);                       // b* = b / (1 - f)
                    for (int g = 0; g < G; ++g) {
                        if (!scat_of_group[g]) continue;
                        double& db = scat_of_group[g]->d_legendre[j + nleg * iw];
                        db += leg[j] * s->d_f[g][iw];          // db* = (db + b* df) / (1 - f)
                        db /= (1.0 - f);
                    }
                }
                for (MappingImpl* m : maps) {
                    if (!m->has_d_extinction) continue;
                    if (!m->has_d_ssa) continue;  // allocated below before the pass
                    double& dk = m->d_extinction[iw];
                    double& dw = m->d_ssa[iw];
                    dk *= one_m_wf;
                    dk -= ext0 * f * dw;
                    dw *= (1.0 - f * (1.0 - ssa1)) / one_m_wf;
                    if (m->is_scattering() && m->scat_deriv_index >= 0) {
                        const double df = s->d_f[m->scat_deriv_index][iw] * m->scat_factor[iw];
                        dk -= ssa0 * ext0 * df;
                        dw += df * ssa0 / one_m_wf * (ssa1 - 1.0);
                    }
                }
            }
        }
    };
    for (MappingImpl* m : maps)
        if (m->has_d_extinction && !m->has_d_ssa) {
            m->d_ssa.assign(nloc * nw, 0.0);
            m->has_d_ssa = true;
        }
    unsigned nt = std::thread::hardware_concurrency();
    nt = std::max(1u, std::min<unsigned>(nt, (unsigned)std::max<size_t>(1, nw / 64)));
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < nt; ++t) pool.emplace_back(pass, nw * t / nt, nw * (t + 1) / nt);
    for (auto& th : pool) th.join();
    s->applied_f_order = order;
    return 0;
}

Surface* sk_surface_create(int nwavel, int nstokes, double* emission) {
    auto* s = new Surface();
    s->nwavel = nwavel;
    s->nstokes = nstokes;
    s->emission = emission;
    s->default_albedo.assign(nwavel > 0 ? nwavel : 0, 0.0);
    return s;
}
void sk_surface_destroy(Surface* s) { delete s; }
int sk_surface_set_brdf(Surface* s, BRDF* brdf, double* brdf_args) {
    if (!s || !brdf) return -1;
    s->brdf = brdf;
    s->brdf_args = brdf_args;
    return 0;
}
int sk_surface_get_derivative_mapping(Surface* s, const char* name, SurfaceDerivativeMapping** mapping) {
    if (!s || !name || !mapping) return -1;
    auto it = s->mappings.find(name);
    if (it == s->mappings.end()) {
        SurfaceMappingImpl m;
        m.nwavel = s->nwavel;
        m.nargs = (s->brdf && s->brdf->kind != 0) ? disco::brdf_num_args(s->brdf->kind) : 1;   // d_brdf is [nwavel, num_args] upstream
        it = s->mappings.emplace(name, std::move(m)).first;
    }
    *mapping = new SurfaceDerivativeMapping{&it->second};
    return 0;
}
int sk_surface_get_num_derivative_mappings(Surface* s, int* n) {
    if (!s || !n) return -1;
    *n = (int)s->mappings.size();
    return 0;
}
int sk_surface_get_derivative_mapping_name(Surface* s, int index, const char** name) {
    if (!s || !name) return -1;
    if (index < 0 || index >= (int)s->mappings.size()) return -2;
    auto it = s->mappings.begin();
    std::advance(it, index);
    *name = it->first.c_str();
    return 0;
}
int sk_surface_set_zero(Surface* s) {
    if (!s) return -1;
    if (s->brdf_args) std::fill(s->brdf_args, s->brdf_args + s->nwavel, 0.0);
    if (s->emission) std::fill(s->emission, s->emission + s->nwavel, 0.0);
    for (auto& kv : s->mappings) std::fill(kv.second.d_brdf.begin(), kv.second.d_brdf.end(), 0.0);
    return 0;
}
int sk_surface_deriv_mapping_get_num_wavel(SurfaceDerivativeMapping* m, int* v) {
    if (!m || !v) return -1;
    *v = m->impl->nwavel;
    return 0;
}
int sk_surface_deriv_mapping_get_num_brdf_args(SurfaceDerivativeMapping* m, int* v) {
    if (!m || !v) return -1;
    *v = m->impl->nargs;
    return 0;
}
int sk_surface_deriv_mapping_get_d_brdf(SurfaceDerivativeMapping* m, double** p) {
    if (!m || !p) return -1;
    if (!m->impl->has_d_brdf) {
        m->impl->d_brdf.assign((size_t)m->impl->nwavel * m->impl->nargs, 0.0);
        m->impl->has_d_brdf = true;
    }
    *p = m->impl->d_brdf.data();
    return 0;
}
int sk_surface_deriv_mapping_set_zero(SurfaceDerivativeMapping* m) {
    if (!m) return -1;
    std::fill(m->impl->d_brdf.begin(), m->impl->d_brdf.end(), 0.0);
    return 0;
}
int sk_surface_deriv_mapping_destroy(SurfaceDerivativeMapping* m) {
    if (!m) return -1;
    delete m;
    return 0;
}

BRDF* sk_brdf_create_lambertian(int nstokes) {
    auto* b = new BRDF();
    b->kind = 0;
    b->nstokes = nstokes;
    return b;
}
// cpp/include/sasktran2/atmosphere/surface.h: Lambertian 1 / 1, SnowKokhanovsky 1 / 1, MODIS 3 / 3
int sk_brdf_get_num_deriv(BRDF* b, int* n) {
    if (!b || !n) return -1;
    *n = disco::brdf_num_args(b->kind);
    return 0;
}
int sk_brdf_get_num_args(BRDF* b, int* n) {
    if (!b || !n) return -1;
    *n = disco::brdf_num_args(b->kind);
    return 0;
}
void sk_brdf_destroy(BRDF* b) { delete b; }

// ---------------------------------------------------------------------------------------------------
// Output
// ---------------------------------------------------------------------------------------------------
OutputC* sk_output_create(double* radiance, int nrad, int nstokes, double* flux, int nflux) {
    auto* o = new OutputC();
    o->radiance = radiance;
    o->nrad = nrad;
    o->nstokes = nstokes;
    o->flux = flux;
    o->nflux = nflux;
    return o;
}
void sk_output_destroy(OutputC* o) { delete o; }
int sk_output_assign_derivative_memory(OutputC* o, const char* name, double* mem, int nrad, int nstokes, int nderiv) {
    if (!o || !name || !mem) return -1;
    o->derivs[name] = DerivMem{mem, nrad, nstokes, nderiv};
    return 0;
}
int sk_output_assign_surface_derivative_memory(OutputC* o, const char* name, double* mem, int nrad, int nstokes) {
    if (!o || !name || !mem) return -1;
    o->surface_derivs[name] = DerivMem{mem, nrad, nstokes, 1};
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// Engine
// ---------------------------------------------------------------------------------------------------
// Spherical geometry (geometrytype::spherical): traced lines of sight, DOSourceInterpolatedPostProcessing for the
// multiple-scatter source and SingleScatterSource<SolarTransmissionExact> for the single scatter
// (cpp/lib/engine/engine.cpp:78-112, 210-220, 285-365).  SingleScatterSource::DiscreteOrdinates adds no
// line-of-sight term in this geometry, exactly as upstream.
static Engine* create_limb_engine(Config* config, Geometry1D* geometry, ViewingGeometry* viewing) {
    const int ms = config->multiple_scatter_source, ss = config->single_scatter_source;
    if (ms != 0 && ms != 3) {
        fail(-2, "B200 limb path: multiple_scatter_source must be DiscreteOrdinates (0) or NoSource (3)");
        return nullptr;
    }
    if (ss != 0 && ss != 2 && ss != 3) {
        fail(-2, "B200 limb path: single_scatter_source must be Exact (0), DiscreteOrdinates (2) or NoSource (3)");
        return nullptr;
    }
    if (config->emission_source != 1 || config->occultation_source != 1) {
        fail(-2, "B200 limb path: emission and occultation sources are not supported");
        return nullptr;
    }
    if (config->solar_refraction || config->los_refraction || config->multiple_scatter_refraction) {
        fail(-2, "B200 limb path: refraction is not supported");
        return nullptr;
    }
    if (!viewing->other_rays.empty() || viewing->num_flux_observers > 0) {
        fail(-2, "B200 limb path: only GroundViewingSolar and TangentAltitudeSolar rays are supported");
        return nullptr;
    }
    if (config->singlescatter_phasemode != 0) {
        fail(-2, "B200 limb path: the single-scatter phase function is evaluated from the Legendre moments only");
        return nullptr;
    }
    if (ss == 0 && ms == 0 && config->num_singlescatter_moments < config->num_streams) {   // cpp/lib/config/config.cpp:98-107
        fail(-2, "Invalid number of single scatter moments, must be at least the number of streams");
        return nullptr;
    }
    try {
        auto* e = new Engine();
        e->cfg = *config;
        e->geometry = geometry;
        e->viewing = viewing;
        disco::LimbOptions lo;
        lo.num_sza = config->num_do_sza;
        lo.ms_do = ms == 0;
        lo.ss_exact = ss == 0;
        lo.num_ss_moments = config->num_singlescatter_moments;
        disco::LimbPlan limb = disco::build_limb_plan(config->num_streams, geometry->spec, viewing->ordered, lo);
        // stream tables of the DO solves: the pseudo-spherical construction without lines of sight
        disco::GeometrySpec gs = geometry->spec;
        gs.geotype = 1;
        gs.cos_sza = limb.sza_grid.empty() ? gs.cos_sza : limb.sza_grid[0];
        if (!(gs.cos_sza > 0.0)) gs.cos_sza = 1.0;   // single scatter only: the DO tables are never used
        disco::HostPlan plan = disco::build_plan(config->num_streams, gs, {});
        disco::EngineOptions opt;
        opt.nstr = config->num_streams;
        opt.include_ss = false;
        opt.validate_inputs = config->input_validation_mode != 2;
        opt.forced_azimuth = -1;   // DOSource::calculate loops over all num_do_streams orders (do_source.cpp:47-57)
        if (const char* env = std::getenv("SK_B200_WORKSPACE_GB")) opt.workspace_gb = std::atof(env);
        e->dev = std::make_unique<disco::DeviceEngine>(opt, plan, limb);
        return e;
    } catch (const std::exception& ex) {
        fail(-3, ex.what());
        return nullptr;
    }
}

Engine* sk_engine_create(Config* config, Geometry1D* geometry, ViewingGeometry* viewing) {
    if (!config || !geometry || !viewing) {
        fail(-1, "sk_engine_create: null handle");
        return nullptr;
    }
    // Refuse everything outside the CUDA path loudly (there is no CPU fallback)
    if (config->num_stokes != 1) {
        fail(-2, "B200 DO path: num_stokes must be 1");
        return nullptr;
    }
    if (geometry->spec.geotype == 2) return create_limb_engine(config, geometry, viewing);
    if (viewing->num_tangent_rays > 0) {
        fail(-2, "TangentAltitude ray construction can only be used in spherical geometry mode.");
        return nullptr;
    }
    // TwoStream (2): the reference's dedicated two-stream source (cpp_twostream_source.cpp) is, for ground-viewing
    // rays in plane-parallel / pseudo-spherical geometry, the multiple-scatter-only 2-stream discrete-ordinates
    // source - upstream asserts their equality to 2e-8 on radiances and weighting functions
    // (tests/engine/test_twostream.py:104-160).  Here it runs through the DO kernels with N = 1.
    if (config->multiple_scatter_source == 2) {
        if (config->num_streams != 2 || config->single_scatter_source != 3) {
            fail(-2, "B200 DO path: multiple_scatter_source TwoStream needs num_streams = 2 and single_scatter_source None (3)");
            return nullptr;
        }
    } else if (config->multiple_scatter_source != 0) {
        fail(-2, "B200 DO path: multiple_scatter_source must be DiscreteOrdinates (0) or TwoStream (2)");
        return nullptr;
    }
    if (config->single_scatter_source != 2 && config->single_scatter_source != 3) {
        fail(-2, "B200 DO path: single_scatter_source must be DiscreteOrdinates (2) or None (3)");
        return nullptr;
    }
    if (config->emission_source != 1 && config->emission_source != 2) {
        fail(-2, "B200 DO path: emission_source must be None (1) or DiscreteOrdinates (2)");
        return nullptr;
    }
    if (config->emission_source == 2 && config->multiple_scatter_source == 2) {
        fail(-2, "B200 two-stream path: emission sources are not supported");
        return nullptr;
    }
    // Config::validate_config, cpp/lib/config/config.cpp:127-141
    if (config->emission_source == 2 && config->single_scatter_source != 2) {
        fail(-2, "emission_source=discrete_ordinates requires single_scatter_source=discrete_ordinates");
        return nullptr;
    }
    if (config->emission_source == 2 && config->multiple_scatter_source != 0) {
        fail(-2, "emission_source=discrete_ordinates requires multiple_scatter_source=discrete_ordinates");
        return nullptr;
    }
    if (config->solar_refraction) {
        fail(-2, "B200 DO path: solar refraction is not supported");
        return nullptr;
    }
    if (!viewing->other_rays.empty()) {
        fail(-2, "B200 DO path: only GroundViewingSolar rays are supported (tangent-altitude / observer-location rays need "
                 "the spherical source-table path)");
        return nullptr;
    }
    if (viewing->num_flux_observers > 0) {
        fail(-2, "B200 DO path: flux observers are not supported");
        return nullptr;
    }
    if (config->los_refraction || config->multiple_scatter_refraction) {
        fail(-2, "B200 DO path: refraction is not supported");
        return nullptr;
    }
    // apply_delta_scaling is acted on by the caller (src/sasktran2/atmosphere.py:846-856 calls
    // sk_atmosphere_apply_delta_m_scaling with order = num_streams); the engine reads the storage's f / d_f.
    for (size_t i = 0; i < viewing->rays.size(); ++i) {
        if (std::abs(viewing->ray_cos_sza[i] - geometry->spec.cos_sza) > 1e-12) {
            fail(-2, "B200 DO path: every ground-viewing ray must use the geometry's cos_sza");
            return nullptr;
        }
    }
    try {
        auto* e = new Engine();
        e->cfg = *config;
        e->geometry = geometry;
        e->viewing = viewing;
        disco::HostPlan plan = disco::build_plan(config->num_streams, geometry->spec, viewing->rays);
        disco::EngineOptions opt;
        opt.nstr = config->num_streams;
        opt.include_ss = config->single_scatter_source == 2;
        opt.forced_azimuth = config->num_do_forced_azimuth;
        opt.twostream = config->multiple_scatter_source == 2;
        opt.validate_inputs = config->input_validation_mode != 2;   // InputValidationMode::disabled
        if (const char* env = std::getenv("SK_B200_WORKSPACE_GB")) opt.workspace_gb = std::atof(env);
        e->dev = std::make_unique<disco::DeviceEngine>(opt, plan);
        return e;
    } catch (const std::exception& ex) {
        fail(-3, ex.what());
        return nullptr;
    }
}
void sk_engine_destroy(Engine* e) { delete e; }

int sk_engine_calculate_radiance(Engine* e, Atmosphere* atm, OutputC* out, int only_initialize) {
    if (!e || !e->dev) return fail(-1, "engine handle is null");
    std::lock_guard<std::mutex> lock(e->mtx);
    int rc = validate(e, atm, out, true);
    if (rc != 0) return rc;
    e->atmosphere = atm;
    if (only_initialize) return 0;
    return run_range(e, atm, out, 0, atm->storage->nwavel);
}

int sk_engine_calculate_radiance_block_thread(Engine* e, OutputC* out, int wavelength_start, int wavelength_count, int) {
    if (!e || !e->dev) return fail(-1, "engine handle is null");
    std::lock_guard<std::mutex> lock(e->mtx);  // one GPU stream: concurrent host threads are serialised
    if (!e->atmosphere) return fail(-1, "sk_engine_calculate_radiance(only_initialize=1) must be called first");
    int rc = validate(e, e->atmosphere, out, true);
    if (rc != 0) return rc;
    if (wavelength_start < 0 || wavelength_count < 0 || wavelength_start + wavelength_count > e->atmosphere->storage->nwavel)
        return fail(-2, "wavelength block out of range");
    return run_range(e, e->atmosphere, out, wavelength_start, wavelength_count);
}

// The reference collapses to 1 for DO sources (cpp/lib/engine/engine.cpp:869-892); a batched GPU engine wants
// the whole spectrum in one call so that the Rayon scheduler hands over one block
int sk_engine_effective_wavelength_batch_size(Engine* e, int num_wavelengths) {
    if (!e) return -1;
    return num_wavelengths > 0 ? num_wavelengths : 1;
}
int sk_engine_supports_linearization(Engine* e, int, int* supported) {
    if (!e || !supported) return -1;
    *supported = 1;
    return 0;
}
int sk_engine_linearization_backend(Engine* e, int, int* backend) {
    if (!e || !backend) return -1;
    *backend = 0;  // Jacobian-only ("StreamingJacobian" on the Rust side)
    return 0;
}
int sk_openmp_support_enabled() { return 0; }

// ---------------------------------------------------------------------------------------------------
// extensions
// ---------------------------------------------------------------------------------------------------
const char* sk_b200_last_error() { return g_last_error.c_str(); }
int sk_b200_device_count() {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}
int sk_b200_set_device(int device) {
    cudaError_t err = cudaSetDevice(device);
    if (err != cudaSuccess) return fail(-3, std::string("cudaSetDevice: ") + cudaGetErrorString(err));
    return 0;
}
int sk_b200_engine_stage_atmosphere(Engine* e, Atmosphere* atm, OutputC* out, int wavelength_start, int wavelength_count) {
    if (!e || !e->dev) return fail(-1, "engine handle is null");
    std::lock_guard<std::mutex> lock(e->mtx);
    int rc = validate(e, atm, out, out != nullptr);
    if (rc != 0) return rc;
    if (wavelength_count < 0) wavelength_count = atm->storage->nwavel - wavelength_start;
    try {
        disco::WfRequest req;
        rc = build_wf_request(e, atm, out, req);
        if (rc != 0) return rc;
        e->dev->stage(arrays_of(e, atm), wavelength_start, wavelength_count, req.enabled() ? &req : nullptr);
        e->atmosphere = atm;
        e->staged_start = wavelength_start;
        e->staged_count = wavelength_count;
        return 0;
    } catch (const std::exception& ex) {
        return fail(-3, ex.what());
    }
}
int sk_b200_engine_solve_staged(Engine* e) {
    if (!e || !e->dev) return fail(-1, "engine handle is null");
    std::lock_guard<std::mutex> lock(e->mtx);
    try {
        e->dev->solve_staged();
        return 0;
    } catch (const std::exception& ex) {
        return fail(-3, ex.what());
    }
}
int sk_b200_engine_fetch_output(Engine* e, OutputC* out) {
    if (!e || !e->dev) return fail(-1, "engine handle is null");
    if (!out || !out->radiance) return fail(-1, "output handle is null");
    std::lock_guard<std::mutex> lock(e->mtx);
    try {
        const int nlos = e->ncols();
        if ((long long)out->nrad < (long long)(e->staged_start + e->staged_count) * nlos)
            return fail(-2, "output radiance too small for the staged wavelength range");
        e->dev->fetch(out->radiance + (size_t)e->staged_start * nlos);
        return 0;
    } catch (const std::exception& ex) {
        return fail(-3, ex.what());
    }
}
int sk_b200_engine_get_timings(Engine* e, double* out_ms, int n) {
    if (!e || !e->dev || !out_ms) return -1;
    const double* t = e->dev->timings_ms();
    for (int i = 0; i < n; ++i) out_ms[i] = i < disco::T_NSLOTS ? t[i] : 0.0;
    return 0;
}
long long sk_b200_engine_kernel_launches(Engine* e) { return (e && e->dev) ? e->dev->kernel_launches() : -1; }
int sk_b200_engine_info(Engine* e, int* num_azimuth, int* chunk_wavelengths, double* workspace_mb_per_wavelength) {
    if (!e || !e->dev) return -1;
    if (num_azimuth) *num_azimuth = e->dev->num_azimuth_solved();
    if (chunk_wavelengths) *chunk_wavelengths = e->dev->chunk_wavelengths();
    if (workspace_mb_per_wavelength) *workspace_mb_per_wavelength = e->dev->workspace_bytes_per_wavelength() / 1048576.0;
    return 0;
}
long long sk_b200_engine_debug_copy(Engine* e, const char* name, double* host, long long max_n) {
    if (!e || !e->dev || !name || !host) return -1;
    return (long long)e->dev->debug_copy(name, host, (size_t)max_n);
}
// ---- wavelength-sharded runs: one process per GPU, final gather over NCCL (disco_comm.cpp) ----
static std::unique_ptr<disco::Comm> g_comm;
int sk_b200_comm_unique_id(char* id, int nbytes) {
    if (!id || nbytes < (int)sizeof(disco::NcclUniqueId)) return fail(-2, "sk_b200_comm_unique_id: need a 128-byte buffer");
    try {
        disco::NcclUniqueId u;
        disco::comm_unique_id(&u);
        std::memcpy(id, u.internal, sizeof(u.internal));
        return 0;
    } catch (const std::exception& ex) {
        return fail(-3, ex.what());
    }
}
int sk_b200_comm_init(const char* id, int rank, int world) {
    if (!id) return fail(-1, "sk_b200_comm_init: null id");
    try {
        disco::NcclUniqueId u;
        std::memcpy(u.internal, id, sizeof(u.internal));
        g_comm.reset();
        g_comm = std::make_unique<disco::Comm>(u, rank, world);
        return 0;
    } catch (const std::exception& ex) {
        return fail(-3, ex.what());
    }
}
int sk_b200_comm_destroy() {
    g_comm.reset();
    return 0;
}
int sk_b200_engine_gather_output(Engine* e, OutputC* root_output, int root, const int* block_start, const int* block_count,
                                 int nw_total, double* ms_out) {
    if (!e || !e->dev) return fail(-1, "engine handle is null");
    if (!g_comm) return fail(-1, "sk_b200_comm_init must be called first");
    if (!block_start || !block_count) return fail(-1, "sk_b200_engine_gather_output: null block tables");
    std::lock_guard<std::mutex> lock(e->mtx);
    try {
        const int nlos = e->ncols();
        std::vector<double*> maps, surfs;
        double* rad = nullptr;
        if (g_comm->rank() == root) {
            if (!root_output || !root_output->radiance) return fail(-1, "gather root needs an output handle");
            if ((long long)root_output->nrad != (long long)nw_total * nlos)
                return fail(-2, "gather root output has the wrong size (expected nw_total * nlos)");
            rad = root_output->radiance;
            for (auto& kv : root_output->derivs) {
                if ((long long)kv.second.nrad != (long long)nw_total * nlos) return fail(-2, "gather root derivative memory has the wrong size");
                maps.push_back(kv.second.ptr);
            }
            for (auto& kv : root_output->surface_derivs) surfs.push_back(kv.second.ptr);
        }
        double ms[2] = {0.0, 0.0};
        e->dev->gather_to_root(*g_comm, root, block_start, block_count, nw_total, rad, maps.data(), surfs.data(), ms);
        if (ms_out) {
            ms_out[0] = ms[0];
            ms_out[1] = ms[1];
        }
        return 0;
    } catch (const std::exception& ex) {
        return fail(-3, ex.what());
    }
}

// Host-only check of the spherical geometry plan (no device needed): traces the viewing rays and returns, per ray,
// [line-of-sight optical depth for the extinction profile `ext` [nloc], number of segments, sum of the DO source
// interpolation weights, cosine of the single-scattering angle, solar optical depth at the far end of the ray, solar
// optical depth at the near end] (6 doubles per ray); *num_points = needed source-table points, sza_grid [num_sza].
int sk_b200_limb_plan_check(Geometry1D* geometry, ViewingGeometry* viewing, int nstr, int num_sza, const double* ext,
                            double* per_ray, int* num_points, double* sza_grid) {
    if (!geometry || !viewing || !ext || !per_ray) return fail(-1, "sk_b200_limb_plan_check: null argument");
    try {
        disco::LimbOptions lo;
        lo.num_sza = num_sza;
        lo.ms_do = true;
        lo.ss_exact = true;
        disco::LimbPlan P = disco::build_limb_plan(nstr, geometry->spec, viewing->ordered, lo);
        for (int r = 0; r < P.nrays; ++r) {
            double od = 0.0, wsum = 0.0;
            for (int sg = P.seg_start[r]; sg < P.seg_start[r + 1]; ++sg) {
                for (int k = 0; k < disco::kLimbStencil; ++k) od += P.od_w[(size_t)sg * disco::kLimbStencil + k] * ext[P.od_idx[(size_t)sg * disco::kLimbStencil + k]];
                for (int e = 0; e < disco::kLimbSrcEntries; ++e) wsum += P.src_w[(size_t)sg * disco::kLimbSrcEntries + e];
            }
            auto solar_od = [&](int b) {
                double v = 0.0;
                for (int e = P.sol_start[b]; e < P.sol_start[b + 1]; ++e) v += P.sol_w[e] * ext[P.sol_idx[e]];
                return P.sol_blocked[b] ? -1.0 : v;
            };
            const int b0 = P.seg_start[r] + r, nb = P.seg_start[r + 1] - P.seg_start[r];
            double* o = per_ray + (size_t)r * 6;
            o[0] = od;
            o[1] = nb;
            o[2] = wsum;
            o[3] = P.ray_cos_scatter[r];
            o[4] = solar_od(b0);
            o[5] = solar_od(b0 + nb);
        }
        if (num_points) *num_points = P.npts;
        if (sza_grid)
            for (int i = 0; i < P.nsza; ++i) sza_grid[i] = P.sza_grid[i];
        return 0;
    } catch (const std::exception& ex) {
        return fail(-3, ex.what());
    }
}

int sk_b200_adjoint_reuses_factors(int n_half_streams, int nlos) { return disco::adjoint_reuses_factors(n_half_streams, nlos) ? 1 : 0; }
double sk_b200_measure_fp64_tflops() { return disco::measure_fp64_tflops(); }
void* sk_b200_host_alloc(size_t nbytes) {
    try {
        return host_alloc(nbytes);
    } catch (...) {
        return nullptr;
    }
}
void sk_b200_host_free(void* p) { host_free(p); }
// Page-lock caller-owned host memory in place (numpy / ndarray buffers, shared-memory segments): copies to and from it
// then run at PCIe speed and overlap with kernels.  Returns 0, or -3 when the range cannot be registered.
int sk_b200_host_register(void* p, size_t nbytes) {
    if (!p || nbytes == 0) return -1;
    cudaError_t err = cudaHostRegister(p, nbytes, cudaHostRegisterPortable);
    if (err == cudaErrorHostMemoryAlreadyRegistered) {
        (void)cudaGetLastError();
        return 0;
    }
    if (err != cudaSuccess) {
        (void)cudaGetLastError();
        return fail(-3, std::string("cudaHostRegister: ") + cudaGetErrorString(err));
    }
    return 0;
}
int sk_b200_host_unregister(void* p) {
    if (!p) return -1;
    cudaError_t err = cudaHostUnregister(p);
    (void)cudaGetLastError();
    return err == cudaSuccess ? 0 : -3;
}
int sk_b200_engine_set_workspace_gb(Engine* e, double gb) {
    if (!e || !e->dev) return -1;
    e->dev->set_workspace_gb(gb);
    return 0;
}

}  // extern "C"
