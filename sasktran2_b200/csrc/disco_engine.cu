#include "disco_engine.h"
#include "disco_wf_body.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>

namespace disco {

#define CUDA_OK(call)                                                                                   \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess)                                                                          \
            throw std::runtime_error(std::string("CUDA error: ") + cudaGetErrorString(e_) + " at " +    \
                                     __FILE__ + ":" + std::to_string(__LINE__));                        \
    } while (0)

template <class T>
T* DeviceEngine::dalloc(size_t n) {
    T* p = nullptr;
    CUDA_OK(cudaMalloc((void**)&p, std::max<size_t>(n, 1) * sizeof(T)));
    return p;
}

template <class T>
static T* upload(const std::vector<T>& v) {
    T* p = nullptr;
    CUDA_OK(cudaMalloc((void**)&p, std::max<size_t>(v.size(), 1) * sizeof(T)));
    if (!v.empty()) CUDA_OK(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return p;
}

DeviceEngine::DeviceEngine(const EngineOptions& opt, const HostPlan& plan) : m_opt(opt), m_plan(plan) {
    m_nrad = plan.nlos;
    init(opt);
}

DeviceEngine::DeviceEngine(const EngineOptions& opt, const HostPlan& plan, const LimbPlan& limb)
    : m_opt(opt), m_plan(plan), m_is_limb(true), m_limb(limb) {
    if (plan.nlos != 0) throw std::runtime_error("limb engine: the DO plan must not carry lines of sight");
    m_nrad = limb.nrays;
    init(opt);
    init_limb();
}

template <class T>
static T* upload_tracked(const std::vector<T>& v, std::vector<void*>& owner) {
    T* p = upload(v);
    owner.push_back((void*)p);
    return p;
}

// device copies of the limb geometry tables and of the per-SZA solar tables
void DeviceEngine::init_limb() {
    const LimbPlan& P = m_limb;
    LimbView& Lv = m_lview;
    std::memset(&Lv, 0, sizeof(Lv));
    Lv.nrays = P.nrays;
    Lv.nsza = P.nsza;
    Lv.npts = P.npts;
    Lv.nseg = P.nseg;
    Lv.nss = P.nss;
    Lv.ms_do = P.ms_do ? 1 : 0;
    Lv.ss_exact = P.ss_exact ? 1 : 0;
    auto& own = m_limb_ptrs;
    Lv.layer_fraction = upload_tracked(P.layer_fraction, own);
    Lv.lp_ang = upload_tracked(P.lp_ang, own);
    Lv.pt_angle = upload_tracked(P.pt_angle, own);
    Lv.pt_alt = upload_tracked(P.pt_alt, own);
    Lv.pt_sza = upload_tracked(P.pt_sza, own);
    Lv.seg_start = upload_tracked(P.seg_start, own);
    Lv.od_idx = upload_tracked(P.od_idx, own);
    Lv.od_w = upload_tracked(P.od_w, own);
    Lv.ent_w = upload_tracked(P.ent_w, own);
    Lv.exit_w = upload_tracked(P.exit_w, own);
    Lv.mid_idx = upload_tracked(P.mid_idx, own);
    Lv.mid_w = upload_tracked(P.mid_w, own);
    Lv.seg_len = upload_tracked(P.seg_len, own);
    Lv.seg_qfrac = upload_tracked(P.seg_qfrac, own);
    Lv.seg_lower = upload_tracked(P.seg_lower, own);
    Lv.src_pt = upload_tracked(P.src_pt, own);
    Lv.src_w = upload_tracked(P.src_w, own);
    Lv.src_cos = upload_tracked(P.src_cos, own);
    Lv.gnd_hit = upload_tracked(P.gnd_hit, own);
    Lv.gnd_sza_idx = upload_tracked(P.gnd_sza_idx, own);
    Lv.gnd_sza_w = upload_tracked(P.gnd_sza_w, own);
    Lv.gnd_mu_in = upload_tracked(P.gnd_mu_in, own);
    Lv.wig_ss = upload_tracked(P.wig_ss, own);
    Lv.sol_start = upload_tracked(P.sol_start, own);
    Lv.sol_idx = upload_tracked(P.sol_idx, own);
    Lv.sol_w = upload_tracked(P.sol_w, own);
    Lv.sol_blocked = upload_tracked(P.sol_blocked, own);
    Lv.ray_order = upload_tracked(P.ray_order, own);
    for (const HostPlan& sp : P.sza_plans) {
        d_sza_lp_csz.push_back(upload_tracked(sp.lp_csz, own));
        d_sza_chapman.push_back(upload_tracked(sp.chapman, own));
    }
}

bool DeviceEngine::limb_shared_factorisation() const {
    static const bool off = [] {
        const char* e = std::getenv("SK_B200_LIMB_SHARED");
        return e && e[0] == '0';
    }();
    return m_is_limb && m_fast && !off && bvp_multi_supported(m_plan.N, m_limb.nsza);
}

void DeviceEngine::init(const EngineOptions& opt) {
    const HostPlan& plan = m_plan;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        throw std::runtime_error("sasktran2_b200: no CUDA device available (there is no CPU fallback)");
    if (opt.device >= 0) CUDA_OK(cudaSetDevice(opt.device));
    CUDA_OK(cudaGetDevice(&m_device));
    if (!nstr_supported(plan.nstr))
        throw std::runtime_error("sasktran2_b200: num_streams must be one of 2, 4, 8, 16, 32");
    if (m_opt.workspace_gb <= 0.0) {
        // a chunk should hold thousands of wavelengths (tens of thousands of warps per launch) so that the last
        // wave of every kernel is a small fraction of it; B200 has 180 GB
        size_t free_b = 0, total_b = 0;
        CUDA_OK(cudaMemGetInfo(&free_b, &total_b));
        m_opt.workspace_gb = std::min(32.0, 0.25 * (double)free_b / (1024.0 * 1024.0 * 1024.0));
    }
    CUDA_OK(cudaStreamCreateWithFlags(&m_stream, cudaStreamNonBlocking));
    CUDA_OK(cudaStreamCreateWithFlags(&m_copy, cudaStreamNonBlocking));
    CUDA_OK(cudaEventCreateWithFlags(&m_ev_h2d_tail, cudaEventDisableTiming));
    CUDA_OK(cudaEventCreateWithFlags(&m_ev_out_ready, cudaEventDisableTiming));
    for (auto& ev : m_ev) CUDA_OK(cudaEventCreate(&ev));
    d_mu = upload(plan.mu);
    d_wt = upload(plan.wt);
    d_lp_mu = upload(plan.lp_mu);
    d_lp_csz = upload(plan.lp_csz);
    d_lp_los = upload(plan.lp_los);
    d_los_mu = upload(plan.los_mu);
    if (fast_path_supported(plan.N)) {
        // Tables::wf_tab: what every block of k_wf_layer_fast used to rebuild from lp_mu / lp_los / lp_csz (one FP64
        // division per entry of tM, 12 % of that kernel's stall samples) - the same products, formed once
        const int nstr = plan.nstr, N = plan.N, nlos = plan.nlos;
        const size_t td = 2 * (size_t)nstr * N + (size_t)nlos * nstr + nstr + N;
        std::vector<double> tab(td * nstr, 0.0);
        for (int m = 0; m < nstr; ++m) {
            double* tW = tab.data() + td * m;
            double* tM = tW + (size_t)nstr * N;
            double* tL = tM + (size_t)nstr * N;
            double* lpc = tL + (size_t)nlos * nstr;
            double* wmu = lpc + nstr;
            for (int l = 0; l < nstr; ++l)
                for (int q = 0; q < N; ++q) {
                    const double lp = plan.lp_mu[((size_t)m * N + q) * nstr + l];
                    tW[l * N + q] = plan.wt[q] * lp;
                    tM[l * N + q] = lp / plan.mu[q];
                }
            for (int los = 0; los < nlos; ++los)
                for (int l = 0; l < nstr; ++l) tL[los * nstr + l] = plan.lp_los[((size_t)los * nstr + m) * nstr + l];
            for (int l = 0; l < nstr; ++l) lpc[l] = plan.lp_csz[(size_t)m * nstr + l];
            for (int i = 0; i < N; ++i) wmu[i] = plan.wt[i] * plan.mu[i];
        }
        d_wf_tab = upload(tab);
        // Tables::los_zero: orders that reach a line of sight with an identically zero Legendre table
        if (nlos > 0) {
            std::vector<unsigned char> lz((size_t)nstr * nlos, 0);
            bool any = false;
            for (int m = 0; m < nstr; ++m)
                for (int los = 0; los < nlos; ++los) {
                    bool zero = true;
                    for (int l = 0; l < nstr && zero; ++l)
                        if (plan.lp_los[((size_t)los * nstr + m) * nstr + l] != 0.0) zero = false;
                    lz[(size_t)m * nlos + los] = zero ? 1 : 0;
                    any = any || zero;
                }
            const char* skip = std::getenv("SK_B200_LOS_SKIP");   // =0: solve the zero orders too (timing / tests)
            if (any && !(skip && skip[0] == '0')) {
                CUDA_OK(cudaMalloc(&d_los_zero, lz.size()));
                CUDA_OK(cudaMemcpy(d_los_zero, lz.data(), lz.size(), cudaMemcpyHostToDevice));
            }
        }
    }
    d_los_cosmphi = upload(plan.los_cosmphi);
    d_layer_dh = upload(plan.layer_dh);
    d_interp_w = upload(plan.interp_w);
    d_interp_idx = upload(plan.interp_idx);
    d_chapman = upload(plan.chapman);
    // Azimuth orders to solve.  The reference sums all nstr orders (do_source_planeparallel.cpp:55-62);
    // an order whose LOS Legendre table P_l^m(mu_los) is identically zero for every line of sight (m > 0 at
    // exactly nadir, d^l_{m0}(0) = 0) contributes exactly zero to every output and is skipped.
    const int nstr = plan.nstr;
    const int nazi = (opt.forced_azimuth > 0) ? std::min(opt.forced_azimuth, nstr) : nstr;
    for (int m = 0; m < nazi; ++m) {
        bool any = false;
        for (int j = 0; j < plan.nlos && !any; ++j)
            for (int l = 0; l < nstr && !any; ++l)
                if (plan.lp_los[((size_t)j * nstr + m) * nstr + l] != 0.0) any = true;
        if (any || plan.nlos == 0) m_mlist.push_back(m);
    }
    if (m_mlist.empty()) m_mlist.push_back(0);
    d_mlist = upload(m_mlist);
    // register-resident layer solve for N = 2, 4, 8; SK_B200_GENERIC=1 forces the generic thread-per-problem
    // kernels (differential testing of the two code paths)
    {
        const char* g = std::getenv("SK_B200_GENERIC");
        m_fast = fast_path_supported(plan.N) && !(g && g[0] == '1');
    }
    d_status = dalloc<unsigned int>(1);
    CUDA_OK(cudaMemset(d_status, 0, sizeof(unsigned int)));
}

void DeviceEngine::free_wf_inputs() {
    if (d_dleg) cudaFree(d_dleg);
    d_dleg = nullptr;
    for (auto& m : m_maps)
        for (void* p : {(void*)m.d_ssa, (void*)m.d_ext, (void*)m.scat, (void*)m.interp, (void*)m.out})
            if (p) cudaFree(p);
    for (auto& s : m_surfs)
        for (void* p : {(void*)s.d_brdf, (void*)s.out})
            if (p) cudaFree(p);
    m_maps.clear();
    m_surfs.clear();
    m_wf_on = false;
    m_ngroups = 0;
}

DeviceEngine::~DeviceEngine() {
    cudaSetDevice(m_device);
    for (auto ev : m_marks) cudaEventDestroy(ev);
    free_wf_inputs();
    free_inputs();
    free_workspace();
    if (d_gather) cudaFree(d_gather);
    for (void* p : m_limb_ptrs)
        if (p) cudaFree(p);
    if (d_los_od) cudaFree(d_los_od);
    for (void* p : {(void*)d_brdf_args, (void*)d_zero_albedo, (void*)d_brdf_Rss, (void*)d_brdf_rsun, (void*)d_brdf_Rls, (void*)d_brdf_rlsun,
                    (void*)d_snow_r0, (void*)d_snow_g, (void*)d_snow_cos, (void*)d_snow_w, (void*)d_snow_scale})
        if (p) cudaFree(p);
    if (d_los_zero) cudaFree(d_los_zero);
    for (void* p : {(void*)d_mu, (void*)d_wt, (void*)d_lp_mu, (void*)d_lp_csz, (void*)d_lp_los, (void*)d_los_mu, (void*)d_wf_tab,
                    (void*)d_los_cosmphi, (void*)d_layer_dh, (void*)d_interp_w, (void*)d_interp_idx,
                    (void*)d_chapman, (void*)d_mlist, (void*)d_status})
        if (p) cudaFree(p);
    for (auto& ev : m_ev)
        if (ev) cudaEventDestroy(ev);
    if (m_ev_h2d_tail) cudaEventDestroy(m_ev_h2d_tail);
    if (m_ev_out_ready) cudaEventDestroy(m_ev_out_ready);
    if (m_copy) cudaStreamDestroy(m_copy);
    if (m_stream) cudaStreamDestroy(m_stream);
}

void DeviceEngine::free_inputs() {
    for (void* p : {(void*)d_ext, (void*)d_ssa, (void*)d_leg, (void*)d_solar, (void*)d_albedo, (void*)d_radiance})
        if (p) cudaFree(p);
    d_ext = d_ssa = d_leg = d_solar = d_albedo = d_radiance = nullptr;
    m_cap_nw = m_cap_nleg = 0;
    if (d_emission) cudaFree(d_emission);
    if (d_semis) cudaFree(d_semis);
    d_emission = d_semis = nullptr;
    m_cap_emission = m_cap_semis = 0;
    if (d_fdm) cudaFree(d_fdm);
    d_fdm = nullptr;
    m_cap_fdm = 0;
}

void DeviceEngine::free_workspace() {
    for (void* p : m_ws_ptrs)
        if (p) cudaFree(p);
    m_ws_ptrs.clear();
    m_ws_chunk = 0;
}

size_t DeviceEngine::workspace_bytes_per_wavelength() const { return ws_bytes(m_wf_on, m_ngroups); }

int DeviceEngine::planned_chunk(int nw, bool wf_on, int ngroups) const {
    const double budget = m_opt.workspace_gb * 1024.0 * 1024.0 * 1024.0;
    long long c = (long long)(budget / (double)ws_bytes(wf_on, ngroups));
    c = std::max<long long>(c, 1);
    return (int)std::min<long long>(c, std::max(nw, 1));
}

size_t DeviceEngine::ws_bytes(bool wf_on, int ngroups) const {
    // the dedicated two-stream kernel keeps everything in registers: no per-wavelength workspace at all
    if (m_opt.twostream && !wf_on && twostream_supported(m_plan.L, m_plan.plane_parallel)) return sizeof(double);
    const size_t N = m_plan.N, L = m_plan.L, nstr = m_plan.nstr, nlos = m_plan.nlos, M = m_mlist.size();
    size_t d = L * (6 + nstr) + 2 * (L + 1);                       // layer optics
    d += 2 * M * L * N * N + M * L * 2 * N + M * L * 4 * N;         // W+, W-, k|theta, G
    d += 2 * N + 1;                                                 // surface sums
    const size_t vw = m_fast ? N : 1;
    d += M * nlos * L * 2 * N + M * nlos * L * vw;                  // wvec, vsrc
    if (m_fast) d += 3 * (N * (N + 1) / 2) * M * L + nlos * (L + 1) + nlos * L * 3;  // eigen planes, LOS exponentials
    d += M * L * 2 * N;                                             // x
    if (m_brdf_kind != 0) d += M * (2 * N * N + 2 * N);             // kernel-based surface sums
    if (m_brdf_kind != 0 && wf_on) d += M * (N + nlos) * (N + 1);   // reflection rows for the weighting functions
    if (m_brdf_kind == kBrdfKokhanovsky) d += M * (N * N + N + nlos * N + nlos);   // per-wavelength Fourier coefficients
    if (m_is_limb) {
        const size_t nsza = m_limb.nsza, npts = m_limb.npts, nrays = m_limb.nrays;
        if (m_limb.ms_do) d += nsza * L * M * nstr + nsza + npts * M;    // Legendre projections, ground source, source table
        if (m_limb.ms_do && limb_shared_factorisation())   // per-SZA copies of secant, beam, G, surface sums, x (+ wider factor rows)
            d += (nsza - 1) * (L + (L + 1) + M * L * 4 * N + (2 * N + 1) + M * L * 2 * N) + M * (bvp_fac_stride((int)N, (int)nsza, (int)L) - bvp_fac_stride((int)N, 1, (int)L));
        if (m_limb.ss_exact) d += nrays * m_plan.nloc;                     // single-scatter phase function
    }
    if (!wf_on) {
        d += M * bvp_fac_stride((int)N, 1, (int)L);                  // LU pivot rows (forward solve)
    } else {
        const size_t G = ngroups, nrhs = adjoint_max_rhs((int)nlos), ngrp = adjoint_groups_per_problem((int)nlos);
        if (adjoint_reuses_factors((int)N, (int)nlos))               // forward pivot rows + multipliers + U^T y
            d += M * (bvp_fac_stride((int)N, 1, (int)L) + bvp_lfac_stride((int)N, (int)L)) + M * nlos * 2 * N * L;
        else
            d += M * ngrp * bvp_fac_stride((int)N, (int)nrhs, (int)L);   // LU pivot rows (forward and adjoint)
        d += M * nlos * 2 * N * L;                                  // adjoint solutions
        d += L * G * nstr;                                          // Legendre derivative directions
        d += M * nlos * L * (G + 4) + M * nlos * L + nlos * 3;      // local lanes, sources, ground terms
        d += nlos * (m_plan.nloc * (2 + G) + 1) + nlos * 3 * (L + 1);  // native derivatives, chain scratch
    }
    return d * sizeof(double);
}

int DeviceEngine::chunk_wavelengths() const {
    const double budget = m_opt.workspace_gb * 1024.0 * 1024.0 * 1024.0;
    long long c = (long long)(budget / (double)workspace_bytes_per_wavelength());
    c = std::max<long long>(c, 1);
    return (int)std::min<long long>(c, std::max(m_nw, 1));
}

void DeviceEngine::ensure_workspace(int chunk) {
    // the derivative arrays are sized by the number of scattering groups: an atmosphere with more groups than the
    // workspace was built for needs a new one even when the chunk fits
    if (chunk <= m_ws_chunk && m_ws_wf == m_wf_on && (!m_wf_on || m_ws_ngroups == m_ngroups) && m_ws_brdf == (m_brdf_kind != 0) && (m_brdf_kind == 0 || m_ws_brdf_kind == m_brdf_kind)) return;
    free_workspace();
    const size_t N = m_plan.N, L = m_plan.L, nstr = m_plan.nstr, nlos = m_plan.nlos, M = m_mlist.size();
    const size_t c = chunk;
    m_dbg.clear();
    auto A = [&](const char* name, size_t n) {
        double* p = dalloc<double>(n);
        m_ws_ptrs.push_back(p);
        m_dbg.push_back({name, {p, n}});
        return p;
    };
    ChunkView& V = m_view;
    if (twostream_direct()) {
        std::memset(&V, 0, sizeof(V));
        m_ws_brdf = m_brdf_kind != 0;
        m_ws_wf = m_wf_on;
        m_ws_ngroups = 0;
        m_ws_chunk = chunk;
        return;
    }
    V.lay_od = A("lay_od", c * L);
    V.lay_ssa = A("lay_ssa", c * L);
    V.lay_beta = A("lay_beta", c * L * nstr);
    const size_t zs = (m_is_limb && m_limb.ms_do && limb_shared_factorisation()) ? (size_t)m_limb.nsza : 1;   // per-SZA slices
    V.lay_secant = A("lay_secant", zs * c * L);
    V.lay_trans = A("lay_trans", zs * c * (L + 1));
    V.lay_cumod = A("lay_cumod", c * (L + 1));
    V.lay_totext = A("lay_totext", c * L);
    V.lay_scatext = A("lay_scatext", c * L);
    V.lay_thermal = A("lay_thermal", c * L * 2);
    V.Wp = A("Wp", c * M * L * N * N);
    V.Wm = A("Wm", c * M * L * N * N);
    V.kth = A("kth", c * M * L * 2 * N);
    V.G = A("G", zs * c * M * L * 4 * N);
    V.surf = A("surf", zs * c * (2 * N + 1));
    V.wvec = A("wvec", c * M * nlos * L * 2 * N);
    V.vsrc_w = m_fast ? (int)N : 1;
    V.vsrc = A("vsrc", c * M * nlos * L * V.vsrc_w);
    V.eigS = V.eigH = V.eigC = V.los_att = V.los_lay = nullptr;
    if (m_fast) {
        const size_t ts = N * (N + 1) / 2;
        V.eigS = A("eigS", ts * c * M * L);
        V.eigH = A("eigH", ts * c * M * L);
        V.eigC = A("eigC", ts * c * M * L);
        V.los_att = A("los_att", c * nlos * (L + 1));
        V.los_lay = A("los_lay", c * nlos * L * 3);
    }
    V.xsol = A("xsol", zs * c * M * L * 2 * N);
    V.gsurf = V.gsurf_out = nullptr;
    V.gsurf_stride = (int)(2 * N * N + 2 * N);
    if (m_brdf_kind != 0) V.gsurf = V.gsurf_out = A("gsurf", c * M * V.gsurf_stride);
    V.gsurf_rows = (m_brdf_kind != 0 && m_wf_on) ? A("gsurf_rows", c * M * (N + nlos) * (N + 1)) : nullptr;
    V.wf_gndk = (m_brdf_kind != 0 && m_wf_on) ? A("wf_gndk", c * nlos * 4) : nullptr;   // at most 3 kernel weights
    V.wf_gnd_part = (m_brdf_kind != 0 && m_wf_on) ? A("wf_gnd_part", c * M * nlos * 5) : nullptr;
    m_ws_brdf = m_brdf_kind != 0;
    m_ws_brdf_kind = m_brdf_kind;
    d_brdf_pw = nullptr;
    if (m_brdf_kind == kBrdfKokhanovsky) d_brdf_pw = A("brdf_pw", c * M * (N * N + N + nlos * N + nlos));
    if (m_is_limb) {
        const size_t nsza = m_limb.nsza, npts = m_limb.npts, nrays = m_limb.nrays;
        m_lview.coef = m_lview.ground = m_lview.table = m_lview.phase = nullptr;
        if (m_limb.ms_do) {
            m_lview.coef = A("limb_coef", c * nsza * L * M * nstr);
            m_lview.ground = A("limb_ground", c * nsza);
            m_lview.table = A("limb_table", c * npts * M);
        }
        if (m_limb.ss_exact) m_lview.phase = A("limb_phase", c * nrays * m_plan.nloc);
    }
    if (!m_wf_on) {
        V.fac_stride = bvp_fac_stride((int)N, (int)zs, (int)L);
        V.fac = A("fac", c * M * V.fac_stride);
        V.zadj = V.lfac = V.yadj = nullptr;
        V.lfac_stride = 0;
        V.lay_dbeta = V.wf_loc = V.wf_src = V.wf_gnd = V.wf_native = V.wf_scratch = nullptr;
    } else {
        const size_t G = m_ngroups, nrhs = adjoint_max_rhs((int)nlos), ngrp = adjoint_groups_per_problem((int)nlos);
        if (adjoint_reuses_factors((int)N, (int)nlos)) {
            V.fac_stride = bvp_fac_stride((int)N, 1, (int)L);
            V.fac = A("fac", c * M * V.fac_stride);
            V.lfac_stride = bvp_lfac_stride((int)N, (int)L);
            V.lfac = A("lfac", c * M * V.lfac_stride);
            V.yadj = A("yadj", c * M * nlos * 2 * N * L);
        } else {
            V.fac_stride = bvp_fac_stride((int)N, (int)nrhs, (int)L);
            V.fac = A("fac", c * M * ngrp * V.fac_stride);
            V.lfac = V.yadj = nullptr;
            V.lfac_stride = 0;
        }
        V.zadj = A("zadj", c * M * nlos * 2 * N * L);
        V.lay_dbeta = A("lay_dbeta", c * L * G * nstr);
        V.wf_loc = A("wf_loc", c * M * nlos * L * (G + 4));
        V.wf_src = A("wf_src", c * M * nlos * L);
        V.wf_gnd = A("wf_gnd", c * nlos * 3);
        V.wf_native = A("wf_native", c * nlos * (m_plan.nloc * (2 + G) + 1));
        V.wf_scratch = A("wf_scratch", c * nlos * 3 * (L + 1));
    }
    m_ws_wf = m_wf_on;
    m_ws_ngroups = m_wf_on ? m_ngroups : 0;
    m_ws_chunk = chunk;
}

void DeviceEngine::stage(const AtmosphereArrays& atm, int w0, int nw, const WfRequest* wf) {
    CUDA_OK(cudaSetDevice(m_device));
    if (atm.nloc != m_plan.nloc) throw std::runtime_error("atmosphere and geometry grids differ in size");
    if (w0 < 0 || nw < 0 || w0 + nw > atm.nwavel) throw std::runtime_error("wavelength range out of bounds");
    const size_t nloc = atm.nloc;
    if (nw > m_cap_nw || atm.nleg != m_cap_nleg) {
        free_inputs();
        d_ext = dalloc<double>(nloc * nw);
        d_ssa = dalloc<double>(nloc * nw);
        d_leg = dalloc<double>((size_t)atm.nleg * nloc * nw);
        d_solar = dalloc<double>(nw);
        d_albedo = dalloc<double>(nw);
        d_radiance = dalloc<double>((size_t)nw * std::max(m_nrad, 1));
        if (m_is_limb) {
            if (d_los_od) cudaFree(d_los_od);
            d_los_od = dalloc<double>((size_t)nw * std::max(m_nrad, 1));
        }
        m_cap_nw = nw;
        m_cap_nleg = atm.nleg;
    }
    m_nw = nw;
    m_nleg = atm.nleg;
    // Inside calculate() (m_overlap) only the wavelengths of the first chunk travel on the compute stream; the rest
    // goes to the copy stream and is awaited by the first chunk that needs it (solve_staged).  Every staged array
    // is wavelength-slowest, so head and tail are contiguous ranges of each.
    {
        const bool wf0 = wf && wf->enabled() && nw > 0;
        m_h2d_head = (m_overlap && nw > 0) ? planned_chunk(nw, wf0, wf0 ? (int)wf->d_legendre.size() : 0) : nw;
    }
    const size_t head = (size_t)m_h2d_head, tail = (size_t)nw - head;
    // the copy engine serves same-direction copies in submission order, whatever their stream: all heads are
    // submitted first, the tails are queued here and submitted at the end of stage()
    struct TailCopy { double* dst; const double* src; size_t bytes; };
    std::vector<TailCopy> tails;
    auto h2d = [&](double* dst, const double* src, size_t per_w) {  // src, dst: first staged wavelength
        if (head > 0)
            CUDA_OK(cudaMemcpyAsync(dst, src, sizeof(double) * per_w * head, cudaMemcpyHostToDevice, m_stream));
        if (tail > 0) tails.push_back({dst + per_w * head, src + per_w * head, sizeof(double) * per_w * tail});
    };
    CUDA_OK(cudaEventRecord(m_ev[0], m_stream));
    if (nw > 0) {
        h2d(d_ext, atm.ext + nloc * w0, nloc);
        h2d(d_ssa, atm.ssa + nloc * w0, nloc);
        h2d(d_leg, atm.leg + (size_t)atm.nleg * nloc * w0, (size_t)atm.nleg * nloc);
        h2d(d_solar, atm.solar + w0, 1);
        if (atm.albedo && atm.brdf_kind == 0) h2d(d_albedo, atm.albedo + w0, 1);
    }
    // multiple_scatter_source = TwoStream outside the dedicated kernel (weighting functions, or a grid too tall for its
    // shared-memory tables) runs the 2-stream discrete-ordinates kernels.  The two-stream source averages a layer as the
    // mean of its two grid points whatever the interpolation method (cpp_twostream_source.cpp:1923-2122); the DO layers
    // follow the geometry's interpolation - the same numbers for linear interpolation only (the identity upstream tests).
    if (m_opt.twostream && !twostream_direct_for(wf && wf->enabled()) && m_plan.interp != 1 && nw > 0)
        throw std::runtime_error("B200 two-stream source: weighting functions (and grids above ~150 layers in pseudo-spherical "
                                 "geometry) are solved by the 2-stream discrete-ordinates kernels, which equal the two-stream "
                                 "source for linear interpolation only");
    // thermal sources (radiances only; the reference's emission derivative lanes are not implemented)
    m_emission_on = atm.emission != nullptr && nw > 0;
    m_semis_on = atm.surface_emission != nullptr && nw > 0;
    if (m_emission_on || m_semis_on) {
        if (m_is_limb) throw std::runtime_error("B200 limb path: emission sources are not supported");
        if (twostream_direct()) throw std::runtime_error("B200 two-stream kernel: emission sources are not supported");
        if (wf && wf->enabled()) throw std::runtime_error("B200 DO path: weighting functions with thermal emission are not supported");
    }
    if (m_emission_on) {
        if (nw > m_cap_emission) {
            if (d_emission) cudaFree(d_emission);
            d_emission = dalloc<double>(nloc * nw);
            m_cap_emission = nw;
        }
        h2d(d_emission, atm.emission + nloc * w0, nloc);
    }
    if (m_semis_on) {
        if (nw > m_cap_semis) {
            if (d_semis) cudaFree(d_semis);
            d_semis = dalloc<double>(nw);
            m_cap_semis = nw;
        }
        h2d(d_semis, atm.surface_emission + w0, 1);
    }
    // kernel-based surface: Fourier coefficients of the kernels (once per engine and model), arguments of the range
    m_brdf_kind = atm.brdf_kind;
    if (m_brdf_kind != 0 && nw > 0) {
        if (m_is_limb) throw std::runtime_error("B200 limb path supports the Lambertian BRDF only");
        if (wf && wf->enabled()) {
            // atmospheric weighting functions above a kernel-based BRDF are solved; those w.r.t. its arguments are not
            // ... except for the weights of a linear kernel model (MODIS), whose tables are the derivatives
            if (!wf->surfaces.empty() && m_brdf_kind != kBrdfModis)
                throw std::runtime_error("B200 DO path: weighting functions w.r.t. the argument of the snow BRDF are not supported");
        }
        if (twostream_direct()) throw std::runtime_error("B200 two-stream kernel supports the Lambertian BRDF only");
        if (m_plan.N > 16) throw std::runtime_error("B200 DO path: kernel-based BRDFs need num_streams <= 32");
        if (m_brdf_tab_kind != m_brdf_kind) {
            for (double** p : {&d_brdf_Rss, &d_brdf_rsun, &d_brdf_Rls, &d_brdf_rlsun, &d_snow_r0, &d_snow_g, &d_snow_cos, &d_snow_w, &d_snow_scale})
                if (*p) {
                    cudaFree(*p);
                    *p = nullptr;
                }
            if (m_brdf_kind == kBrdfKokhanovsky) {
                const SnowTables T = build_snow_tables(m_plan);
                d_snow_r0 = upload(T.r0);
                d_snow_g = upload(T.g);
                d_snow_cos = upload(T.cosphi);
                d_snow_w = upload(T.weight);
                d_snow_scale = upload(T.scale);
                m_snow_npairs = T.npairs;
                m_snow_nsamples = T.nsamples;
                m_brdf_nk = 0;
            } else {
                const BrdfTables T = build_brdf_tables(m_brdf_kind, m_plan);
                d_brdf_Rss = upload(T.Rss);
                d_brdf_rsun = upload(T.rsun);
                d_brdf_Rls = upload(T.Rls);
                d_brdf_rlsun = upload(T.rlsun);
                m_brdf_nk = T.nk;
            }
            m_brdf_tab_kind = m_brdf_kind;
        }
        m_brdf_nargs = atm.brdf_nargs;
        if (m_brdf_nargs != brdf_num_args(m_brdf_kind) || !atm.brdf_args) throw std::runtime_error("BRDF arguments do not match the BRDF model");
        const size_t need = (size_t)m_brdf_nargs * nw;
        if (need > m_cap_brdf) {
            if (d_brdf_args) cudaFree(d_brdf_args);
            if (d_zero_albedo) cudaFree(d_zero_albedo);
            d_brdf_args = dalloc<double>(need);
            d_zero_albedo = dalloc<double>(need);
            CUDA_OK(cudaMemset(d_zero_albedo, 0, sizeof(double) * need));
            m_cap_brdf = need;
        }
        h2d(d_brdf_args, atm.brdf_args + (size_t)m_brdf_nargs * w0, (size_t)m_brdf_nargs);
    }
    // delta-M truncation fraction and its derivatives (set by sk_atmosphere_apply_delta_m_scaling)
    m_has_f = atm.f != nullptr && nw > 0;
    if (m_has_f) {
        const size_t G = (wf && wf->enabled()) ? wf->d_f.size() : 0;
        const size_t n2 = nloc * (size_t)nw, need = (1 + G) * n2;
        if (need > m_cap_fdm) {
            if (d_fdm) cudaFree(d_fdm);
            d_fdm = dalloc<double>(need);
            m_cap_fdm = need;
        }
        h2d(d_fdm, atm.f + nloc * w0, nloc);
        for (size_t g = 0; g < G; ++g) {
            if (wf->d_f[g])
                h2d(d_fdm + (1 + g) * n2, wf->d_f[g] + nloc * w0, nloc);
            else
                CUDA_OK(cudaMemsetAsync(d_fdm + (1 + g) * n2, 0, sizeof(double) * n2, m_stream));
        }
    }
    // weighting-function inputs of the staged range.  Device buffers are kept from the previous call when the
    // request has the same shape (same number of wavelengths, groups, mappings and output sizes).
    m_w0 = w0;
    m_nw_total = atm.nwavel;
    const bool want_wf = wf && wf->enabled() && nw > 0;
    if (want_wf && m_is_limb) throw std::runtime_error("B200 limb path: weighting functions are not supported in spherical geometry");
    bool same_shape = want_wf && m_wf_on && m_wf_nw == nw && m_wf_nleg == atm.nleg &&
                      m_ngroups == (int)wf->d_legendre.size() && m_maps.size() == wf->mappings.size() &&
                      m_surfs.size() == wf->surfaces.size();
    if (same_shape)
        for (size_t i = 0; i < m_maps.size(); ++i) {
            const auto& a = m_maps[i].host;
            const auto& b = wf->mappings[i];
            if (a.nout != b.nout || (a.scat_factor == nullptr) != (b.scat_factor == nullptr) ||
                (a.interpolator == nullptr) != (b.interpolator == nullptr))
                same_shape = false;
        }
    if (!same_shape) free_wf_inputs();
    if (want_wf) {
        if (wf->d_legendre.size() > 2) throw std::runtime_error("B200 DO path supports at most 2 scattering derivative groups");
        const size_t nl3 = (size_t)atm.nleg * nloc * nw;
        const size_t n2 = nloc * (size_t)nw;
        if (!same_shape) {
            m_wf_on = true;
            m_wf_nw = nw;
            m_wf_nleg = atm.nleg;
            m_ngroups = (int)wf->d_legendre.size();
            if (m_ngroups > 0) d_dleg = dalloc<double>(nl3 * m_ngroups);
            for (const auto& mp : wf->mappings) {
                DevMapping dm;
                dm.d_ssa = dalloc<double>(n2);
                dm.d_ext = dalloc<double>(n2);
                if (mp.scat_factor) dm.scat = dalloc<double>(n2);
                if (mp.interpolator) dm.interp = dalloc<double>(nloc * (size_t)mp.nout);
                dm.out = dalloc<double>((size_t)mp.nout * nw * m_plan.nlos);
                m_maps.push_back(dm);
            }
            for (size_t i = 0; i < wf->surfaces.size(); ++i) {
                DevSurface ds;
                ds.d_brdf = dalloc<double>((size_t)nw * std::max(wf->surfaces[i].nargs, 1));
                ds.out = dalloc<double>((size_t)nw * m_plan.nlos);
                m_surfs.push_back(ds);
            }
        }
        for (int g = 0; g < m_ngroups; ++g)
            h2d(d_dleg + nl3 * g, wf->d_legendre[g] + (size_t)atm.nleg * nloc * w0, (size_t)atm.nleg * nloc);
        for (size_t i = 0; i < m_maps.size(); ++i) {
            DevMapping& dm = m_maps[i];
            const WfMapping& mp = wf->mappings[i];
            dm.host = mp;
            h2d(dm.d_ssa, mp.d_ssa + nloc * w0, nloc);
            h2d(dm.d_ext, mp.d_extinction + nloc * w0, nloc);
            if (mp.scat_factor) h2d(dm.scat, mp.scat_factor + nloc * w0, nloc);
            if (mp.interpolator)
                CUDA_OK(cudaMemcpyAsync(dm.interp, mp.interpolator, sizeof(double) * nloc * mp.nout, cudaMemcpyHostToDevice, m_stream));
        }
        for (size_t i = 0; i < m_surfs.size(); ++i) {
            m_surfs[i].host = wf->surfaces[i];
            for (int k = 0; k < std::max(wf->surfaces[i].nargs, 1); ++k)   // one column per BRDF argument
                h2d(m_surfs[i].d_brdf + (size_t)k * nw, wf->surfaces[i].d_brdf + (size_t)k * wf->surfaces[i].nwavel + w0, 1);
        }
    }
    CUDA_OK(cudaEventRecord(m_ev[1], m_stream));
    for (const auto& t : tails) CUDA_OK(cudaMemcpyAsync(t.dst, t.src, t.bytes, cudaMemcpyHostToDevice, m_copy));
    m_tail_pending = !tails.empty();
    if (m_tail_pending) CUDA_OK(cudaEventRecord(m_ev_h2d_tail, m_copy));
    if (m_overlap) return;  // calculate(): the kernels queue right behind the head copies; fetch() reads the timing
    CUDA_OK(cudaStreamSynchronize(m_stream));
    float ms = 0;
    CUDA_OK(cudaEventElapsedTime(&ms, m_ev[0], m_ev[1]));
    m_ms[T_H2D] = ms;
}

void DeviceEngine::solve_staged() {
    for (int i = T_OPTICS; i < T_NSLOTS; ++i)
        if (i != T_D2H) m_ms[i] = 0.0;
    m_launches = 0;
    CUDA_OK(cudaSetDevice(m_device));
    if (m_nw == 0 || m_nrad == 0) return;
    const int chunk = chunk_wavelengths();
    ensure_workspace(chunk);
    CUDA_OK(cudaMemsetAsync(d_status, 0, sizeof(unsigned int), m_stream));
    ChunkView V = m_view;
    V.T.nstr = m_plan.nstr;
    V.T.N = m_plan.N;
    V.T.L = m_plan.L;
    V.T.nloc = m_plan.nloc;
    V.T.nlos = m_plan.nlos;
    V.T.csz = m_plan.csz;
    V.T.mu = d_mu;
    V.T.wt = d_wt;
    V.T.lp_mu = d_lp_mu;
    V.T.lp_csz = d_lp_csz;
    V.T.lp_los = d_lp_los;
    V.T.los_mu = d_los_mu;
    V.T.wf_tab = d_wf_tab;
    V.T.los_zero = d_los_zero;
    V.T.los_cosmphi = d_los_cosmphi;
    V.layer_dh = d_layer_dh;
    V.interp_idx = d_interp_idx;
    V.interp_w = d_interp_w;
    V.chapman = d_chapman;
    V.plane_parallel = m_plan.plane_parallel ? 1 : 0;
    V.nleg = m_nleg;
    V.include_ss = m_opt.include_ss ? 1 : 0;
    V.M = (int)m_mlist.size();
    V.m_list = d_mlist;
    V.status = d_status;
    V.ngroups = m_wf_on ? m_ngroups : 0;
    const size_t nloc = m_plan.nloc;
    struct Span { cudaEvent_t a, b; int slot; };
    // per-kernel timing: events are recorded around every launch; elapsed times are read after the final sync
    // (the events live in m_marks and are reused by every solve: nothing to leak when a launch throws)
    std::vector<cudaEvent_t> evs;
    std::vector<int> slots;
    auto mark = [&]() {
        if (evs.size() == m_marks.size()) {
            cudaEvent_t ev;
            CUDA_OK(cudaEventCreate(&ev));
            m_marks.push_back(ev);
        }
        cudaEvent_t ev = m_marks[evs.size()];
        CUDA_OK(cudaEventRecord(ev, m_stream));
        evs.push_back(ev);
    };
    mark();
    for (int w0 = 0; w0 < m_nw; w0 += chunk) {
        V.nw = std::min(chunk, m_nw - w0);
        if (m_tail_pending && w0 + V.nw > m_h2d_head) {  // first chunk that reads inputs copied on m_copy
            CUDA_OK(cudaStreamWaitEvent(m_stream, m_ev_h2d_tail, 0));
            m_tail_pending = false;
        }
        if (m_overlap && w0 > 0 && w0 + V.nw >= m_nw) {
            // last chunk: everything before it is final - its way to the host overlaps this chunk's kernels
            CUDA_OK(cudaEventRecord(m_ev_out_ready, m_stream));
            CUDA_OK(cudaStreamWaitEvent(m_copy, m_ev_out_ready, 0));
            copy_outputs(0, w0, m_early_radiance, m_copy);
            m_early_done = w0;
        }
        V.ext = d_ext + nloc * w0;
        V.ssa = d_ssa + nloc * w0;
        V.leg = d_leg + (size_t)m_nleg * nloc * w0;
        V.albedo = (m_brdf_kind != 0) ? d_zero_albedo + w0 : d_albedo + w0;
        V.solar = d_solar + w0;
        V.emission = m_emission_on ? d_emission + nloc * w0 : nullptr;
        V.semis = m_semis_on ? d_semis + w0 : nullptr;
        V.radiance = d_radiance + (size_t)w0 * m_nrad;
        V.dleg = d_dleg ? d_dleg + (size_t)m_nleg * nloc * w0 : nullptr;
        V.dleg_gstride = (size_t)m_nleg * nloc * m_nw;
        V.fdm = m_has_f ? d_fdm + nloc * w0 : nullptr;
        V.fdm_gstride = nloc * (size_t)m_nw;
        if (m_opt.validate_inputs) {
            launch_validate_inputs(V, m_stream);
            m_launches += 1;
        }
        if (twostream_direct()) {
            launch_twostream(V, m_stream);
            mark(); slots.push_back(T_LAYER);
            m_launches += 1;
            continue;
        }
        if (m_is_limb) {
            // Spherical line-of-sight path: the layer optics once, then per SZA of the DO grid the beam, the layer
            // solutions and the BVP with that SZA's chapman factors and solar Legendre table, each followed by the
            // projection of the diffuse field; finally the source table and the line-of-sight integration.
            LimbView Lv = m_lview;
            Lv.radiance = V.radiance;
            Lv.los_od = d_los_od ? d_los_od + (size_t)w0 * m_nrad : nullptr;
            if (m_limb.ms_do) {
                launch_layer_optics(V, m_stream);
                mark(); slots.push_back(T_OPTICS);
                m_launches += 1;
                if (limb_shared_factorisation()) {
                    // The homogeneous solutions and the BVP matrix do not depend on the solar zenith angle: one eigen-solve
                    // and one factorisation serve every SZA of the source table; only the beam, the particular solutions
                    // and the right-hand sides are per SZA (the reference redoes everything per SZA, do_source.cpp:35-58).
                    const size_t cw = (size_t)m_ws_chunk;
                    const size_t sSec = cw * V.T.L, sTr = cw * (V.T.L + 1), sG = cw * V.M * V.T.L * 4 * V.T.N,
                                 sSurf = cw * (2 * V.T.N + 1), sX = cw * V.M * V.T.L * 2 * V.T.N;
                    ChunkView B0 = V;   // slice 0 = the base pointers
                    auto slice = [&](int sz) {
                        ChunkView S = B0;
                        S.T.csz = m_limb.sza_grid[sz];
                        S.T.lp_csz = d_sza_lp_csz[sz];
                        S.chapman = d_sza_chapman[sz];
                        S.lay_secant = B0.lay_secant + sz * sSec;
                        S.lay_trans = B0.lay_trans + sz * sTr;
                        S.G = B0.G + sz * sG;
                        S.surf = B0.surf + sz * sSurf;
                        S.xsol = B0.xsol + sz * sX;
                        return S;
                    };
                    for (int sz = 0; sz < m_limb.nsza; ++sz) launch_beam(slice(sz), m_stream, sz == 0);
                    mark(); slots.push_back(T_OPTICS);
                    launch_layer_eig_fast(B0, m_stream);
                    for (int sz = 0; sz < m_limb.nsza; ++sz) launch_layer_post_fast(slice(sz), m_stream);
                    mark(); slots.push_back(T_LAYER);
                    ChunkView Mv = B0;
                    Mv.nsza = m_limb.nsza;
                    Mv.T.csz = m_limb.sza_grid[0];
                    Mv.T.lp_csz = d_sza_lp_csz[0];
                    Mv.sza_G = sG;
                    Mv.sza_surf = sSurf;
                    Mv.sza_trans = sTr;
                    Mv.sza_xsol = sX;
                    for (int sz = 0; sz < m_limb.nsza; ++sz) Mv.sza_csz[sz] = m_limb.sza_grid[sz];
                    launch_bvp_multi(Mv, m_stream);
                    mark(); slots.push_back(T_BVP);
                    for (int sz = 0; sz < m_limb.nsza; ++sz) launch_limb_coef(slice(sz), Lv, sz, m_stream);
                    mark(); slots.push_back(T_LIMB_SOURCE);
                    m_launches += 2 * m_limb.nsza + 2 + m_limb.nsza + 1;
                }
                for (int sz = 0; sz < (limb_shared_factorisation() ? 0 : m_limb.nsza); ++sz) {
                    V.T.csz = m_limb.sza_grid[sz];
                    V.T.lp_csz = d_sza_lp_csz[sz];
                    V.chapman = d_sza_chapman[sz];
                    launch_beam(V, m_stream, sz == 0);
                    mark(); slots.push_back(T_OPTICS);
                    if (m_fast) {
                        launch_layer_solve_fast(V, m_stream);
                        m_launches += 2;
                    } else {
                        launch_layer_solve(V, m_stream);
                    }
                    mark(); slots.push_back(T_LAYER);
                    launch_bvp(V, m_stream);
                    mark(); slots.push_back(T_BVP);
                    launch_limb_coef(V, Lv, sz, m_stream);
                    mark(); slots.push_back(T_LIMB_SOURCE);
                    m_launches += 4;
                }
                launch_limb_table(V, Lv, m_stream);
                mark(); slots.push_back(T_LIMB_SOURCE);
                m_launches += 1;
            }
            if (m_limb.ss_exact) {
                launch_limb_phase(V, Lv, m_stream);
                m_launches += 1;
            }
            launch_limb_integrate(V, Lv, m_stream);
            mark(); slots.push_back(T_LIMB_INTEGRATE);
            m_launches += 1;
            continue;
        }
        launch_layer_optics(V, m_stream);
        launch_beam(V, m_stream);
        mark(); slots.push_back(T_OPTICS);
        if (m_fast) {
            launch_layer_solve_fast(V, m_stream);
            m_launches += 3;
        } else {
            launch_layer_solve(V, m_stream);
        }
        if (m_brdf_kind != 0) {
            BrdfView B;
            B.nk = m_brdf_nk;
            B.nargs = m_brdf_nargs;
            B.Rss = d_brdf_Rss;
            B.rsun = d_brdf_rsun;
            B.Rls = d_brdf_Rls;
            B.rlsun = d_brdf_rlsun;
            B.args = d_brdf_args + (size_t)m_brdf_nargs * w0;
            B.pw = nullptr;
            B.pw_out = nullptr;
            B.npairs = m_snow_npairs;
            B.nsamples = m_snow_nsamples;
            B.snow_r0 = d_snow_r0;
            B.snow_g = d_snow_g;
            B.snow_cos = d_snow_cos;
            B.snow_w = d_snow_w;
            B.snow_scale = d_snow_scale;
            if (m_brdf_kind == kBrdfKokhanovsky) {
                B.pw_out = d_brdf_pw;
                launch_brdf_expand_snow(V, B, m_stream);
                B.pw = d_brdf_pw;
                m_launches += 1;
            }
            launch_surface_general(V, B, m_stream);
            m_launches += 1;
        }
        mark(); slots.push_back(T_LAYER);
        launch_bvp(V, m_stream);
        mark(); slots.push_back(T_BVP);
        launch_radiance(V, m_stream);
        mark(); slots.push_back(T_RADIANCE);
        m_launches += 5;
        if (m_wf_on) {
            launch_bvp_adjoint(V, m_stream);
            mark(); slots.push_back(T_WF_ADJOINT);
            const bool general_brdf = m_brdf_kind != 0;
            if (general_brdf) {   // per-order ground pieces (k_wf_layer), summed in slot order afterwards
                V.brdf_nk = (m_brdf_kind == kBrdfModis) ? m_brdf_nk : 0;
                V.brdf_Rss = d_brdf_Rss;
                V.brdf_rsun = d_brdf_rsun;
                V.brdf_Rls = d_brdf_Rls;
                V.brdf_rlsun = d_brdf_rlsun;
                V.wf_gnd_part = m_view.wf_gnd_part;
                V.wf_gndk = m_surfs.empty() ? nullptr : m_view.wf_gndk;   // weights of the MODIS kernels
            }
            if (m_fast && wf_layer_fast_tile(m_plan.N, m_ngroups, m_plan.nlos) > 0) {
                ChunkView Vw = V;
                Vw.wf_bottom_only = general_brdf ? 1 : 0;
                launch_wf_layer_fast(Vw, m_stream);
                if (general_brdf) {   // the layer on the ground: generic kernel with the reflection rows
                    launch_wf_layer(Vw, m_stream);
                    m_launches += 1;
                }
            } else {
                launch_wf_layer(V, m_stream);
            }
            if (general_brdf) {
                launch_wf_ground_reduce(V, m_stream);
                m_launches += 1;
            }
            mark(); slots.push_back(T_WF_LAYER);
            launch_wf_chain(V, m_stream);
            mark(); slots.push_back(T_WF_CHAIN);
            m_launches += 3;
            for (auto& dm : m_maps) {
                MappingView mv;
                mv.d_ssa = dm.d_ssa;
                mv.d_ext = dm.d_ext;
                mv.scat_factor = dm.scat;
                mv.scat_index = dm.host.scat_index;
                mv.interp = dm.interp;
                mv.nout = dm.host.nout;
                mv.out = dm.out;
                launch_wf_map(V, mv, w0, m_nw, dm.host.log_radiance_space, m_stream);
                m_launches += dm.host.log_radiance_space ? 2 : 1;
            }
            for (auto& ds : m_surfs) {
                if (general_brdf)
                    launch_wf_surface_args(V, ds.d_brdf, (size_t)m_nw, ds.out, w0, m_stream);
                else
                    launch_wf_surface(V, ds.d_brdf, ds.out, w0, m_stream);
                m_launches += 1;
            }
            mark(); slots.push_back(T_WF_MAP);
        }
    }
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaStreamSynchronize(m_stream));
    for (size_t i = 0; i + 1 < evs.size(); ++i) {
        float ms = 0;
        CUDA_OK(cudaEventElapsedTime(&ms, evs[i], evs[i + 1]));
        m_ms[slots[i]] += ms;
    }
    m_ms[T_WF] = m_ms[T_WF_ADJOINT] + m_ms[T_WF_LAYER] + m_ms[T_WF_CHAIN] + m_ms[T_WF_MAP];
    float tot = 0;
    CUDA_OK(cudaEventElapsedTime(&tot, evs.front(), evs.back()));
    m_ms[T_TOTAL_KERNELS] = tot;
    unsigned int st = 0;
    CUDA_OK(cudaMemcpy(&st, d_status, sizeof(st), cudaMemcpyDeviceToHost));
    // input validation (cpp/include/sasktran2/validation/validation.h:12-64) comes first: bad inputs also trip the solver bits
    if (st & 8u) throw std::runtime_error("Invalid input: Atmosphere total extinction contains non-finite values");
    if (st & 16u) throw std::runtime_error("Invalid input: Atmosphere total extinction contains values less than 0");
    if (st & 32u) throw std::runtime_error("Invalid input: Atmosphere single scatter albedo contains non-finite values");
    if (st & 64u) throw std::runtime_error("Invalid input: Atmosphere single scatter albedo contains values less than 0");
    if (st & 128u) throw std::runtime_error("Invalid input: Atmosphere single scatter albedo contains values greater than 1");
    if (st & 1u) throw std::runtime_error("DO homogeneous solution: S- is not positive definite (invalid phase moments?)");
    if (st & 2u)
        throw std::runtime_error(
            "An homogeneous solution was found to be imaginary. An insufficient number of streams is likely.");
    if (st & 4u) throw std::runtime_error("BVP could not be solved since the coefficient matrix was singular.");
}

// device -> host copies of the results of staged wavelengths [w_begin, w_end) (radiance [nw][nlos]; weighting
// functions [nout][nw_total][nlos], one strided copy per mapping)
void DeviceEngine::copy_outputs(int w_begin, int w_end, double* radiance_host, cudaStream_t s) {
    if (w_end <= w_begin || m_nrad <= 0) return;
    const size_t nlos = m_nrad, n = (size_t)(w_end - w_begin), off = (size_t)w_begin * nlos;
    CUDA_OK(cudaMemcpyAsync(radiance_host + off, d_radiance + off, sizeof(double) * n * nlos, cudaMemcpyDeviceToHost, s));
    if (m_wf_on) {
        const size_t width = sizeof(double) * n * nlos;
        for (auto& dm : m_maps)
            CUDA_OK(cudaMemcpy2DAsync(dm.host.out + (size_t)m_w0 * nlos + off, sizeof(double) * (size_t)m_nw_total * nlos,
                                      dm.out + off, sizeof(double) * (size_t)m_nw * nlos, width, dm.host.nout,
                                      cudaMemcpyDeviceToHost, s));
        for (auto& ds : m_surfs)
            CUDA_OK(cudaMemcpyAsync(ds.host.out + (size_t)m_w0 * nlos + off, ds.out + off, width, cudaMemcpyDeviceToHost, s));
    }
}

void DeviceEngine::fetch(double* radiance_host) {
    CUDA_OK(cudaSetDevice(m_device));
    CUDA_OK(cudaEventRecord(m_ev[2], m_stream));
    copy_outputs(m_early_done, m_nw, radiance_host, m_stream);   // everything, unless calculate() sent a part ahead
    CUDA_OK(cudaEventRecord(m_ev[3], m_stream));
    CUDA_OK(cudaStreamSynchronize(m_stream));
    CUDA_OK(cudaStreamSynchronize(m_copy));
    m_early_done = 0;
    float ms = 0;
    CUDA_OK(cudaEventElapsedTime(&ms, m_ev[2], m_ev[3]));
    m_ms[T_D2H] = ms;
    if (m_overlap) {  // stage() did not wait: the exposed part of the host -> device copies is read here
        CUDA_OK(cudaEventElapsedTime(&ms, m_ev[0], m_ev[1]));
        m_ms[T_H2D] = ms;
    }
}

void DeviceEngine::fetch_los_optical_depth(double* host) {
    CUDA_OK(cudaSetDevice(m_device));
    if (!m_is_limb || !d_los_od) throw std::runtime_error("line-of-sight optical depths are produced by the spherical path only");
    CUDA_OK(cudaMemcpyAsync(host, d_los_od, sizeof(double) * (size_t)m_nw * m_nrad, cudaMemcpyDeviceToHost, m_stream));
    CUDA_OK(cudaStreamSynchronize(m_stream));
}

// Final gather of a wavelength-sharded solve: every rank holds its block's results on its device
// (d_radiance [nw_r][nlos], per mapping [nout][nw_r][nlos], per surface mapping [nw_r][nlos]).  One NCCL group: the
// other ranks send their arrays to the root over NVLink, the root receives them back to back into one staging buffer
// and scatters every block into the caller's full-spectrum host arrays (strided rows for the weighting functions).
void DeviceEngine::gather_to_root(Comm& comm, int root, const int* block_start, const int* block_count, int nw_total,
                                  double* radiance_host, double* const* mapping_host, double* const* surface_host,
                                  double ms_out[2]) {
    CUDA_OK(cudaSetDevice(m_device));
    const int rank = comm.rank(), world = comm.world();
    const size_t nlos = (size_t)std::max(m_nrad, 0);
    if (block_count[rank] != m_nw) throw std::runtime_error("gather_to_root: this rank's block size differs from the staged range");
    size_t per_w = nlos;  // doubles per wavelength over all outputs
    if (m_wf_on) {
        for (auto& dm : m_maps) per_w += (size_t)dm.host.nout * nlos;
        per_w += m_surfs.size() * nlos;
    }
    CUDA_OK(cudaEventRecord(m_ev[4], m_stream));
    if (rank != root) {
        comm.group_start();
        comm.send(d_radiance, (size_t)m_nw * nlos, root, m_stream);
        if (m_wf_on) {
            for (auto& dm : m_maps) comm.send(dm.out, (size_t)dm.host.nout * m_nw * nlos, root, m_stream);
            for (auto& ds : m_surfs) comm.send(ds.out, (size_t)m_nw * nlos, root, m_stream);
        }
        comm.group_end();
        CUDA_OK(cudaEventRecord(m_ev[5], m_stream));
        CUDA_OK(cudaStreamSynchronize(m_stream));
        float ms = 0;
        CUDA_OK(cudaEventElapsedTime(&ms, m_ev[4], m_ev[5]));
        ms_out[0] = ms;
        ms_out[1] = 0.0;
        return;
    }
    size_t others = 0;
    for (int r = 0; r < world; ++r)
        if (r != root) others += (size_t)block_count[r];
    if (others * per_w > m_cap_gather) {
        if (d_gather) cudaFree(d_gather);
        d_gather = dalloc<double>(others * per_w);
        m_cap_gather = others * per_w;
    }
    // staging layout per source rank: radiance | mapping 0 | mapping 1 | ... | surface 0 | ...
    comm.group_start();
    {
        double* p = d_gather;
        for (int r = 0; r < world; ++r) {
            if (r == root) continue;
            const size_t n = (size_t)block_count[r];
            comm.recv(p, n * nlos, r, m_stream);
            p += n * nlos;
            if (m_wf_on) {
                for (auto& dm : m_maps) {
                    comm.recv(p, (size_t)dm.host.nout * n * nlos, r, m_stream);
                    p += (size_t)dm.host.nout * n * nlos;
                }
                for (size_t i = 0; i < m_surfs.size(); ++i) {
                    comm.recv(p, n * nlos, r, m_stream);
                    p += n * nlos;
                }
            }
        }
    }
    comm.group_end();
    CUDA_OK(cudaEventRecord(m_ev[5], m_stream));
    // device -> host: every block into its rows of the full-spectrum arrays
    const size_t pitch = sizeof(double) * (size_t)nw_total * nlos;
    auto put = [&](int r, const double* rad, const double* const* maps, const double* const* surfs) {
        const size_t n = (size_t)block_count[r], off = (size_t)block_start[r] * nlos, width = sizeof(double) * n * nlos;
        if (n == 0) return;
        CUDA_OK(cudaMemcpyAsync(radiance_host + off, rad, width, cudaMemcpyDeviceToHost, m_stream));
        if (!m_wf_on) return;
        for (size_t i = 0; i < m_maps.size(); ++i)
            CUDA_OK(cudaMemcpy2DAsync(mapping_host[i] + off, pitch, maps[i], width, width, m_maps[i].host.nout,
                                      cudaMemcpyDeviceToHost, m_stream));
        for (size_t i = 0; i < m_surfs.size(); ++i)
            CUDA_OK(cudaMemcpyAsync(surface_host[i] + off, surfs[i], width, cudaMemcpyDeviceToHost, m_stream));
    };
    std::vector<const double*> mp(m_maps.size()), sp(m_surfs.size());
    for (size_t i = 0; i < m_maps.size(); ++i) mp[i] = m_maps[i].out;
    for (size_t i = 0; i < m_surfs.size(); ++i) sp[i] = m_surfs[i].out;
    put(root, d_radiance, mp.data(), sp.data());
    {
        const double* p = d_gather;
        for (int r = 0; r < world; ++r) {
            if (r == root) continue;
            const size_t n = (size_t)block_count[r];
            const double* rad = p;
            p += n * nlos;
            if (m_wf_on) {
                for (size_t i = 0; i < m_maps.size(); ++i) {
                    mp[i] = p;
                    p += (size_t)m_maps[i].host.nout * n * nlos;
                }
                for (size_t i = 0; i < m_surfs.size(); ++i) {
                    sp[i] = p;
                    p += n * nlos;
                }
            }
            put(r, rad, mp.data(), sp.data());
        }
    }
    CUDA_OK(cudaEventRecord(m_ev[6], m_stream));
    CUDA_OK(cudaStreamSynchronize(m_stream));
    float a = 0, b = 0;
    CUDA_OK(cudaEventElapsedTime(&a, m_ev[4], m_ev[5]));
    CUDA_OK(cudaEventElapsedTime(&b, m_ev[5], m_ev[6]));
    ms_out[0] = a;
    ms_out[1] = b;
}

// One call = copy in, solve, copy out, with the copies of all but one chunk hidden behind the kernels
// (SK_B200_OVERLAP=0 serialises them).
void DeviceEngine::calculate(const AtmosphereArrays& atm, int w0, int nw, double* radiance_host, const WfRequest* wf) {
    static const bool overlap = [] {
        const char* e = std::getenv("SK_B200_OVERLAP");
        return !(e && e[0] == '0');
    }();
    m_overlap = overlap;
    m_early_radiance = radiance_host;
    m_early_done = 0;
    try {
        stage(atm, w0, nw, wf);
        solve_staged();
        fetch(radiance_host);
    } catch (...) {
        m_overlap = false;
        m_tail_pending = false;
        m_early_done = 0;
        cudaStreamSynchronize(m_copy);
        cudaStreamSynchronize(m_stream);
        throw;
    }
    m_overlap = false;
}

}  // namespace disco

namespace disco {
size_t DeviceEngine::debug_copy(const char* name, double* host, size_t max_n) {
    for (auto& e : m_dbg)
        if (e.first == name) {
            const size_t n = std::min(max_n, e.second.second);
            cudaMemcpy(host, e.second.first, n * sizeof(double), cudaMemcpyDeviceToHost);
            return n;
        }
    return 0;
}
}  // namespace disco
