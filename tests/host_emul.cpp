// TEST HARNESS (not shipped, not a fallback): runs the product's per-thread kernel bodies
// (sasktran2_b200/csrc/disco_bodies.h, disco_core.h) on the CPU so their arithmetic can be checked against
// the oracle without a GPU.  The warp-cooperative BVP kernel cannot run on the host; it is mirrored here by a
// lane-serial transcription of the same staircase elimination (same slots, ranks, pivot-row storage and back
// substitution), which checks the algorithm, leaving only the SIMT mechanics to the GPU tests.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../sasktran2_b200/csrc/disco_bodies.h"
#include "../sasktran2_b200/csrc/disco_bvp_rows.h"
#include "../sasktran2_b200/csrc/disco_wf_body.h"
#include "../sasktran2_b200/csrc/disco_plan.h"
#include "../sasktran2_b200/csrc/disco_brdf.h"

using namespace disco;

static std::string g_err;
static std::vector<std::pair<std::string, std::vector<double>>> g_dump;

// Lane-serial transcription of staircase_solve (sasktran2_b200/csrc/disco_bvp.cuh): same slots, ranks, pivot
// search, pivot-row storage and back substitution, driven by the same row loaders as the CUDA kernels.
template <int N, class Prob>
static void staircase_emul(const Prob& prob, unsigned* status) {
    constexpr int NRHS = Prob::NRHS;
    constexpr int NC = 2 * N, ROWS = 3 * N, ROWLEN = 4 * N + NRHS;
    constexpr int GL = ROWS <= 4 ? 4 : (ROWS <= 8 ? 8 : (ROWS <= 16 ? 16 : 32));
    constexpr int R = (ROWS + GL - 1) / GL;
    constexpr int NSLOT = GL * R;
    const int nsteps = prob.nsteps();
    std::vector<double> a((size_t)NSLOT * ROWLEN, 0.0);
    std::vector<char> act(NSLOT, 0);
    std::vector<double> fac((size_t)nsteps * NC * ROWLEN, 0.0), facs(NC * ROWLEN), xs(NRHS * NC, 0.0);
    for (int step = 0; step < nsteps; ++step) {
        const int needed = prob.nnew(step);
        int rank = 0;
        std::vector<char> wasfree(NSLOT);
        for (int sid = 0; sid < NSLOT; ++sid) wasfree[sid] = !act[sid];
        for (int sid = 0; sid < NSLOT; ++sid) {
            if (!wasfree[sid]) continue;
            if (rank < needed) {
                act[sid] = 1;
                prob.load(step, rank, &a[sid * ROWLEN]);
            }
            ++rank;
        }
        const int nleft = prob.nleft(step);
        for (int c = 0; c < nleft; ++c) {
            double best = -1.0;
            int bsid = -1;
            for (int sid = 0; sid < NSLOT; ++sid)
                if (act[sid] && std::fabs(a[sid * ROWLEN + c]) > best) {
                    best = std::fabs(a[sid * ROWLEN + c]);
                    bsid = sid;
                }
            if (!(best > 0.0)) {
                *status |= 4u;
                return;
            }
            act[bsid] = 0;
            for (int cc = 0; cc < ROWLEN; ++cc) facs[c * ROWLEN + cc] = (cc >= c) ? a[bsid * ROWLEN + cc] : 0.0;
            const double pinv = 1.0 / facs[c * ROWLEN + c];
            for (int sid = 0; sid < NSLOT; ++sid)
                if (act[sid]) {
                    const double f = a[sid * ROWLEN + c] * pinv;
                    for (int cc = c + 1; cc < ROWLEN; ++cc) a[sid * ROWLEN + cc] -= f * facs[c * ROWLEN + cc];
                    a[sid * ROWLEN + c] = 0.0;
                }
        }
        std::copy(facs.begin(), facs.begin() + nleft * ROWLEN, fac.begin() + (size_t)step * NC * ROWLEN);
        if (step < nsteps - 1)
            for (int sid = 0; sid < NSLOT; ++sid)
                for (int j = 0; j < NC; ++j) {
                    a[sid * ROWLEN + j] = a[sid * ROWLEN + NC + j];
                    a[sid * ROWLEN + NC + j] = 0.0;
                }
    }
    for (int step = nsteps - 1; step >= 0; --step) {
        const int nleft = prob.nleft(step), nright = prob.nright(step);
        const double* f = &fac[(size_t)step * NC * ROWLEN];
        for (int r = 0; r < NRHS; ++r) {
            std::vector<double> acc(NC, 0.0), x(NC, 0.0);
            for (int c = 0; c < nleft; ++c) {
                acc[c] = f[c * ROWLEN + 4 * N + r];
                for (int j = 0; j < nright; ++j) acc[c] -= f[c * ROWLEN + NC + j] * xs[r * NC + j];
            }
            for (int cc = nleft - 1; cc >= 0; --cc) {
                x[cc] = acc[cc] * (1.0 / f[cc * ROWLEN + cc]);
                for (int c = 0; c < cc; ++c) acc[c] -= f[c * ROWLEN + cc] * x[cc];
            }
            for (int c = 0; c < nleft; ++c) {
                xs[r * NC + c] = x[c];
                prob.store(step, c, r, x[c]);
            }
        }
    }
}

template <int N>
static void bvp_emul(const ChunkView& V, int w, int ms, unsigned* status) {
    ForwardRows<N> rows(V, w, ms);
    staircase_emul<N>(rows, status);
}

template <int N>
static void bvp_adjoint_emul(const ChunkView& V, int w, int ms, unsigned* status) {
    const int nlos = V.T.nlos;
    if (nlos <= 4) {
        for (int los0 = 0; los0 < nlos; los0 += 4) {
            AdjointRows<N, 4> rows(V, w, ms, los0);
            staircase_emul<N>(rows, status);
        }
    } else {
        for (int los0 = 0; los0 < nlos; los0 += 10) {
            AdjointRows<N, 10> rows(V, w, ms, los0);
            staircase_emul<N>(rows, status);
        }
    }
}

template <int N, int G>
static void run_wf(ChunkView& V, unsigned* status) {
    for (int w = 0; w < V.nw; ++w)
        for (int ms = 0; ms < V.M; ++ms) bvp_adjoint_emul<N>(V, w, ms, status);
    for (long long i = 0; i < (long long)V.nw * V.M * V.T.L; ++i) wf_layer_body<N, G>(V, i);
    if (V.gsurf_rows)
        for (long long i = 0; i < (long long)V.nw * V.T.nlos; ++i) wf_ground_reduce_body(V, i);
    for (long long i = 0; i < (long long)V.nw * V.T.nlos; ++i) wf_chain_body(V, i, G);
}

static const BrdfView* g_brdf_view = nullptr;   // set by emul_do_radiance when a MODIS surface was requested

template <int N>
static void run_all(ChunkView& V, unsigned* status, bool wf) {
    const int L = V.T.L;
    for (long long i = 0; i < (long long)V.nw * L; ++i) optics_body(V, i);
    for (int w = 0; w < V.nw; ++w) beam_body(V, w);
    for (long long i = 0; i < (long long)V.nw * V.M * L; ++i) layer_problem_body<N>(V, i);
    if (g_brdf_view)   // kernel-based surface: one "thread" per (wavelength, order, stream / LOS row)
        for (int w = 0; w < V.nw; ++w)
            for (int ms = 0; ms < V.M; ++ms)
                for (int t = 0; t < V.T.N + V.T.nlos; ++t) surface_general_body(V, *g_brdf_view, w, ms, t);
    for (int w = 0; w < V.nw; ++w)
        for (int ms = 0; ms < V.M; ++ms) bvp_emul<N>(V, w, ms, status);
    for (long long i = 0; i < (long long)V.nw * V.T.nlos; ++i) radiance_body(V, i);
    if (wf) {
        switch (V.ngroups) {
            case 0: run_wf<N, 0>(V, status); break;
            case 1: run_wf<N, 1>(V, status); break;
            case 2: run_wf<N, 2>(V, status); break;
            default: throw std::runtime_error("unsupported number of scattering groups");
        }
    }
}

extern "C" const char* emul_last_error() { return g_err.c_str(); }
extern "C" long long emul_dump(const char* name, double* out, long long max_n) {
    for (auto& e : g_dump)
        if (e.first == name) {
            long long n = std::min<long long>(max_n, (long long)e.second.size());
            std::copy(e.second.begin(), e.second.begin() + n, out);
            return n;
        }
    return 0;
}

// thermal sources of the next emul_do_radiance call: emission_source [nloc, nwavel] and surface emission [nwavel]
static const double* g_emission = nullptr;
static const double* g_semis = nullptr;
extern "C" void emul_set_emission(const double* emission, const double* surface_emission) {
    g_emission = emission;
    g_semis = surface_emission;
}

// MODIS surface of the next emul_do_radiance call: args [3, nwavel] column-major (null: Lambertian)
static const double* g_modis_args = nullptr;
extern "C" void emul_set_modis(const double* args) { g_modis_args = args; }

extern "C" int emul_do_radiance(int nstr, int nloc, int nwavel, int nleg, int nlos, const double* alt, int interp,
                                int geotype, double cos_sza, double earth_radius, const double* los_cos_vza,
                                const double* los_rel_az, const double* ssa, const double* ext, const double* leg,
                                const double* solar, const double* albedo, int include_ss, double* radiance,
                                int* num_azimuth_solved, const double* d_leg, int ngroups, double* native) {
    try {
        GeometrySpec geo;
        geo.altitudes.assign(alt, alt + nloc);
        geo.interp = interp;
        geo.geotype = geotype;
        geo.cos_sza = cos_sza;
        geo.earth_radius = earth_radius;
        std::vector<LineOfSight> los(nlos);
        for (int j = 0; j < nlos; ++j) los[j] = {los_cos_vza[j], los_rel_az[j], alt[nloc - 1] + 1.0};
        HostPlan P = build_plan(nstr, geo, los);
        std::vector<int> mlist;
        for (int m = 0; m < nstr; ++m) {
            bool any = false;
            for (int j = 0; j < nlos && !any; ++j)
                for (int l = 0; l < nstr && !any; ++l)
                    if (P.lp_los[((size_t)j * nstr + m) * nstr + l] != 0.0) any = true;
            if (any) mlist.push_back(m);
        }
        if (mlist.empty()) mlist.push_back(0);
        if (num_azimuth_solved) *num_azimuth_solved = (int)mlist.size();
        const size_t N = P.N, L = P.L, M = mlist.size(), c = nwavel;
        ChunkView V{};
        V.T.nstr = nstr; V.T.N = (int)N; V.T.L = (int)L; V.T.nloc = nloc; V.T.nlos = nlos; V.T.csz = P.csz;
        V.T.mu = P.mu.data(); V.T.wt = P.wt.data(); V.T.lp_mu = P.lp_mu.data(); V.T.lp_csz = P.lp_csz.data();
        V.T.lp_los = P.lp_los.data(); V.T.los_mu = P.los_mu.data(); V.T.los_cosmphi = P.los_cosmphi.data();
        V.layer_dh = P.layer_dh.data(); V.interp_idx = P.interp_idx.data(); V.interp_w = P.interp_w.data();
        V.chapman = P.chapman.data(); V.plane_parallel = P.plane_parallel;
        V.nw = nwavel; V.nleg = nleg; V.ext = ext; V.ssa = ssa; V.leg = leg; V.albedo = albedo; V.solar = solar;
        V.include_ss = include_ss; V.M = (int)M; V.m_list = mlist.data();
        std::vector<std::vector<double>> store;
        auto A = [&](size_t n) { store.emplace_back(n, 0.0); return store.back().data(); };
        V.lay_od = A(c * L); V.lay_ssa = A(c * L); V.lay_beta = A(c * L * nstr); V.lay_secant = A(c * L);
        V.lay_trans = A(c * (L + 1)); V.lay_cumod = A(c * (L + 1)); V.lay_totext = A(c * L); V.lay_scatext = A(c * L);
        V.lay_thermal = A(c * L * 2); V.emission = g_emission; V.semis = g_semis;
        V.Wp = A(c * M * L * N * N); V.Wm = A(c * M * L * N * N); V.kth = A(c * M * L * 2 * N);
        V.G = A(c * M * L * 4 * N); V.surf = A(c * (2 * N + 1)); V.wvec = A(c * M * nlos * L * 2 * N);
        V.vsrc_w = 1;
        V.vsrc = A(c * M * nlos * L); V.xsol = A(c * M * L * 2 * N); V.fac = A(c * M * L * 2 * N * (4 * N + 1));
        V.radiance = radiance;
        const bool wf = native != nullptr;
        const size_t G = wf ? ngroups : 0;
        V.ngroups = (int)G;
        V.dleg = d_leg;
        V.dleg_gstride = (size_t)nleg * nloc * nwavel;
        if (wf) {
            V.zadj = A(c * M * nlos * 2 * N * L);
            V.lay_dbeta = A(c * L * (G ? G : 1) * nstr);
            V.wf_loc = A(c * M * nlos * L * (G + 4));
            V.wf_src = A(c * M * nlos * L);
            V.wf_gnd = A(c * nlos * 3);
            V.wf_scratch = A(c * nlos * 3 * (L + 1));
            V.wf_native = native;
        }
        BrdfTables BT;
        BrdfView BV{};
        std::vector<double> zero_albedo(nwavel, 0.0);
        g_brdf_view = nullptr;
        if (g_modis_args) {
            BT = build_brdf_tables(kBrdfModis, P);
            BV.nk = BT.nk; BV.nargs = 3; BV.Rss = BT.Rss.data(); BV.rsun = BT.rsun.data(); BV.Rls = BT.Rls.data();
            BV.rlsun = BT.rlsun.data(); BV.args = g_modis_args;
            V.albedo = zero_albedo.data();
            V.gsurf_stride = (int)(2 * N * N + 2 * N);
            V.gsurf_out = A(c * M * V.gsurf_stride);
            V.gsurf = V.gsurf_out;
            if (wf) {
                V.gsurf_rows = A(c * M * (N + nlos) * (N + 1));
                V.brdf_Rss = BT.Rss.data(); V.brdf_rsun = BT.rsun.data(); V.brdf_Rls = BT.Rls.data(); V.brdf_rlsun = BT.rlsun.data();
                V.brdf_nk = BT.nk;
                V.wf_gndk = A(c * nlos * BT.nk);
                V.wf_gnd_part = A(c * M * nlos * (2 + BT.nk));
            }
            g_brdf_view = &BV;
        }
        unsigned status = 0;
        V.status = &status;
        switch (N) {
            case 1: run_all<1>(V, &status, wf); break;
            case 2: run_all<2>(V, &status, wf); break;
            case 4: run_all<4>(V, &status, wf); break;
            case 8: run_all<8>(V, &status, wf); break;
            case 16: run_all<16>(V, &status, wf); break;
            default: throw std::runtime_error("unsupported nstr");
        }
        {
            g_dump.clear();
            auto D = [&](const char* nm, const double* p, size_t n) { if (p) g_dump.push_back({nm, std::vector<double>(p, p + n)}); };
            D("lay_od", V.lay_od, c * L); D("lay_secant", V.lay_secant, c * L); D("lay_trans", V.lay_trans, c * (L + 1));
            D("lay_beta", V.lay_beta, c * L * nstr); D("kth", V.kth, c * M * L * 2 * N); D("Wp", V.Wp, c * M * L * N * N);
            D("G", V.G, c * M * L * 4 * N); D("wvec", V.wvec, c * M * nlos * L * 2 * N); D("vsrc", V.vsrc, c * M * nlos * L);
            D("xsol", V.xsol, c * M * L * 2 * N); D("Wm", V.Wm, c * M * L * N * N); D("lay_ssa", V.lay_ssa, c * L);
            D("lay_cumod", V.lay_cumod, c * (L + 1));
            if (wf) {
                D("zadj", V.zadj, c * M * nlos * 2 * N * L); D("lay_dbeta", V.lay_dbeta, c * L * G * nstr);
                D("wf_loc", V.wf_loc, c * M * nlos * L * (G + 4)); D("wf_src", V.wf_src, c * M * nlos * L);
                D("wf_gnd", V.wf_gnd, c * nlos * 3);
                if (V.gsurf_rows) D("gsurf_rows", V.gsurf_rows, c * M * (N + nlos) * (N + 1));
                if (V.wf_gndk) D("wf_gndk", V.wf_gndk, c * nlos * V.brdf_nk);
            }
        }
        if (status) {
            g_err = "status bits " + std::to_string(status);
            return -3;
        }
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -3;
    }
}

// Dedicated two-stream kernel body (sasktran2_b200/csrc/disco_twostream_body.h) run on the host, one wavelength after the
// other, two lines of sight per pass like k_twostream<2>.  fdm: delta-M fraction [nloc, nwavel] or null.
#include "../sasktran2_b200/csrc/disco_twostream_body.h"
extern "C" int emul_twostream(int nloc, int nwavel, int nleg, int nlos, const double* alt, int interp, int geotype,
                              double cos_sza, double earth_radius, const double* los_cos_vza, const double* los_rel_az,
                              const double* ssa, const double* ext, const double* leg, const double* solar,
                              const double* albedo, const double* fdm, double* radiance) {
    try {
        GeometrySpec geo;
        geo.altitudes.assign(alt, alt + nloc);
        geo.interp = interp;
        geo.geotype = geotype;
        geo.cos_sza = cos_sza;
        geo.earth_radius = earth_radius;
        std::vector<LineOfSight> los(nlos);
        for (int j = 0; j < nlos; ++j) los[j] = {los_cos_vza[j], los_rel_az[j], alt[nloc - 1] + 1.0};
        HostPlan P = build_plan(2, geo, los);
        ChunkView V{};
        V.T.nstr = 2; V.T.N = 1; V.T.L = P.L; V.T.nloc = nloc; V.T.nlos = nlos; V.T.csz = P.csz;
        V.T.los_mu = P.los_mu.data(); V.T.los_cosmphi = P.los_cosmphi.data();
        V.layer_dh = P.layer_dh.data(); V.chapman = P.chapman.data(); V.plane_parallel = P.plane_parallel;
        V.nw = nwavel; V.nleg = nleg; V.ext = ext; V.ssa = ssa; V.leg = leg; V.albedo = albedo; V.solar = solar;
        V.fdm = fdm;
        V.radiance = radiance;
        std::vector<double> od(P.L);
        for (int w = 0; w < nwavel; ++w)
            for (int l0 = 0; l0 < nlos; l0 += 2) disco::ts::twostream_body<2>(V, w, l0, P.chapman.data(), od.data(), 1);
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -3;
    }
}
