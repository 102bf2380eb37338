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
#include "../sasktran2_b200/csrc/disco_plan.h"

using namespace disco;

static std::string g_err;

template <int N>
static void bvp_emul(const ChunkView& V, int w, int ms, unsigned* status) {
    constexpr int NC = 2 * N, ROWS = 3 * N, ROWLEN = 4 * N + 1;
    constexpr int GL = ROWS <= 4 ? 4 : (ROWS <= 8 ? 8 : (ROWS <= 16 ? 16 : 32));
    constexpr int R = (ROWS + GL - 1) / GL;
    constexpr int NSLOT = GL * R;
    const int L = V.T.L, M = V.M;
    const int m = V.m_list[ms];
    const size_t lay0 = ((size_t)w * M + ms) * L;
    const double* Wp = V.Wp + lay0 * N * N;
    const double* Wm = V.Wm + lay0 * N * N;
    const double* kth = V.kth + lay0 * 2 * N;
    const double* G = V.G + lay0 * 4 * N;
    double* fac = V.fac + lay0 * NC * ROWLEN;
    double* xout = V.xsol + lay0 * NC;
    std::vector<double> a((size_t)NSLOT * ROWLEN, 0.0);
    std::vector<char> act(NSLOT, 0);
    for (int sid = 0; sid < N; ++sid) {
        act[sid] = 1;
        for (int j = 0; j < N; ++j) {
            a[sid * ROWLEN + j] = Wp[sid * N + j];
            a[sid * ROWLEN + N + j] = Wm[sid * N + j] * kth[N + j];
        }
        a[sid * ROWLEN + 4 * N] = -G[sid];
    }
    std::vector<double> facs(NC * ROWLEN), xs(NC, 0.0);
    for (int p = 0; p < L; ++p) {
        const bool last = (p == L - 1);
        const int needed = last ? N : NC;
        int rank = 0;
        std::vector<char> wasfree(NSLOT);
        for (int sid = 0; sid < NSLOT; ++sid) wasfree[sid] = !act[sid];
        for (int sid = 0; sid < NSLOT; ++sid) {
            if (!wasfree[sid]) continue;
            if (rank < needed) {
                act[sid] = 1;
                double* row = &a[sid * ROWLEN];
                const double* Wpu = Wp + (size_t)p * N * N;
                const double* Wmu = Wm + (size_t)p * N * N;
                const double* thu = kth + (size_t)p * 2 * N + N;
                const double* Gu = G + (size_t)p * 4 * N;
                if (!last) {
                    const double* Wpl = Wpu + N * N;
                    const double* Wml = Wmu + N * N;
                    const double* thl = thu + 2 * N;
                    const double* Gl = Gu + 4 * N;
                    const bool first = rank < N;
                    const int i = first ? rank : rank - N;
                    const double* A1 = first ? Wmu : Wpu;
                    const double* A2 = first ? Wpu : Wmu;
                    const double* B1 = first ? Wml : Wpl;
                    const double* B2 = first ? Wpl : Wml;
                    for (int j = 0; j < N; ++j) {
                        row[j] = A1[i * N + j] * thu[j];
                        row[N + j] = A2[i * N + j];
                        row[2 * N + j] = -B1[i * N + j];
                        row[3 * N + j] = -(B2[i * N + j] * thl[j]);
                    }
                    row[4 * N] = first ? (-Gu[3 * N + i] + Gl[N + i]) : (-Gu[2 * N + i] + Gl[i]);
                } else {
                    const int i = rank;
                    const bool refl = (m == 0);
                    const double alb2 = refl ? 2.0 * V.albedo[w] : 0.0;
                    const double* surf = V.surf + (size_t)w * (2 * N + 1);
                    for (int j = 0; j < N; ++j) {
                        double vm = Wmu[i * N + j], vp = Wpu[i * N + j];
                        if (refl) {
                            vm -= alb2 * surf[j];
                            vp -= alb2 * surf[N + j];
                        }
                        row[j] = vm * thu[j];
                        row[N + j] = vp;
                        row[2 * N + j] = 0.0;
                        row[3 * N + j] = 0.0;
                    }
                    double rhs = -Gu[3 * N + i];
                    if (refl) {
                        rhs += alb2 * surf[2 * N];
                        rhs += V.T.csz * V.albedo[w] / kPi * V.lay_trans[(size_t)w * (L + 1) + L];
                    }
                    row[4 * N] = rhs;
                }
            }
            ++rank;
        }
        for (int c = 0; c < NC; ++c) {
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
        std::copy(facs.begin(), facs.end(), fac + (size_t)p * NC * ROWLEN);
        if (!last)
            for (int sid = 0; sid < NSLOT; ++sid)
                for (int j = 0; j < NC; ++j) {
                    a[sid * ROWLEN + j] = a[sid * ROWLEN + NC + j];
                    a[sid * ROWLEN + NC + j] = 0.0;
                }
    }
    for (int p = L - 1; p >= 0; --p) {
        const double* f = fac + (size_t)p * NC * ROWLEN;
        std::vector<double> acc(NC), x(NC);
        for (int c = 0; c < NC; ++c) {
            acc[c] = f[c * ROWLEN + 4 * N];
            if (p < L - 1)
                for (int j = 0; j < NC; ++j) acc[c] -= f[c * ROWLEN + NC + j] * xs[j];
        }
        for (int cc = NC - 1; cc >= 0; --cc) {
            x[cc] = acc[cc] / f[cc * ROWLEN + cc];
            for (int c = 0; c < cc; ++c) acc[c] -= f[c * ROWLEN + cc] * x[cc];
        }
        for (int c = 0; c < NC; ++c) {
            xs[c] = x[c];
            xout[(size_t)p * NC + c] = x[c];
        }
    }
}

template <int N>
static void run_all(ChunkView& V, unsigned* status) {
    const int L = V.T.L;
    for (long long i = 0; i < (long long)V.nw * L; ++i) optics_body(V, i);
    for (int w = 0; w < V.nw; ++w) beam_body(V, w);
    for (long long i = 0; i < (long long)V.nw * V.M * L; ++i) layer_problem_body<N>(V, i);
    for (int w = 0; w < V.nw; ++w)
        for (int ms = 0; ms < V.M; ++ms) bvp_emul<N>(V, w, ms, status);
    for (long long i = 0; i < (long long)V.nw * V.T.nlos; ++i) radiance_body(V, i);
}

extern "C" const char* emul_last_error() { return g_err.c_str(); }

extern "C" int emul_do_radiance(int nstr, int nloc, int nwavel, int nleg, int nlos, const double* alt, int interp,
                                int geotype, double cos_sza, double earth_radius, const double* los_cos_vza,
                                const double* los_rel_az, const double* ssa, const double* ext, const double* leg,
                                const double* solar, const double* albedo, int include_ss, double* radiance,
                                int* num_azimuth_solved) {
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
        V.Wp = A(c * M * L * N * N); V.Wm = A(c * M * L * N * N); V.kth = A(c * M * L * 2 * N);
        V.G = A(c * M * L * 4 * N); V.surf = A(c * (2 * N + 1)); V.wvec = A(c * M * nlos * L * 2 * N);
        V.vsrc = A(c * M * nlos * L); V.xsol = A(c * M * L * 2 * N); V.fac = A(c * M * L * 2 * N * (4 * N + 1));
        V.radiance = radiance;
        unsigned status = 0;
        V.status = &status;
        switch (N) {
            case 1: run_all<1>(V, &status); break;
            case 2: run_all<2>(V, &status); break;
            case 4: run_all<4>(V, &status); break;
            case 8: run_all<8>(V, &status); break;
            case 16: run_all<16>(V, &status); break;
            default: throw std::runtime_error("unsupported nstr");
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
