// Per-thread bodies of the batched DO kernels, written as host/device functions so that the CUDA kernels
// (disco_kernels.cu) stay thin and the same arithmetic can be exercised on the CPU by tests/host_emul.cpp.
#pragma once
#include <stddef.h>

#include "disco_core.h"

namespace disco {

// Device-side view of one wavelength chunk (all pointers device memory)
struct ChunkView {
    // geometry
    Tables T;
    const double* layer_dh;   // [L]
    const int* interp_idx;    // [L][2]
    const double* interp_w;   // [L][2]
    const double* chapman;    // [L][L]
    int plane_parallel;
    // atmosphere inputs of the chunk (reference C-ABI layouts, wavelength slowest)
    int nw;                   // wavelengths in this chunk
    int nleg;
    const double* ext;        // [nloc, nw]
    const double* ssa;        // [nloc, nw]
    const double* leg;        // [nleg, nloc, nw]
    const double* albedo;     // [nw]
    const double* solar;      // [nw]
    const double* emission;   // [nloc, nw] thermal emission source (emission_source = discrete_ordinates), null: none
    const double* semis;      // [nw] surface emission, null: none
    int include_ss;
    int M;                    // azimuth orders actually solved
    const int* m_list;        // [M] azimuth order of each solved slot
    // per-layer optics
    double* lay_od;           // [nw][L]   optical thickness
    double* lay_ssa;          // [nw][L]
    double* lay_beta;         // [nw][L][nstr]
    double* lay_secant;       // [nw][L]
    double* lay_trans;        // [nw][L+1] beam transmittance at layer boundaries (x solar irradiance)
    double* lay_cumod;        // [nw][L+1] vertical optical depth above each boundary
    double* lay_totext;       // [nw][L]
    double* lay_scatext;      // [nw][L]
    double* lay_thermal;      // [nw][L][2] thermal source b0 exp(-b1 x) of the layer: b0 | b1 (written when emission is set)
    // per (w, m, layer) solution
    double* Wp;               // [nw][M][L][N*N] row-major (stream, solution)
    double* Wm;
    double* kth;              // [nw][M][L][2N]  k | theta
    double* G;                // [nw][M][L][4N]  G+top | G-top | G+bot | G-bot
    double* surf;             // [nw][2N+1]      s+_j | s-_j | sG   (bottom layer, m = 0)
    double* wvec;             // [nw][M][nlos][L][2N]  d(radiance_m)/d(L_j, M_j)
    double* vsrc;             // [nw][M][nlos][L][vsrc_w]  particular + single-scatter (+ ground direct) terms
    double* xsol;             // [nw][M][L][2N]        BVP solution L | M
    double* fac;              // [group][fac_stride]   pivot rows of the staircase LU (forward and adjoint)
    size_t fac_stride;        // doubles per solve group: (L+1) * 2N * (4N + max nrhs)
    double* zadj;             // [nw][M][2N*L][nlos]   adjoint BVP solutions A^T z = wvec (weighting functions), LOS fastest
    double* lfac;             // [nw*M][lfac_stride]   elimination multipliers of the forward LU, [pivot][LS] (null: not kept)
    size_t lfac_stride;       // doubles per problem: L * 2N * LS
    double* yadj;             // [nw][M][2N*L][nlos]   U^T y = wvec intermediate of the transposed solve, LOS fastest
    double* radiance;         // [nw][nlos]
    // ---- weighting functions (null / 0 when not requested)
    int ngroups;              // scattering derivative groups G
    const double* dleg;       // [G][nleg, nloc, nw_staged]: d_leg_coeff of group g (chunk-offset applied)
    size_t dleg_gstride;      // doubles between groups
    const double* fdm;        // [1 + G][nloc, nw_staged]: delta-M fraction f | d_f of group g; null: no scaling applied
    size_t fdm_gstride;       // doubles between the planes of fdm
    double* lay_dbeta;        // [nw][L][G][nstr] layer-level Legendre derivative direction per group
    double* wf_loc;           // [nw][M][nlos][L][G+4]
    double* wf_src;           // [nw][M][nlos][L]
    double* wf_gnd;           // [nw][nlos][3]
    double* wf_native;        // [nw][nlos][nloc*(2+G)+1]
    double* wf_scratch;       // [nw][nlos][3][L+1]
    // ---- register-resident fast path (N <= 8): [element][problem] staging planes of the eigen-solve
    double* eigS;             // [N(N+1)/2][nw*M*L] packed symmetric S~+
    double* eigH;             // [N(N+1)/2][nw*M*L] packed Cholesky factor of S~-
    double* eigC;             // [N(N+1)/2][nw*M*L] packed symmetric C = H^T S~+ H
    double* los_att;          // [nw][nlos][L+1]  exp(-cum_od / mu_los)
    double* los_lay;          // [nw][nlos][L][3] exp(-od / mu_los) | E | 1 / (1 + mu secant)
    int vsrc_w;               // entries of vsrc per (w, m, los, layer): 1 (generic path) or N (per-solution partials)
    unsigned int* status;     // error bits
    // ---- kernel-based (non-Lambertian) surface: null for the Lambertian closed form.  With gsurf set, `albedo` points
    //      at zeros, so every Lambertian term of the layer kernels and row loaders vanishes (disco_brdf.h).
    const double* gsurf;      // [nw][M][2N^2 + 2N]: SP[i][j] | SM[i][j] | SG[i] | rho_m(mu_i, mu_0)
    double* gsurf_out;        // the same array, writable (k_surface_general)
    int gsurf_stride;
    // weighting functions above such a surface: [nw][M][N + nlos][N + 1] reflection rows of the streams and the LOS:
    // (1 + delta_m0) w_q mu_q rho_m(row, mu_q) | rho_m(row, mu_0)   (null: radiances only)
    double* gsurf_rows;
    int wf_bottom_only;       // 1: k_wf_layer solves only the layer on the ground (k_wf_layer_fast leaves it out)
    // weighting functions w.r.t. the weights of a linear kernel model (MODIS): the kernels' own Fourier tables
    // (disco_brdf.h layouts) and the accumulated derivatives wf_gndk[nw][nlos][brdf_nk]; null: not requested
    const double *brdf_Rss, *brdf_rsun, *brdf_Rls, *brdf_rlsun;
    int brdf_nk;
    double* wf_gndk;
    // per-order ground pieces above a kernel-based BRDF, [nw][M][nlos][2 + brdf_nk]: d/dT_floor | ground term | weights;
    // summed over the orders in a fixed order by wf_ground_reduce_body (bit-reproducible, no floating-point atomics)
    double* wf_gnd_part;
    // ---- several solar zenith angles sharing one homogeneous solution and one factorisation (spherical path): the
    //      arrays that depend on the SZA exist nsza times; slice s starts s * stride doubles after the pointers above
    int nsza;                 // 0 / 1: single SZA
    size_t sza_G, sza_surf, sza_trans, sza_xsol;   // strides of G, surf, lay_trans, xsol
    double sza_csz[4];        // cos(SZA) of every slice
};


DISCO_HD void raise_status(unsigned int* status, unsigned int bits) {
#if defined(__CUDA_ARCH__)
    atomicOr(status, bits);
#else
    *status |= bits;
#endif
}

// K1 body: one (wavelength, layer)
DISCO_HD void optics_body(const ChunkView& V, long long idx) {
    const int L = V.T.L, nstr = V.T.nstr, nloc = V.T.nloc;
    const int w = (int)(idx / L), p = (int)(idx % L);
    const double* ext = V.ext + (size_t)nloc * w;
    const double* ssa = V.ssa + (size_t)nloc * w;
    const double* leg = V.leg + (size_t)V.nleg * nloc * w;
    double od = 0.0, sc = 0.0;
    double* beta = V.lay_beta + (size_t)idx * nstr;
    for (int l = 0; l < nstr; ++l) beta[l] = 0.0;
    for (int c = 0; c < 2; ++c) {
        const int q = V.interp_idx[p * 2 + c];
        if (q < 0) continue;
        const double wgt = V.interp_w[p * 2 + c];
        const double kext = ext[q];
        const double kscat = ssa[q] * kext;
        od += kext * wgt;
        sc += kscat * wgt;
        const int nl = nstr < V.nleg ? nstr : V.nleg;
        if (V.fdm) {
            // delta-M: the truncated forward peak leaves every moment (sktran_do_layerarray.cpp:396-404)
            const double f = V.fdm[(size_t)nloc * w + q];
            const double ff = f / (1.0 - f);
            for (int l = 0; l < nstr; ++l)
                beta[l] += wgt * kscat * ((l < nl ? leg[l + (size_t)V.nleg * q] : 0.0) - (2 * l + 1) * ff);
        } else {
            for (int l = 0; l < nl; ++l) beta[l] += wgt * kscat * leg[l + (size_t)V.nleg * q];
        }
    }
    if (sc > 0.0) {
        for (int l = 0; l < nstr; ++l) beta[l] /= sc;
    } else {
        beta[0] = 0.0;
    }
    if (V.ngroups > 0) {
        // Layer-level derivative direction of the Legendre moments for each scattering group.  The reference
        // overwrites it for every contributing grid point, so the last (highest) one wins
        // (sktran_do_layerarray.cpp:761-800).  With delta-M scaling applied the direction also carries the
        // truncated peak and d_f (:773, :792-800).
        int ql = V.interp_idx[p * 2 + 1] >= 0 ? V.interp_idx[p * 2 + 1] : V.interp_idx[p * 2];
        const double f = (V.fdm && ql >= 0) ? V.fdm[(size_t)nloc * w + ql] : 0.0;
        for (int g = 0; g < V.ngroups; ++g) {
            const double* dl = V.dleg + g * V.dleg_gstride + (size_t)V.nleg * nloc * w;
            double* db = V.lay_dbeta + ((size_t)idx * V.ngroups + g) * nstr;
            const double df = (V.fdm && ql >= 0) ? V.fdm[(1 + g) * V.fdm_gstride + (size_t)nloc * w + ql] : 0.0;
            for (int l = 0; l < nstr; ++l) {
                const double ph = (l < V.nleg && ql >= 0) ? leg[l + (size_t)V.nleg * ql] : 0.0;
                const double dv = (l < V.nleg && ql >= 0) ? dl[l + (size_t)V.nleg * ql] : 0.0;
                double d = dv + (ph - (2 * l + 1) * f / (1.0 - f) - beta[l]);
                if (V.fdm) d += -(2 * l + 1) / (1.0 - f) / (1.0 - f) * df;
                db[l] = d;
            }
        }
    }
    double ssa_l = sc / od;
    const double dh = V.layer_dh[p];
    od *= dh;
    if (V.emission) {
        // thermal source S(x) = b0 exp(-b1 x), x from the layer top: emission at the highest / lowest grid point that
        // contributes to the layer (sktran_do_layerarray.cpp:341-370, 459-470)
        const double* em = V.emission + (size_t)nloc * w;
        int min_q = -1, max_q = -1;
        for (int c = 0; c < 2; ++c) {
            const int q = V.interp_idx[p * 2 + c];
            if (q < 0 || !(V.interp_w[p * 2 + c] > 0.0)) continue;
            if (min_q < 0 || q < min_q) min_q = q;
            if (max_q < 0 || q > max_q) max_q = q;
        }
        const double b0_top = max_q >= 0 ? em[max_q] : 0.0;
        const double b0_bot = (min_q >= 0 && min_q != max_q) ? em[min_q] : b0_top;
        double b1 = 0.0;
        if (od > 1e-10 && b0_top > 1e-30 && b0_bot > 1e-30 && fabs(b0_top - b0_bot) > 1e-15 * fmax(b0_top, b0_bot))
            b1 = log(b0_top / b0_bot) / od;
        V.lay_thermal[(size_t)idx * 2] = b0_top;
        V.lay_thermal[(size_t)idx * 2 + 1] = b1;
    }
    const double total_ext = od / dh;
    double scat_ext = total_ext * ssa_l;
    scat_ext = fmax(scat_ext, total_ext * kSsaDither);
    double m_ssa = scat_ext / total_ext;
    if (1.0 - m_ssa < kSsaDither) m_ssa = 1.0 - kSsaDither;
    V.lay_od[idx] = od;  // raw; k_beam turns it into the reference's floor - ceiling difference
    V.lay_ssa[idx] = m_ssa;
    V.lay_totext[idx] = total_ext;
    V.lay_scatext[idx] = scat_ext;
}

// K1b body: one wavelength
DISCO_HD void beam_body(const ChunkView& V, int w) {
    const int L = V.T.L;
    double* od = V.lay_od + (size_t)w * L;
    double* cum = V.lay_cumod + (size_t)w * (L + 1);
    double* sec = V.lay_secant + (size_t)w * L;
    double* tr = V.lay_trans + (size_t)w * (L + 1);
    double ceiling = 0.0, floor_d = 0.0;
    cum[0] = 0.0;
    for (int p = 0; p < L; ++p) {
        floor_d += od[p];
        od[p] = floor_d - ceiling;  // M_OPTICAL_THICKNESS (sktran_do_opticallayer.cpp:21)
        ceiling = floor_d;
        cum[p + 1] = floor_d;
    }
    const double f0 = V.solar[w];
    tr[0] = f0;
    double prev = 0.0;
    for (int p = 0; p < L; ++p) {
        double slant = 0.0;
        const double* ch = V.chapman + (size_t)p * L;
        for (int q = 0; q <= p; ++q) slant += ch[q] * od[q];
        sec[p] = (slant - prev) / od[p];
        tr[p + 1] = exp(-slant) * f0;
        prev = slant;
    }
}

// K2 body: one (wavelength, azimuth slot, layer)
template <int N>
DISCO_HD void layer_problem_body(const ChunkView& V, long long idx) {
    constexpr int NSTR = 2 * N;
    const int L = V.T.L, M = V.M, nlos = V.T.nlos;
    const int p = (int)(idx % L);
    const int ms = (int)((idx / L) % M);
    const int w = (int)(idx / ((long long)L * M));
    const int m = V.m_list[ms];
    const size_t wl = (size_t)w * L + p;
    const double od = V.lay_od[wl], ssa = V.lay_ssa[wl], secant = V.lay_secant[wl];
    const double trans_top = V.lay_trans[(size_t)w * (L + 1) + p];
    double beta[NSTR];
#pragma unroll
    for (int l = 0; l < NSTR; ++l) beta[l] = V.lay_beta[wl * NSTR + l];

    LayerSol<N> S;
    layer_solve<N>(V.T, m, od, ssa, beta, secant, trans_top, S);
    if (S.status) raise_status(V.status, (unsigned)S.status);
    const bool thermal = V.emission && m == 0;   // thermal sources only enter order 0 (sktran_do_rte.cpp:1337-1340)
    const double b0 = thermal ? V.lay_thermal[wl * 2] : 0.0, b1 = thermal ? V.lay_thermal[wl * 2 + 1] : 0.0;
    if (thermal) thermal_particular<N>(V.T, od, ssa, b0, b1, S);

    double* Wp = V.Wp + (size_t)idx * N * N;
    double* Wm = V.Wm + (size_t)idx * N * N;
    for (int i = 0; i < N * N; ++i) {
        Wp[i] = S.Wp[i];
        Wm[i] = S.Wm[i];
    }
    double* kth = V.kth + (size_t)idx * 2 * N;
    double* G = V.G + (size_t)idx * 4 * N;
    for (int j = 0; j < N; ++j) {
        kth[j] = S.k[j];
        kth[N + j] = S.theta[j];
        G[j] = S.Gpt[j];
        G[N + j] = S.Gmt[j];
        G[2 * N + j] = S.Gpb[j];
        G[3 * N + j] = S.Gmb[j];
    }
    // Lambertian surface couples only m = 0 (sktran_do_surface.h:53-60, sktran_do_rte.h:116-152)
    const bool ground = (p == L - 1) && (m == 0);
    double sp[N], sm[N], sG = 0.0;
    if (ground) {
        for (int j = 0; j < N; ++j) {
            double a = 0.0, b = 0.0;
            for (int i = 0; i < N; ++i) {
                a += V.T.wt[i] * V.T.mu[i] * S.Wp[i * N + j];
                b += V.T.wt[i] * V.T.mu[i] * S.Wm[i * N + j];
            }
            sp[j] = a;
            sm[j] = b;
        }
        for (int i = 0; i < N; ++i) sG += V.T.wt[i] * V.T.mu[i] * S.Gpb[i];
        double* surf = V.surf + (size_t)w * (2 * N + 1);
        for (int j = 0; j < N; ++j) {
            surf[j] = sp[j];
            surf[N + j] = sm[j];
        }
        surf[2 * N] = sG;
    }
    const double cum_top = V.lay_cumod[(size_t)w * (L + 1) + p];
    const double cum_all = V.lay_cumod[(size_t)w * (L + 1) + L];
    const double albedo = V.albedo[w];
    const double trans_floor = V.lay_trans[(size_t)w * (L + 1) + L];
    for (int los = 0; los < nlos; ++los) {
        double cpos[N], cneg[N], v;
        los_layer_terms<N>(V.T, m, los, od, ssa, beta, secant, trans_top, V.include_ss != 0, S, cpos, cneg, v, thermal, b0, b1);
        const double mu = V.T.los_mu[los];
        const double att = exp(-cum_top / mu);
        const size_t o = (((size_t)w * M + ms) * nlos + los) * L + p;
        double* wv = V.wvec + o * 2 * N;
        double vv = v * att;
        if (ground) {
            // ground-leaving radiance toward the LOS (sktran_do_layerarray.cpp:5-288) attenuated by the
            // whole column: linear in (L, M) of the bottom layer plus constants
            const double attg = exp(-cum_all / mu) * albedo;
            for (int j = 0; j < N; ++j) {
                wv[j] = cpos[j] * att + attg * 2.0 * sp[j] * S.theta[j];
                wv[N + j] = cneg[j] * att + attg * 2.0 * sm[j];
            }
            double direct = V.include_ss ? V.T.csz / kPi * trans_floor : 0.0;
            vv += attg * (direct + 2.0 * sG);
            // surface emission leaves the ground unreflected, inside the reference's direct-bounce branch
            // (sktran_do_layerarray.cpp:225-266)
            if (V.semis && V.include_ss) vv += exp(-cum_all / mu) * V.semis[w];
        } else {
            for (int j = 0; j < N; ++j) {
                wv[j] = cpos[j] * att;
                wv[N + j] = cneg[j] * att;
            }
        }
        V.vsrc[o] = vv;
    }
}

// Kernel-based (non-Lambertian) surface, device view of the tables of disco_brdf.h
struct BrdfView {
    int nk, nargs;
    const double *Rss, *rsun, *Rls, *rlsun;
    const double* args;   // [nargs][nw] of the chunk (k + nargs * w)
    // snow model: per-wavelength coefficients pw[w][M][npairs] written by k_brdf_expand_snow from the sample tables
    const double* pw;
    double* pw_out;
    int npairs, nsamples;
    const double *snow_r0, *snow_g, *snow_cos, *snow_w, *snow_scale;
};

// K2s body: one (wavelength, azimuth slot, row): row t < N is stream t, row N + los a line of sight.  R = sum_k args_k
// R^k_m(row, .) from the host tables (or the per-wavelength coefficients of the snow model), then the products with the
// bottom layer's solution that the BVP ground rows need (stream rows) / the ground-leaving radiance toward the LOS
// (sktran_do_rte.h:116-345, sktran_do_layerarray.cpp:5-288), and the rows themselves for the weighting functions.
DISCO_HD void surface_general_body(const ChunkView& V, const BrdfView& B, int w, int ms, int t) {
    const int N = V.T.N, M = V.M, L = V.T.L, nlos = V.T.nlos, nstr = V.T.nstr;
    const int m = V.m_list[ms];
    if (t >= N + nlos) return;
    const size_t idxb = ((size_t)w * M + ms) * L + (L - 1);
    const double* Wp = V.Wp + idxb * N * N;
    const double* Wm = V.Wm + idxb * N * N;
    const double* th = V.kth + idxb * 2 * N + N;
    const double* Gpb = V.G + idxb * 4 * N + 2 * N;
    const bool stream = t < N;
    const int los = t - N;
    double R[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) R[q] = 0.0;
    double rsun = 0.0;
    if (B.pw) {   // per-wavelength coefficients (snow model): pw[w][ms][pair]
        const double* pw = B.pw + ((size_t)w * M + ms) * B.npairs;
        const double* tab = stream ? pw + t * N : pw + N * N + N + los * N;
        rsun = stream ? pw[N * N + t] : pw[N * N + N + nlos * N + los];
#pragma unroll
        for (int q = 0; q < 16; ++q)
            if (q < N) R[q] = tab[q];
    }
    for (int k = 0; k < (B.pw ? 0 : B.nk); ++k) {
        const double a = B.args[k + (size_t)B.nargs * w];
        const double* tab = stream ? B.Rss + (((size_t)k * nstr + m) * N + t) * N
                                                : B.Rls + (((size_t)k * nstr + m) * nlos + los) * N;
        rsun += a * (stream ? B.rsun[((size_t)k * nstr + m) * N + t] : B.rlsun[((size_t)k * nstr + m) * nlos + los]);
#pragma unroll
        for (int q = 0; q < 16; ++q)
            if (q < N) R[q] = fma(a, tab[q], R[q]);
    }
    double sg = 0.0;
#pragma unroll
    for (int q = 0; q < 16; ++q)
        if (q < N) sg = fma(R[q], Gpb[q], sg);
    const double t_floor = V.lay_trans[(size_t)w * (L + 1) + L];
    if (stream) {
        double* gs = V.gsurf_out + ((size_t)w * M + ms) * V.gsurf_stride;
        for (int j = 0; j < N; ++j) {
            double sp = 0.0, sm = 0.0;
#pragma unroll
            for (int q = 0; q < 16; ++q)
                if (q < N) {
                    sp = fma(R[q], Wp[q * N + j], sp);
                    sm = fma(R[q], Wm[q * N + j], sm);
                }
            gs[t * N + j] = sp;
            gs[N * N + t * N + j] = sm;
        }
        gs[2 * N * N + t] = sg;
        gs[2 * N * N + N + t] = rsun;
    }
    if (V.gsurf_rows) {   // reflection rows for the weighting-function kernels
        double* row = V.gsurf_rows + (((size_t)w * M + ms) * (N + nlos) + t) * (N + 1);
#pragma unroll
        for (int q = 0; q < 16; ++q)
            if (q < N) row[q] = R[q];
        row[N] = rsun;
    }
    if (!stream) {
        const size_t o = (((size_t)w * M + ms) * nlos + los) * L + (L - 1);
        const double attg = exp(-V.lay_cumod[(size_t)w * (L + 1) + L] / V.T.los_mu[los]);
        double* wv = V.wvec + o * 2 * N;
        for (int j = 0; j < N; ++j) {
            double lp = 0.0, lm = 0.0;
#pragma unroll
            for (int q = 0; q < 16; ++q)
                if (q < N) {
                    lp = fma(R[q], Wp[q * N + j], lp);
                    lm = fma(R[q], Wm[q * N + j], lm);
                }
            wv[j] += attg * lp * th[j];
            wv[N + j] += attg * lm;
        }
        const double direct = V.include_ss ? V.T.csz / kPi * t_floor * rsun : 0.0;
        V.vsrc[o * V.vsrc_w] += attg * (sg + direct);
    }
}

// K4 body: one (wavelength, LOS)
DISCO_HD void radiance_body(const ChunkView& V, long long idx) {
    const int L = V.T.L, M = V.M, nlos = V.T.nlos, N = V.T.N, nstr = V.T.nstr;
    const int w = (int)(idx / nlos), los = (int)(idx % nlos);
    double total = 0.0;
    for (int ms = 0; ms < M; ++ms) {
        const int m = V.m_list[ms];
        const size_t o = (((size_t)w * M + ms) * nlos + los) * L;
        const double* wv = V.wvec + o * 2 * N;
        const double* vs = V.vsrc + o * V.vsrc_w;
        const double* x = V.xsol + ((size_t)w * M + ms) * L * 2 * N;
        double comp = 0.0;
        // same accumulation order as the reference's upward recursion: ground/bottom layer first
        for (int p = L - 1; p >= 0; --p) {
            double s = 0.0;
            for (int c = 0; c < V.vsrc_w; ++c) s += vs[(size_t)p * V.vsrc_w + c];
            for (int j = 0; j < 2 * N; ++j) s += wv[(size_t)p * 2 * N + j] * x[(size_t)p * 2 * N + j];
            comp += s;
        }
        total += comp * V.T.los_cosmphi[(size_t)los * nstr + m];
    }
    V.radiance[idx] = total;
}

}  // namespace disco
