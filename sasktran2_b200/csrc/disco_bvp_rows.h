// Row loaders of the staircase systems (forward A x = b and adjoint A^T z = wvec), host/device so that the
// CPU emulation in tests/host_emul.cpp exercises exactly the rows the CUDA kernels eliminate.
#pragma once
#include "disco_bodies.h"

namespace disco {

DISCO_HD void prefetch_line(const void* p) {
#if defined(__CUDA_ARCH__)
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#else
    (void)p;
#endif
}

// Row loader of the forward system A x = b (sktran_do_rte.cpp:1898-2294, sktran_do_rte.h:116-345)
template <int N>
struct ForwardRows {
    static constexpr int NRHS = 1;
    static constexpr bool KEEPS_L = true;  // the forward elimination can record its multipliers (transposed solves)
    static constexpr bool ROW_OWNER_BACKSUB = false;
    const ChunkView& V;
    int w, ms, m, L;
    const double *Wp, *Wm, *kth, *G;
    // index of new row `rank` of `step` in the row order of A (N TOA rows, 2N per interface, N ground rows)
    DISCO_HD int row_of(int step, int rank) const { return (step == 0 ? 0 : N + 2 * N * step) + rank; }
    DISCO_HD ForwardRows(const ChunkView& V_, int w_, int ms_) : V(V_), w(w_), ms(ms_) {
        L = V.T.L;
        m = V.m_list[ms];
        const size_t lay0 = ((size_t)w * V.M + ms) * L;
        Wp = V.Wp + lay0 * N * N;
        Wm = V.Wm + lay0 * N * N;
        kth = V.kth + lay0 * 2 * N;
        G = V.G + lay0 * 4 * N;
    }
    DISCO_HD int nsteps() const { return L; }
    DISCO_HD int nleft(int) const { return 2 * N; }
    DISCO_HD int nright(int step) const { return step < L - 1 ? 2 * N : 0; }
    DISCO_HD int nnew(int step) const { return (step == 0 ? N : 0) + (step < L - 1 ? 2 * N : N); }
    DISCO_HD void load(int step, int rank, double* a) const {
        const int p = step;
        if (step == 0 && rank < N) {
            // TOA rows: W+_0 L + W-_0 Theta_0 M = -G+top_0
            const int i = rank;
#pragma unroll
            for (int j = 0; j < N; ++j) {
                a[j] = Wp[i * N + j];
                a[N + j] = Wm[i * N + j] * kth[N + j];
                a[2 * N + j] = 0.0;
                a[3 * N + j] = 0.0;
            }
            a[4 * N] = -G[i];
            return;
        }
        if (step == 0) rank -= N;
        const double* Wpu = Wp + (size_t)p * N * N;
        const double* Wmu = Wm + (size_t)p * N * N;
        const double* thu = kth + (size_t)p * 2 * N + N;
        const double* Gu = G + (size_t)p * 4 * N;
        if (p < L - 1) {
            // continuity between layer p (upper) and p+1 (lower)
            const double* Wpl = Wpu + N * N;
            const double* Wml = Wmu + N * N;
            const double* thl = thu + 2 * N;
            const double* Gl = Gu + 4 * N;
            const bool first = rank < N;  // rows i: W- family; rows i+N: W+ family
            const int i = first ? rank : rank - N;
            const double* A1 = first ? Wmu : Wpu;  // multiplies L_upper (with theta)
            const double* A2 = first ? Wpu : Wmu;  // multiplies M_upper
            const double* B1 = first ? Wml : Wpl;  // multiplies L_lower
            const double* B2 = first ? Wpl : Wml;  // multiplies M_lower (with theta)
#pragma unroll
            for (int j = 0; j < N; ++j) {
                a[j] = A1[i * N + j] * thu[j];
                a[N + j] = A2[i * N + j];
                a[2 * N + j] = -B1[i * N + j];
                a[3 * N + j] = -(B2[i * N + j] * thl[j]);
            }
            a[4 * N] = first ? (-Gu[3 * N + i] + Gl[N + i]) : (-Gu[2 * N + i] + Gl[i]);
        } else {
            // ground rows (Lambertian: only m = 0 reflects)
            const int i = rank;
            const bool refl = (m == 0);
            const double alb2 = refl ? 2.0 * V.albedo[w] : 0.0;
            const double* surf = V.surf + (size_t)w * (2 * N + 1);
            // kernel-based BRDF: every order reflects, with per-stream sums prepared by k_surface_general
            const double* gs = V.gsurf ? V.gsurf + ((size_t)w * V.M + ms) * V.gsurf_stride : nullptr;
#pragma unroll
            for (int j = 0; j < N; ++j) {
                double vm = Wmu[i * N + j], vp = Wpu[i * N + j];
                if (refl) {
                    vm -= alb2 * surf[j];      // - (1+d_m0) rho sum_q w mu W+_qj
                    vp -= alb2 * surf[N + j];  // - (1+d_m0) rho sum_q w mu W-_qj
                }
                if (gs) {
                    vm -= gs[i * N + j];
                    vp -= gs[N * N + i * N + j];
                }
                a[j] = vm * thu[j];
                a[N + j] = vp;
                a[2 * N + j] = 0.0;
                a[3 * N + j] = 0.0;
            }
            double rhs = -Gu[3 * N + i];
            if (refl) {
                rhs += alb2 * surf[2 * N];
                rhs += V.T.csz * V.albedo[w] / kPi * V.lay_trans[(size_t)w * (L + 1) + L];
            }
            if (gs) rhs += gs[2 * N * N + i] + V.T.csz * gs[2 * N * N + N + i] / kPi * V.lay_trans[(size_t)w * (L + 1) + L];
            if (refl && V.semis) rhs += V.semis[w];   // surface emission (ground_direct_sun, sktran_do_rte.h:229-235)
            a[4 * N] = rhs;
        }
    }
    // ---- the same rows from shared-memory tiles of the layers' solutions (brought in by the TMA engine, disco_bvp.cuh):
    //      tile of one layer = W+ rows | W- rows (row stride TILE_RS doubles: conflict-free LDS.128 for 8 lanes) |
    //      k, theta [2N] | G [4N]
    static constexpr int TILE_RS = N + 2;
    static constexpr int TILE_DOUBLES = 2 * N * TILE_RS + 6 * N;
    DISCO_HD void load_tiles(int step, int rank, double* a, const double* tu, const double* tl) const {
        const int p = step;
        const double* Wpu = tu;
        const double* Wmu = tu + N * TILE_RS;
        const double* thu = tu + 2 * N * TILE_RS + N;
        const double* Gu = tu + 2 * N * TILE_RS + 2 * N;
        if (step == 0 && rank < N) {
            const int i = rank;
#pragma unroll
            for (int j = 0; j < N; ++j) {
                a[j] = Wpu[i * TILE_RS + j];
                a[N + j] = Wmu[i * TILE_RS + j] * thu[j];
                a[2 * N + j] = 0.0;
                a[3 * N + j] = 0.0;
            }
            a[4 * N] = -Gu[i];
            return;
        }
        if (step == 0) rank -= N;
        if (p < L - 1) {
            const double* Wpl = tl;
            const double* Wml = tl + N * TILE_RS;
            const double* thl = tl + 2 * N * TILE_RS + N;
            const double* Gl = tl + 2 * N * TILE_RS + 2 * N;
            const bool first = rank < N;
            const int i = first ? rank : rank - N;
            const double* A1 = first ? Wmu : Wpu;
            const double* A2 = first ? Wpu : Wmu;
            const double* B1 = first ? Wml : Wpl;
            const double* B2 = first ? Wpl : Wml;
#pragma unroll
            for (int j = 0; j < N; ++j) {
                a[j] = A1[i * TILE_RS + j] * thu[j];
                a[N + j] = A2[i * TILE_RS + j];
                a[2 * N + j] = -B1[i * TILE_RS + j];
                a[3 * N + j] = -(B2[i * TILE_RS + j] * thl[j]);
            }
            a[4 * N] = first ? (-Gu[3 * N + i] + Gl[N + i]) : (-Gu[2 * N + i] + Gl[i]);
        } else {
            const int i = rank;
            const bool refl = (m == 0);
            const double alb2 = refl ? 2.0 * V.albedo[w] : 0.0;
            const double* surf = V.surf + (size_t)w * (2 * N + 1);
            const double* gs = V.gsurf ? V.gsurf + ((size_t)w * V.M + ms) * V.gsurf_stride : nullptr;
#pragma unroll
            for (int j = 0; j < N; ++j) {
                double vm = Wmu[i * TILE_RS + j], vp = Wpu[i * TILE_RS + j];
                if (refl) {
                    vm -= alb2 * surf[j];
                    vp -= alb2 * surf[N + j];
                }
                if (gs) {
                    vm -= gs[i * N + j];
                    vp -= gs[N * N + i * N + j];
                }
                a[j] = vm * thu[j];
                a[N + j] = vp;
                a[2 * N + j] = 0.0;
                a[3 * N + j] = 0.0;
            }
            double rhs = -Gu[3 * N + i];
            if (refl) {
                rhs += alb2 * surf[2 * N];
                rhs += V.T.csz * V.albedo[w] / kPi * V.lay_trans[(size_t)w * (L + 1) + L];
            }
            if (gs) rhs += gs[2 * N * N + i] + V.T.csz * gs[2 * N * N + N + i] / kPi * V.lay_trans[(size_t)w * (L + 1) + L];
            if (refl && V.semis) rhs += V.semis[w];   // surface emission (ground_direct_sun, sktran_do_rte.h:229-235)
            a[4 * N] = rhs;
        }
    }
    // global sources of a layer's tile (device only)
    DISCO_HD const double* tile_src_wp(int p) const { return Wp + (size_t)p * N * N; }
    DISCO_HD const double* tile_src_wm(int p) const { return Wm + (size_t)p * N * N; }
    DISCO_HD const double* tile_src_kth(int p) const { return kth + (size_t)p * 2 * N; }
    DISCO_HD const double* tile_src_g(int p) const { return G + (size_t)p * 4 * N; }

    // Segment loaders of the 2D-distributed elimination: window block wb (0..3) of new row `rank` of `step`,
    // i.e. entries a[wb * N .. wb * N + N) of load(), and its right-hand side.
    DISCO_HD void load_seg(int step, int rank, int wb, double* seg) const {
        const int p = step;
#pragma unroll
        for (int j = 0; j < N; ++j) seg[j] = 0.0;
        if (step == 0 && rank < N) {  // TOA rows
            const int i = rank;
            if (wb == 0) {
#pragma unroll
                for (int j = 0; j < N; ++j) seg[j] = Wp[i * N + j];
            } else if (wb == 1) {
#pragma unroll
                for (int j = 0; j < N; ++j) seg[j] = Wm[i * N + j] * kth[N + j];
            }
            return;
        }
        if (step == 0) rank -= N;
        const double* Wpu = Wp + (size_t)p * N * N;
        const double* Wmu = Wm + (size_t)p * N * N;
        const double* thu = kth + (size_t)p * 2 * N + N;
        if (p < L - 1) {
            // interior interface rows, branch-free: the matrix is W+ when (wb odd) == (first half), the layer is
            // p + (wb >> 1), blocks 0 and 3 carry the layer's theta, blocks 2 and 3 a minus sign
            const bool first = rank < N;
            const int i = first ? rank : rank - N;
            {
                const int lay = p + (wb >> 1);
                const bool plus = ((wb & 1) != 0) == first;
                const double* Mx = (plus ? Wp : Wm) + (size_t)lay * N * N + i * N;
                const double* th = kth + (size_t)lay * 2 * N + N;
                const bool scaled = (wb == 0) || (wb == 3);
                const double sgn = (wb >= 2) ? -1.0 : 1.0;
#pragma unroll
                for (int j = 0; j < N; ++j) seg[j] = sgn * Mx[j] * (scaled ? th[j] : 1.0);
                return;
            }
            if (wb == 0) {
                const double* A1 = first ? Wmu : Wpu;
#pragma unroll
                for (int j = 0; j < N; ++j) seg[j] = A1[i * N + j] * thu[j];
            } else if (wb == 1) {
                const double* A2 = first ? Wpu : Wmu;
#pragma unroll
                for (int j = 0; j < N; ++j) seg[j] = A2[i * N + j];
            } else if (wb == 2) {
                const double* B1 = first ? Wmu + N * N : Wpu + N * N;
#pragma unroll
                for (int j = 0; j < N; ++j) seg[j] = -B1[i * N + j];
            } else {
                const double* B2 = first ? Wpu + N * N : Wmu + N * N;
                const double* thl = thu + 2 * N;
#pragma unroll
                for (int j = 0; j < N; ++j) seg[j] = -(B2[i * N + j] * thl[j]);
            }
        } else if (wb < 2) {  // ground rows
            const int i = rank;
            const bool refl = (m == 0);
            const double alb2 = refl ? 2.0 * V.albedo[w] : 0.0;
            const double* surf = V.surf + (size_t)w * (2 * N + 1);
            const double* gs = V.gsurf ? V.gsurf + ((size_t)w * V.M + ms) * V.gsurf_stride : nullptr;
#pragma unroll
            for (int j = 0; j < N; ++j) {
                if (wb == 0) {
                    double vm = Wmu[i * N + j];
                    if (refl) vm -= alb2 * surf[j];
                    if (gs) vm -= gs[i * N + j];
                    seg[j] = vm * thu[j];
                } else {
                    double vp = Wpu[i * N + j];
                    if (refl) vp -= alb2 * surf[N + j];
                    if (gs) vp -= gs[N * N + i * N + j];
                    seg[j] = vp;
                }
            }
        }
    }
    DISCO_HD double load_rhs(int step, int rank, int) const {
        const int p = step;
        if (step == 0 && rank < N) return -G[rank];
        if (step == 0) rank -= N;
        const double* Gu = G + (size_t)p * 4 * N;
        if (p < L - 1) {
            const double* Gl = Gu + 4 * N;
            const bool first = rank < N;
            const int i = first ? rank : rank - N;
            return first ? (-Gu[3 * N + i] + Gl[N + i]) : (-Gu[2 * N + i] + Gl[i]);
        }
        const int i = rank;
        double rhs = -Gu[3 * N + i];
        if (m == 0) {
            const double* surf = V.surf + (size_t)w * (2 * N + 1);
            rhs += 2.0 * V.albedo[w] * surf[2 * N];
            rhs += V.T.csz * V.albedo[w] / kPi * V.lay_trans[(size_t)w * (L + 1) + L];
        }
        if (V.gsurf) {
            const double* gs = V.gsurf + ((size_t)w * V.M + ms) * V.gsurf_stride;
            rhs += gs[2 * N * N + i] + V.T.csz * gs[2 * N * N + N + i] / kPi * V.lay_trans[(size_t)w * (L + 1) + L];
        }
        if (m == 0 && V.semis) rhs += V.semis[w];
        return rhs;
    }
    // lines of layer step+1 (the lower layer of interface step+1), one 128-byte line per lane
    DISCO_HD void prefetch(int step, int lane) const {
        const int p = step + 1;
        if (p >= L) return;
        constexpr int WL = (N * N * 8 + 127) / 128, KL = (2 * N * 8 + 127) / 128, GLn = (4 * N * 8 + 127) / 128;
        const char* ptr = nullptr;
        if (lane < WL) ptr = (const char*)(Wp + (size_t)p * N * N) + 128 * lane;
        else if (lane < 2 * WL) ptr = (const char*)(Wm + (size_t)p * N * N) + 128 * (lane - WL);
        else if (lane < 2 * WL + KL) ptr = (const char*)(kth + (size_t)p * 2 * N) + 128 * (lane - 2 * WL);
        else if (lane < 2 * WL + KL + GLn) ptr = (const char*)(G + (size_t)p * 4 * N) + 128 * (lane - 2 * WL - KL);
        if (ptr) prefetch_line(ptr);
    }
    // unknown c of block `step`, right-hand side r
    DISCO_HD void store(int step, int c, int, double v) const {
        V.xsol[(((size_t)w * V.M + ms) * L + step) * 2 * N + c] = v;
    }
};

// Forward system of several solar zenith angles at once (spherical path): the matrix - homogeneous solutions, surface
// coupling - does not depend on the SZA, only the right-hand side does (Green's particular solutions G, direct beam on
// the ground).  Right-hand side r comes from slice r of the per-SZA arrays (ChunkView::sza_*); one factorisation, NRHS
// substitutions.  The reference factorises once per SZA (DOSource::calculate, source_term/do_source.cpp:35-58).
template <int N, int NRHS_>
struct ForwardRowsMulti : ForwardRows<N> {
    using Base = ForwardRows<N>;
    static constexpr int NRHS = NRHS_;
    static constexpr bool KEEPS_L = false;
    static constexpr bool ROW_OWNER_BACKSUB = true;   // few right-hand sides: lane c owns pivot row c for each of them
    DISCO_HD ForwardRowsMulti(const ChunkView& V_, int w_, int ms_) : Base(V_, w_, ms_) {}
    DISCO_HD double rhs_of(int step, int rank, int r) const {
        const ChunkView& V = this->V;
        const int L = this->L, w = this->w;
        const double* G = this->G + (size_t)r * V.sza_G;
        const int p = step;
        if (step == 0 && rank < N) return -G[rank];
        if (step == 0) rank -= N;
        const double* Gu = G + (size_t)p * 4 * N;
        if (p < L - 1) {
            const double* Gl = Gu + 4 * N;
            const bool first = rank < N;
            const int i = first ? rank : rank - N;
            return first ? (-Gu[3 * N + i] + Gl[N + i]) : (-Gu[2 * N + i] + Gl[i]);
        }
        double rhs = -Gu[3 * N + rank];
        if (this->m == 0) {
            const double* surf = V.surf + (size_t)r * V.sza_surf + (size_t)w * (2 * N + 1);
            rhs += 2.0 * V.albedo[w] * surf[2 * N];
            rhs += V.sza_csz[r] * V.albedo[w] / kPi * V.lay_trans[(size_t)r * V.sza_trans + (size_t)w * (L + 1) + L];
        }
        return rhs;
    }
    DISCO_HD void load(int step, int rank, double* a) const {
        Base::load(step, rank, a);   // matrix part and the right-hand side of slice 0
#pragma unroll
        for (int r = 1; r < NRHS; ++r) a[4 * N + r] = rhs_of(step, rank, r);
    }
    DISCO_HD void store(int step, int c, int r, double v) const {
        const ChunkView& V = this->V;
        V.xsol[(size_t)r * V.sza_xsol + (((size_t)this->w * V.M + this->ms) * this->L + step) * 2 * N + c] = v;
    }
};

// Row loader of the transposed system A^T z = wvec(los): equation block b = columns of layer b, unknown
// blocks = rows of A (TOA rows, interface rows, ground rows).
template <int N, int NRHS_>
struct AdjointRows {
    static constexpr int NRHS = NRHS_;
    static constexpr bool KEEPS_L = false;
    static constexpr bool ROW_OWNER_BACKSUB = false;
    DISCO_HD int row_of(int, int) const { return 0; }
    const ChunkView& V;
    int w, ms, m, L, los0, nl;
    const double *Wp, *Wm, *kth;
    DISCO_HD AdjointRows(const ChunkView& V_, int w_, int ms_, int los0_) : V(V_), w(w_), ms(ms_), los0(los0_) {
        L = V.T.L;
        m = V.m_list[ms];
        nl = V.T.nlos - los0 < NRHS ? V.T.nlos - los0 : NRHS;
        const size_t lay0 = ((size_t)w * V.M + ms) * L;
        Wp = V.Wp + lay0 * N * N;
        Wm = V.Wm + lay0 * N * N;
        kth = V.kth + lay0 * 2 * N;
    }
    DISCO_HD int nsteps() const { return L + 1; }
    DISCO_HD int nleft(int step) const { return (step == 0 || step == L) ? N : 2 * N; }
    DISCO_HD int nright(int step) const { return step < L - 1 ? 2 * N : (step == L - 1 ? N : 0); }
    DISCO_HD int nnew(int step) const { return step < L ? 2 * N : 0; }
    DISCO_HD void load(int step, int rank, double* a) const {
        const int b = step;  // layer whose unknown column `rank` this equation belongs to
        const bool isL = rank < N;
        const int j = isL ? rank : rank - N;
        const double* Wpb = Wp + (size_t)b * N * N;
        const double* Wmb = Wm + (size_t)b * N * N;
        const double th = kth[(size_t)b * 2 * N + N + j];
#pragma unroll
        for (int c = 0; c < 4 * N; ++c) a[c] = 0.0;
        if (b == 0) {
            // column of the TOA block [W+ | W- Theta]
#pragma unroll
            for (int i = 0; i < N; ++i) a[i] = isL ? Wpb[i * N + j] : Wmb[i * N + j] * th;
        } else {
            // column of -V_b (layer b is the lower layer of interface b)
#pragma unroll
            for (int i = 0; i < N; ++i) {
                a[i] = isL ? -Wmb[i * N + j] : -(Wpb[i * N + j] * th);
                a[N + i] = isL ? -Wpb[i * N + j] : -(Wmb[i * N + j] * th);
            }
        }
        if (b < L - 1) {
            // column of U_{b+1} (layer b is the upper layer of interface b+1)
#pragma unroll
            for (int i = 0; i < N; ++i) {
                a[2 * N + i] = isL ? Wmb[i * N + j] * th : Wpb[i * N + j];
                a[3 * N + i] = isL ? Wpb[i * N + j] * th : Wmb[i * N + j];
            }
        } else {
            // column of the ground block [v- Theta | v+]
            const bool refl = (m == 0);
            const double alb2 = refl ? 2.0 * V.albedo[w] : 0.0;
            const double* surf = V.surf + (size_t)w * (2 * N + 1);
            const double* gs = V.gsurf ? V.gsurf + ((size_t)w * V.M + ms) * V.gsurf_stride : nullptr;
#pragma unroll
            for (int i = 0; i < N; ++i) {
                double vm = Wmb[i * N + j], vp = Wpb[i * N + j];
                if (refl) {
                    vm -= alb2 * surf[j];
                    vp -= alb2 * surf[N + j];
                }
                if (gs) {   // kernel-based BRDF: every order reflects (rows of k_surface_general)
                    vm -= gs[i * N + j];
                    vp -= gs[N * N + i * N + j];
                }
                a[2 * N + i] = isL ? vm * th : vp;
            }
        }
#pragma unroll
        for (int r = 0; r < NRHS; ++r) {
            const int los = los0 + (r < nl ? r : 0);
            const size_t o = (((size_t)w * V.M + ms) * V.T.nlos + los) * L + b;
            a[4 * N + r] = (r < nl) ? V.wvec[o * 2 * N + rank] : 0.0;
        }
    }
    // Segment loaders of the 2D-distributed elimination (entries a[wb * N .. wb * N + N) of load())
    DISCO_HD void load_seg(int step, int rank, int wb, double* seg) const {
        const int b = step;
        const bool isL = rank < N;
        const int j = isL ? rank : rank - N;
        const double* Wpb = Wp + (size_t)b * N * N;
        const double* Wmb = Wm + (size_t)b * N * N;
        const double th = kth[(size_t)b * 2 * N + N + j];
#pragma unroll
        for (int i = 0; i < N; ++i) seg[i] = 0.0;
        if (b > 0 && b < L - 1) {
            // interior layers, branch-free: W+ when (wb odd) == isL, theta_j when (wb >= 2) == isL, minus for wb < 2
            const bool plus = ((wb & 1) != 0) == isL;
            const double* Mx = (plus ? Wpb : Wmb) + j;
            const double sc = (((wb >> 1) != 0) == isL) ? th : 1.0;
            const double f = (wb < 2) ? -sc : sc;
#pragma unroll
            for (int i = 0; i < N; ++i) seg[i] = f * Mx[i * N];
            return;
        }
        if (wb == 0) {
            if (b == 0) {
#pragma unroll
                for (int i = 0; i < N; ++i) seg[i] = isL ? Wpb[i * N + j] : Wmb[i * N + j] * th;
            } else {
#pragma unroll
                for (int i = 0; i < N; ++i) seg[i] = isL ? -Wmb[i * N + j] : -(Wpb[i * N + j] * th);
            }
        } else if (wb == 1) {
            if (b > 0) {
#pragma unroll
                for (int i = 0; i < N; ++i) seg[i] = isL ? -Wpb[i * N + j] : -(Wmb[i * N + j] * th);
            }
        } else if (b < L - 1) {
            if (wb == 2) {
#pragma unroll
                for (int i = 0; i < N; ++i) seg[i] = isL ? Wmb[i * N + j] * th : Wpb[i * N + j];
            } else {
#pragma unroll
                for (int i = 0; i < N; ++i) seg[i] = isL ? Wpb[i * N + j] * th : Wmb[i * N + j];
            }
        } else if (wb == 2) {  // ground block [v- Theta | v+]
            const bool refl = (m == 0);
            const double alb2 = refl ? 2.0 * V.albedo[w] : 0.0;
            const double* surf = V.surf + (size_t)w * (2 * N + 1);
            const double* gs = V.gsurf ? V.gsurf + ((size_t)w * V.M + ms) * V.gsurf_stride : nullptr;
#pragma unroll
            for (int i = 0; i < N; ++i) {
                double vm = Wmb[i * N + j], vp = Wpb[i * N + j];
                if (refl) {
                    vm -= alb2 * surf[j];
                    vp -= alb2 * surf[N + j];
                }
                if (gs) {
                    vm -= gs[i * N + j];
                    vp -= gs[N * N + i * N + j];
                }
                seg[i] = isL ? vm * th : vp;
            }
        }
    }
    DISCO_HD double load_rhs(int step, int rank, int r) const {
        if (r >= nl) return 0.0;
        const size_t o = (((size_t)w * V.M + ms) * V.T.nlos + (los0 + r)) * L + step;
        return V.wvec[o * 2 * N + rank];
    }
    // lines of layer `step`: W+-, k|theta and the wvec rows of the batch's lines of sight
    DISCO_HD void prefetch(int step, int lane) const {
        const int b = step;
        if (b >= L) return;
        constexpr int WL = (N * N * 8 + 127) / 128, KL = (2 * N * 8 + 127) / 128, VL = (2 * N * 8 + 127) / 128;
        const char* ptr = nullptr;
        if (lane < WL) ptr = (const char*)(Wp + (size_t)b * N * N) + 128 * lane;
        else if (lane < 2 * WL) ptr = (const char*)(Wm + (size_t)b * N * N) + 128 * (lane - WL);
        else if (lane < 2 * WL + KL) ptr = (const char*)(kth + (size_t)b * 2 * N) + 128 * (lane - 2 * WL);
        else if (lane < 2 * WL + KL + nl * VL) {
            const int q = lane - 2 * WL - KL;
            const int los = los0 + q / VL;
            const size_t o = (((size_t)w * V.M + ms) * V.T.nlos + los) * L + b;
            ptr = (const char*)(V.wvec + o * 2 * N) + 128 * (q % VL);
        }
        if (ptr) prefetch_line(ptr);
    }
    DISCO_HD void store(int step, int c, int r, double v) const {
        if (r >= nl) return;
        const int row = (step == 0) ? c : N + (step - 1) * 2 * N + c;
        V.zadj[(((size_t)w * V.M + ms) * ((size_t)2 * N * L) + row) * V.T.nlos + (los0 + r)] = v;
    }
};

}  // namespace disco
