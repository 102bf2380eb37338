// Staircase LU (forward and adjoint boundary-value solves) — templates, instantiated per stream count in
// disco_bvp_inst.cu (one translation unit per N so that the build parallelises).
#pragma once
#include <cstdlib>
#include <type_traits>

#include "disco_bvp_rows.h"
#include "disco_kernels.cuh"

namespace disco {

#define FULL_MASK 0xffffffffu

// -------------------------------------------------------------------------------------------------
// K3: boundary value problem per (w, m), forward (A x = b) and adjoint (A^T z = w_los).
//
// Unknowns x = [L_0 M_0 | L_1 M_1 | ...] (2N per layer).  Rows: N TOA rows, 2N continuity rows per interface,
// N ground rows.  Eliminating the 2N unknowns of layer p only ever involves the N rows left over from the
// layers above plus the 2N rows of interface p+1, i.e. a 3N x (4N + nrhs) panel ("staircase").  One lane owns
// one panel row in registers; the pivot row is broadcast through shared memory.  The candidate rows of every
// column are exactly the rows LAPACK's banded partial pivoting (kl = ku = 3N-1) would search, so the
// factorisation is the reference's dgbsv in a different storage scheme.  The transposed system has the same
// staircase shape with the roles of layers and interfaces exchanged (unknown blocks N, 2N, ..., 2N, N), so
// the adjoint solve (the reference's dgbtrs('T') in RTESolver::backprop, sktran_do_rte.cpp:1793-1836) is the
// same elimination with a different row loader and one right-hand side per line of sight.
// -------------------------------------------------------------------------------------------------
template <int N, int NRHS>
struct BvpCfg {
    static constexpr int NC = 2 * N;
    static constexpr int ROWS = 3 * N;
    static constexpr int GL = ROWS <= 4 ? 4 : (ROWS <= 8 ? 8 : (ROWS <= 16 ? 16 : 32));
    static constexpr int R = (ROWS + GL - 1) / GL;
    static constexpr int ROWLEN = 4 * N + NRHS;
    static constexpr int GROUPS_PER_WARP = 32 / GL;
    static constexpr int WARPS_PER_BLOCK = (N >= 16) ? 2 : 4;
    static constexpr int GROUPS_PER_BLOCK = GROUPS_PER_WARP * WARPS_PER_BLOCK;
    static constexpr int BUF = ROWLEN + 1;  // padded
    static constexpr int SMEM_DOUBLES_PER_GROUP = 2 * BUF + NC * ROWLEN + NRHS * NC;
};

template <int N, class Prob>
__device__ __forceinline__ void staircase_solve(const Prob& prob, double* gs, double* fac, int lane, unsigned gbase,
                                                unsigned gmask, bool valid, unsigned int* status) {
    constexpr int NRHS = Prob::NRHS;
    using C = BvpCfg<N, NRHS>;
    constexpr int NC = C::NC, GL = C::GL, R = C::R, ROWLEN = C::ROWLEN;
    double* buf = gs;                      // [2][BUF]
    double* facs = gs + 2 * C::BUF;        // [NC][ROWLEN]
    double* xs = facs + NC * ROWLEN;       // [NRHS][NC]

    double a[R][ROWLEN];
    bool act[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        act[r] = false;
#pragma unroll
        for (int c = 0; c < ROWLEN; ++c) a[r][c] = 0.0;
    }
    const unsigned lt_mask = (lane == 0) ? 0u : (((1u << lane) - 1u) << gbase);
    bool singular = false;
    const int nsteps = prob.nsteps();

    for (int step = 0; step < nsteps; ++step) {
        // ---- bring the new rows of this step into free slots
        {
            unsigned freeb[R];
            bool wasfree[R];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                wasfree[r] = !act[r];
                freeb[r] = __ballot_sync(FULL_MASK, wasfree[r]) & gmask;
            }
            const int needed = prob.nnew(step);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if (wasfree[r]) {
                    // rank of this free slot in slot-id order (slot id = lane * R + r)
                    int rank = 0;
#pragma unroll
                    for (int r2 = 0; r2 < R; ++r2) rank += __popc(freeb[r2] & lt_mask);
#pragma unroll
                    for (int r2 = 0; r2 < R; ++r2)
                        if (r2 < r && wasfree[r2]) rank += 1;
                    if (rank < needed) {
                        act[r] = true;
                        prob.load(step, rank, a[r]);
                    }
                }
            }
        }
        // ---- eliminate the unknowns of this block
        const int nleft = prob.nleft(step);
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            if (c < nleft) {
                double best = -1.0;
                int bsid = 0x7fffffff;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    if (act[r]) {
                        const double v = fabs(a[r][c]);
                        if (v > best) {
                            best = v;
                            bsid = lane * R + r;
                        }
                    }
                }
#pragma unroll
                for (int off = GL / 2; off > 0; off >>= 1) {
                    const double ov = __shfl_xor_sync(FULL_MASK, best, off);
                    const int oi = __shfl_xor_sync(FULL_MASK, bsid, off);
                    if (ov > best || (ov == best && oi < bsid)) {
                        best = ov;
                        bsid = oi;
                    }
                }
                if (!(best > 0.0)) singular = true;
                double* bc = buf + (c & 1) * C::BUF;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    if (act[r] && bsid == lane * R + r) {
                        act[r] = false;
#pragma unroll
                        for (int cc = 0; cc < ROWLEN; ++cc) {
                            const double v = (cc >= c) ? a[r][cc] : 0.0;
                            if (cc >= c) bc[cc] = v;
                            facs[c * ROWLEN + cc] = v;
                        }
                    }
                }
                __syncwarp();
                const double pinv = 1.0 / bc[c];
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    if (act[r]) {
                        const double f = a[r][c] * pinv;
#pragma unroll
                        for (int cc = c + 1; cc < ROWLEN; ++cc) a[r][cc] -= f * bc[cc];
                        a[r][c] = 0.0;
                    }
                }
            }
        }
        __syncwarp();
        // ---- flush the pivot rows of this block (coalesced) and slide the panel window
        if (valid) {
            double* dst = fac + (size_t)step * NC * ROWLEN;
            for (int e = lane; e < nleft * ROWLEN; e += GL) dst[e] = facs[e];
        }
        __syncwarp();
        if (step < nsteps - 1) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
#pragma unroll
                for (int j = 0; j < NC; ++j) {
                    a[r][j] = a[r][NC + j];
                    a[r][NC + j] = 0.0;
                }
            }
        }
    }
    if (singular && valid) atomicOr(status, 4u);

    // ---- back substitution, last block first; lane c owns pivot row c
    for (int step = nsteps - 1; step >= 0; --step) {
        const int nleft = prob.nleft(step);
        const int nright = prob.nright(step);
        if (step < nsteps - 1) {
            const double* src = fac + (size_t)step * NC * ROWLEN;
            for (int e = lane; e < nleft * ROWLEN; e += GL) facs[e] = src[e];
            __syncwarp();
        }
        double acc[NRHS], myx[NRHS];
#pragma unroll
        for (int r = 0; r < NRHS; ++r) {
            acc[r] = 0.0;
            myx[r] = 0.0;
        }
        if (lane < nleft) {
#pragma unroll
            for (int r = 0; r < NRHS; ++r) acc[r] = facs[lane * ROWLEN + 4 * N + r];
            for (int j = 0; j < nright; ++j) {
                const double rj = facs[lane * ROWLEN + NC + j];
#pragma unroll
                for (int r = 0; r < NRHS; ++r) acc[r] -= rj * xs[r * NC + j];
            }
        }
#pragma unroll
        for (int cc = NC - 1; cc >= 0; --cc) {
            if (cc < nleft) {
                const double dinv = 1.0 / facs[cc * ROWLEN + cc];
                const double u = (lane < cc) ? facs[lane * ROWLEN + cc] : 0.0;
#pragma unroll
                for (int r = 0; r < NRHS; ++r) {
                    double xv = (lane == cc) ? acc[r] * dinv : 0.0;
                    xv = __shfl_sync(FULL_MASK, xv, (int)gbase + cc);
                    if (lane < cc) acc[r] -= u * xv;
                    if (lane == cc) myx[r] = xv;
                }
            }
        }
        __syncwarp();
        if (lane < nleft) {
#pragma unroll
            for (int r = 0; r < NRHS; ++r) {
                xs[r * NC + lane] = myx[r];
                if (valid) prob.store(step, lane, r, myx[r]);
            }
        }
        __syncwarp();
    }
}

// -------------------------------------------------------------------------------------------------
// Version 2 of the staircase solve for 3N <= lanes per group (N = 1, 2, 4, 8), one panel row per lane:
//   * pivot search in 3 instructions: REDUX.MAX over the high words of |a_c| (non-negative doubles order like
//     their bit patterns), ballot, find-first-set.  The pivot is the largest candidate up to the 2^-20 resolution
//     of the high word - threshold partial pivoting with threshold 1 - 1e-6, the same candidate set as LAPACK's
//     dgbtf2 (kl = ku = 3N - 1);
//   * the pivot row goes to shared memory once: the per-column rows of the step's factor block double as the
//     broadcast buffer (uniform LDS.128 for the rank-1 update, 2 DFMA per load) and are flushed to HBM coalesced at
//     the end of the step together with 1/pivot, so the back substitution does not divide;
//   * back substitution: one right-hand side - lane c owns pivot row c, x broadcast by shuffle; several right-hand
//     sides (adjoint, one per line of sight) - lane r owns right-hand side r and runs the whole block solve from
//     uniform shared-memory loads, no shuffles (a 64-bit shuffle costs 4 DFMA issue slots on B200).
// Factor layout in HBM: fac[step][c][FS], FS = even(4N + NRHS) + 2, entry FS - 2 = 1 / pivot; entries left of the
// diagonal are unspecified.
// -------------------------------------------------------------------------------------------------
template <int N, int NRHS>
struct BvpCfg2 {
    static constexpr int NC = 2 * N;
    static constexpr int ROWS = 3 * N;
    // panel rows per lane.  R = 2 (two problems per warp for N = 8, every broadcast operand feeding two DFMAs) was
    // measured SLOWER on B200 (bvp 9.2 -> 15.3 ms per 1000 wavelengths): registers halve the resident warps while the
    // per-pivot broadcast stays bound by shared-memory instruction throughput (tools/microbench/lu_patterns.cu).
    static constexpr int R = 1;
    static constexpr int LANES_ROWS = (ROWS + R - 1) / R > NC ? (ROWS + R - 1) / R : NC;
    static constexpr int LANES_NEEDED = (NRHS > 1 && NRHS > LANES_ROWS) ? NRHS : LANES_ROWS;  // one RHS per lane in the multi-RHS solve
    static constexpr int GL = LANES_NEEDED <= 2 ? 2 : (LANES_NEEDED <= 4 ? 4 : (LANES_NEEDED <= 8 ? 8 : (LANES_NEEDED <= 16 ? 16 : 32)));
    static constexpr int ROWLEN = 4 * N + NRHS;
    static constexpr int RL2 = (ROWLEN + 1) & ~1;
    static constexpr int FS = RL2 + 2;                    // row stride: padded row | 1/pivot | pad
    static constexpr int LS = (ROWS + 2) & ~1;            // multiplier row stride: one per panel lane | (pivot lane, row) | pad
    static constexpr int GROUPS_PER_WARP = 32 / GL;
    static constexpr int WARPS_PER_BLOCK = 4;
    static constexpr int GROUPS_PER_BLOCK = GROUPS_PER_WARP * WARPS_PER_BLOCK;
    static constexpr int MIN_BLOCKS = (R == 1) ? 4 : (NRHS == 1 ? 3 : 2);  // register budget: 128 / 168 / 255 per thread
    // factor blocks resident during the back substitution: the one being solved + 2 (1 for the long multi-RHS steps) in flight
    static constexpr int STAGES = (NRHS == 1) ? 3 : 2;
    static constexpr int SMEM_DOUBLES_PER_GROUP = STAGES * NC * FS + NC * (NRHS <= 4 ? NRHS : 1) + 2;   // factor block ring | x of the block below (row-owner substitution) | two mbarriers (TMA row staging)
    static_assert(ROWS <= GL * R && NC <= GL, "panel rows fit the group; one pivot row per lane in the back substitution");
    static_assert(NRHS == 1 || NRHS <= GL, "one right-hand side per lane");
    static_assert(NRHS > 1 || (R == 1 && ROWS < GL), "a spare lane records the pivot's (lane, row) next to the multipliers");
};

// 1/x without the library's slow-path call on the critical path of every pivot: MUFU seed, two Newton steps and
// a residual correction (0 or denormal pivots give inf/garbage; those systems are flagged singular anyway)
__device__ __forceinline__ double rcp_pivot(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    return fma(y, e, y);
}

template <int N, class Prob>
__device__ __forceinline__ void staircase_solve_v2(const Prob& prob, double* gs, double* fac, int lane, unsigned gbase,
                                                   unsigned gmask, bool valid, unsigned int* status, double* lf = nullptr) {
    constexpr int NRHS = Prob::NRHS;
    using C = BvpCfg2<N, NRHS>;
    constexpr int NC = C::NC, GL = C::GL, ROWLEN = C::ROWLEN, RL2 = C::RL2, FS = C::FS;
    constexpr int CPL = (NC * FS + GL - 1) / GL;  // factor-block elements per lane
    constexpr int STAGES = C::STAGES;
    double* ring = gs;                     // [STAGES][NC][FS]; block of step s lives in slot s % STAGES
    double* xs = gs + STAGES * NC * FS;    // [NC]
    const unsigned lane_bit = 1u << (gbase + lane);
    const unsigned lt_mask = (lane_bit - 1u) & gmask;

    constexpr int R = C::R;
    double a[R][ROWLEN];
    bool act[R];
    int myrow = 0;  // row of A held by this lane (kept for the multiplier record, R = 1)
#pragma unroll
    for (int r = 0; r < R; ++r) {
        act[r] = false;
#pragma unroll
        for (int c = 0; c < ROWLEN; ++c) a[r][c] = 0.0;
    }
    bool singular = false;
    const int nsteps = prob.nsteps();
    constexpr int LS = C::LS;

    for (int step = 0; step < nsteps; ++step) {
        {   // new rows of this step go to the lowest free slots (slot id = lane * R + r)
            unsigned freeb[R];
            bool wasfree[R];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                wasfree[r] = !act[r];
                freeb[r] = __ballot_sync(gmask, wasfree[r]) & gmask;
            }
            const int needed = prob.nnew(step);
            int below = 0;  // free slots in lower lanes
#pragma unroll
            for (int r = 0; r < R; ++r) below += __popc(freeb[r] & lt_mask);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if (wasfree[r]) {
                    int rank = below;
#pragma unroll
                    for (int r2 = 0; r2 < r; ++r2) rank += wasfree[r2] ? 1 : 0;
                    if (rank < needed) {
                        act[r] = true;
                        prob.load(step, rank, a[r]);
                        if (Prob::KEEPS_L) myrow = prob.row_of(step, rank);
                    }
                }
            }
        }
        if (step + 1 < nsteps) prob.prefetch(step + 1, lane);  // next step's rows -> L1 while this block is eliminated
        const int nleft = prob.nleft(step);
        double* facs = ring + (step % STAGES) * NC * FS;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            if (c < nleft) {
                unsigned key = 0u;
                int rbest = 0;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const unsigned kr = act[r] ? (unsigned)__double2hiint(fabs(a[r][c])) : 0u;
                    if (act[r] && (kr > key || (r == 0))) {
                        if (kr > key || key == 0u) {
                            key = kr;
                            rbest = r;
                        }
                    }
                }
                bool any_act = false;
#pragma unroll
                for (int r = 0; r < R; ++r) any_act |= act[r];
                const unsigned mx = __reduce_max_sync(gmask, key);
                const unsigned cand = __ballot_sync(gmask, any_act && key == mx);
                if (mx == 0u) singular = true;
                double* bc = facs + c * FS;
                if (any_act && key == mx && (cand & lt_mask) == 0u) {  // lowest candidate lane holds the pivot row
                    const int c0 = c & ~1;
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        if (r == rbest) {
                            act[r] = false;
#pragma unroll
                            for (int cc = c0; cc < RL2; cc += 2) {
                                const double v0 = (cc < ROWLEN) ? a[r][cc < ROWLEN ? cc : 0] : 0.0;
                                const double v1 = (cc + 1 < ROWLEN) ? a[r][cc + 1 < ROWLEN ? cc + 1 : 0] : 0.0;
                                *reinterpret_cast<double2*>(bc + cc) = make_double2(v0, v1);
                            }
                            bc[RL2] = rcp_pivot(a[r][c]);
                        }
                    }
                }
                __syncwarp(gmask);
                if (Prob::KEEPS_L && lf != nullptr) {
                    // record of this pivot for transposed solves with the same factors (k_bvp_tsolve): lane i < 3N
                    // writes its multiplier (0 when it holds no candidate row), the spare lane 3N the pivot's lane and row
                    const int plane = __ffs(cand) - 1;
                    const int prow = __shfl_sync(gmask, myrow, plane);
                    const double fi = act[0] ? a[0][c] * bc[RL2] : 0.0;
                    const double rec = (lane < C::ROWS) ? fi : __hiloint2double(prow, plane - (int)gbase);
                    if (valid && lane <= C::ROWS) lf[((size_t)step * NC + c) * LS + lane] = rec;
                }
                {
                    bool any_left = false;
#pragma unroll
                    for (int r = 0; r < R; ++r) any_left |= act[r];
                    if (any_left) {
                        double f[R];
                        const double pinv = bc[RL2];
#pragma unroll
                        for (int r = 0; r < R; ++r) {
                            f[r] = act[r] ? a[r][c] * pinv : 0.0;
                            a[r][c] = act[r] ? 0.0 : a[r][c];
                        }
#pragma unroll
                        for (int cc = c + 1; cc < ROWLEN; ++cc) {
                            const double b = bc[cc];
#pragma unroll
                            for (int r = 0; r < R; ++r) a[r][cc] = fma(-f[r], b, a[r][cc]);
                        }
                    }
                }
            }
        }
        __syncwarp(gmask);
        {
            double tmp[CPL];
#pragma unroll
            for (int i = 0; i < CPL; ++i) tmp[i] = facs[(lane + i * GL) < NC * FS ? lane + i * GL : 0];
            if (valid) {
                double* dst = fac + (size_t)step * NC * FS;
#pragma unroll
                for (int i = 0; i < CPL; ++i)
                    if (lane + i * GL < nleft * FS) dst[lane + i * GL] = tmp[i];
            }
        }
        __syncwarp(gmask);
        if (step < nsteps - 1) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
#pragma unroll
                for (int j = 0; j < NC; ++j) {
                    a[r][j] = a[r][NC + j];
                    a[r][NC + j] = 0.0;
                }
            }
        }
    }
    if (singular && valid) atomicOr(status, 4u);

    // factor blocks come back from HBM through a 3-slot cp.async ring: while block s is solved, blocks s-1 and
    // s-2 are in flight (16 bytes per lane per copy, no registers)
    auto fetch_block = [&](int step) {
        if (step >= 0) {
            const double* src = fac + (size_t)step * NC * FS;
            double* dst = ring + (step % STAGES) * NC * FS;
#pragma unroll
            for (int i = 0; i < (NC * FS / 2 + GL - 1) / GL; ++i) {
                const int e = 2 * (lane + i * GL);
                if (e < NC * FS) {
                    const unsigned sa = (unsigned)__cvta_generic_to_shared(dst + e);
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(src + e) : "memory");
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    __syncwarp(gmask);
#pragma unroll
    for (int k = 2; k <= STAGES; ++k) fetch_block(nsteps - k);

    if constexpr (NRHS > 1 && Prob::ROW_OWNER_BACKSUB) {
        // ---- a few right-hand sides (one per solar zenith angle): lane c owns pivot row c for every one of them
        for (int step = nsteps - 1; step >= 0; --step) {
            const int nleft = prob.nleft(step);
            const int nright = prob.nright(step);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
            __syncwarp(gmask);
            const double* facs = ring + (step % STAGES) * NC * FS;
            const int row = lane < nleft ? lane : 0;
            const double* my = facs + row * FS;
            double acc[NRHS], myx[NRHS];
#pragma unroll
            for (int r = 0; r < NRHS; ++r) {
                acc[r] = my[4 * N + r];
                myx[r] = 0.0;
            }
            for (int jx = 0; jx < nright; ++jx) {
                const double u = my[NC + jx];
#pragma unroll
                for (int r = 0; r < NRHS; ++r) acc[r] = fma(-u, xs[r * NC + jx], acc[r]);
            }
            const double pinv = my[RL2];
#pragma unroll
            for (int cc = NC - 1; cc >= 0; --cc) {
                if (cc < nleft) {
                    const double u = my[cc];
#pragma unroll
                    for (int r = 0; r < NRHS; ++r) {
                        double xv = acc[r] * pinv;
                        xv = __shfl_sync(gmask, xv, (int)gbase + cc);
                        if (lane < cc) acc[r] = fma(-u, xv, acc[r]);
                        if (lane == cc) myx[r] = xv;
                    }
                }
            }
            __syncwarp(gmask);
            if (lane < nleft) {
#pragma unroll
                for (int r = 0; r < NRHS; ++r) {
                    xs[r * NC + lane] = myx[r];
                    if (valid) prob.store(step, lane, r, myx[r]);
                }
            }
            __syncwarp(gmask);
            fetch_block(step - STAGES);
        }
    } else if (NRHS == 1) {
        // ---- lane c owns pivot row c of the block; x of the block below sits in xs
        for (int step = nsteps - 1; step >= 0; --step) {
            const int nleft = prob.nleft(step);
            const int nright = prob.nright(step);
            asm volatile("cp.async.wait_group 1;" ::: "memory");  // all but the newest group: block `step` has landed
            __syncwarp(gmask);
            const double* facs = ring + (step % STAGES) * NC * FS;
            const int row = lane < nleft ? lane : 0;
            const double* my = facs + row * FS;
            double acc = my[4 * N];
            for (int jx = 0; jx < nright; ++jx) acc = fma(-my[NC + jx], xs[jx], acc);
            const double pinv = my[RL2];
            double myx = 0.0;
#pragma unroll
            for (int cc = NC - 1; cc >= 0; --cc) {
                if (cc < nleft) {
                    double xv = acc * pinv;
                    xv = __shfl_sync(gmask, xv, (int)gbase + cc);
                    if (lane < cc) acc = fma(-my[cc], xv, acc);
                    if (lane == cc) myx = xv;
                }
            }
            __syncwarp(gmask);
            if (lane < nleft) {
                xs[lane] = myx;
                if (valid) prob.store(step, lane, 0, myx);
            }
            __syncwarp(gmask);
            fetch_block(step - STAGES);  // into the slot this step just released
        }
    } else {
        // ---- lane r owns right-hand side r: the whole block solve from uniform shared-memory loads
        const int r = lane < NRHS ? lane : 0;
        double xn[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) xn[c] = 0.0;
        for (int step = nsteps - 1; step >= 0; --step) {
            const int nleft = prob.nleft(step);
            const int nright = prob.nright(step);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
            __syncwarp(gmask);
            const double* facs = ring + (step % STAGES) * NC * FS;
            double acc[NC];
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                acc[c] = 0.0;
                if (c < nleft) {
                    const double* u = facs + c * FS;
                    double s = u[4 * N + r];
#pragma unroll
                    for (int jx = 0; jx < NC; ++jx)
                        if (jx < nright) s = fma(-u[NC + jx], xn[jx], s);
                    acc[c] = s;
                }
            }
#pragma unroll
            for (int cc = NC - 1; cc >= 0; --cc) {
                if (cc < nleft) {
                    const double xv = acc[cc] * facs[cc * FS + RL2];
                    acc[cc] = xv;
#pragma unroll
                    for (int c = 0; c < cc; ++c) acc[c] = fma(-facs[c * FS + cc], xv, acc[c]);
                }
            }
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                xn[c] = acc[c];
                if (valid && lane < NRHS && c < nleft) prob.store(step, c, r, acc[c]);
            }
            __syncwarp(gmask);
            fetch_block(step - STAGES);
        }
    }
}

// -------------------------------------------------------------------------------------------------
// Version 4 of the forward staircase solve (one right-hand side, one panel row per lane): the elimination of
// version 2 in blocks of B = 4 pivot columns, so that the shared-memory broadcast - the bound of version 2, 14
// STS.128 from ONE lane per pivot at ~4 MIO cycles each (profiles/README.md) - is paid once per block by the
// block's B pivot lanes together:
//   panel phase   per pivot: REDUX pivot search as in version 2; the pivot lane's B panel entries and 1/pivot travel
//                 by shuffle (<= B 64-bit shuffles), every candidate row forms its multiplier f_k and updates its
//                 own panel entries - pivots and multipliers are exactly those of the column-by-column elimination;
//   block phase   the B pivot lanes publish their RAW trailing entries together (one STS.128 stream for the block),
//                 every lane applies the rank-B update a -= sum_j M_j raw_j with the composite multipliers
//                 M = f L11^-1 (L11 = the block's unit lower triangle, exchanged by 6 shuffles): mathematically the
//                 B rank-1 updates, ~B^2 extra DFMA per lane;
//   factor rows   the update also runs on the pivot lanes, whose rows thereby become the U rows; they overwrite
//                 their raw rows in the factor block, which is flushed as in version 2 - layout, back substitution
//                 and the transposed solves (k_bvp_tsolve) are unchanged.
// -------------------------------------------------------------------------------------------------
// ---- TMA helpers (1-D bulk copies global -> shared, completion on an mbarrier of the warp's shared-memory area)
__device__ __forceinline__ void bvp_mbar_init(unsigned long long* bar, unsigned count) {
    const unsigned mb = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mb), "r"(count) : "memory");
}
__device__ __forceinline__ void bvp_mbar_expect(unsigned long long* bar, unsigned bytes) {
    const unsigned mb = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bvp_tma_load(void* smem_dst, const void* gmem_src, unsigned bytes, unsigned long long* bar) {
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    const unsigned mb = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(gmem_src), "r"(bytes), "r"(mb)
                 : "memory");
}
__device__ __forceinline__ void bvp_mbar_wait(unsigned long long* bar, unsigned phase) {
    const unsigned mb = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(mb),
        "r"(phase)
        : "memory");
}

template <int N, class Prob, bool TMA = false>
__device__ __forceinline__ void staircase_solve_v4(const Prob& prob, double* gs, double* fac, int lane, unsigned gbase,
                                                   unsigned gmask, bool valid, unsigned int* status, double* lf) {
    static_assert(Prob::NRHS == 1, "forward solve");
    using C = BvpCfg2<N, 1>;
    constexpr int NC = C::NC, GL = C::GL, ROWLEN = C::ROWLEN, RL2 = C::RL2, FS = C::FS, LS = C::LS;
    constexpr int B = NC >= 4 ? 4 : 2;
    constexpr int CPL = (NC * FS + GL - 1) / GL;
    constexpr int STAGES = C::STAGES;
    double* ring = gs;
    double* xs = gs + STAGES * NC * FS;
    const unsigned lane_bit = 1u << (gbase + lane);
    const unsigned lt_mask = (lane_bit - 1u) & gmask;

    double a[ROWLEN];
    bool act = false;
    int myrow = 0;
#pragma unroll
    for (int c = 0; c < ROWLEN; ++c) a[c] = 0.0;
    bool singular = false;
    const int nsteps = prob.nsteps();

    // TMA variant: the solutions of layer p (W+-, k | theta, G) arrive as one tile per layer in ring slots 1 and 2 (free
    // during the elimination: the factor block of every step but the last is assembled in slot 0), one step ahead of
    // their use, by per-row bulk copies issued from the lanes of the warp - off the LSU data pipe that bounds this kernel.
    constexpr int TRS = Prob::TILE_RS, TD = Prob::TILE_DOUBLES;
    static_assert(!TMA || TD <= NC * FS, "a layer tile fits a ring slot");
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(xs + NC);   // [2], after x of the block below
    auto tile_of = [&](int layer) { return ring + (1 + (layer & 1)) * NC * FS; };
    auto issue_tile = [&](int layer) {
        if (layer >= nsteps) return;
        double* t = tile_of(layer);
        unsigned long long* bar = bars + (layer & 1);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if (lane == 0) bvp_mbar_expect(bar, (unsigned)((2 * N * N + 6 * N) * sizeof(double)));
        __syncwarp(gmask);
        if (lane < N) bvp_tma_load(t + lane * TRS, prob.tile_src_wp(layer) + lane * N, N * sizeof(double), bar);
        else if (lane < 2 * N) bvp_tma_load(t + N * TRS + (lane - N) * TRS, prob.tile_src_wm(layer) + (lane - N) * N, N * sizeof(double), bar);
        else if (lane == 2 * N) bvp_tma_load(t + 2 * N * TRS, prob.tile_src_kth(layer), 2 * N * sizeof(double), bar);
        else if (lane == 2 * N + 1) bvp_tma_load(t + 2 * N * TRS + 2 * N, prob.tile_src_g(layer), 4 * N * sizeof(double), bar);
    };
    if (TMA) {
        if (lane == 0) {
            bvp_mbar_init(bars, 1);
            bvp_mbar_init(bars + 1, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp(gmask);
        issue_tile(0);
        issue_tile(1);
    }

    for (int step = 0; step < nsteps; ++step) {
        if (TMA) {
            if (step == 0) bvp_mbar_wait(bars, 0);
            if (step + 1 < nsteps) bvp_mbar_wait(bars + ((step + 1) & 1), (unsigned)(((step + 1) >> 1) & 1));
        }
        {   // new rows of this step go to the lowest free lanes
            const unsigned freeb = __ballot_sync(gmask, !act) & gmask;
            const int rank = __popc(freeb & lt_mask);
            if (!act && rank < prob.nnew(step)) {
                act = true;
                if (TMA)
                    prob.load_tiles(step, rank, a, tile_of(step), tile_of(step + 1));
                else
                    prob.load(step, rank, a);
                myrow = prob.row_of(step, rank);
            }
        }
        if (TMA) {
            __syncwarp(gmask);          // every lane has read the tile of layer `step`: its slot takes layer step + 2
            issue_tile(step + 2);
        } else if (step + 1 < nsteps) {
            prob.prefetch(step + 1, lane);
        }
        // TMA: slot 0 until the last step, whose block must sit where the back substitution expects it (its slot is a
        // tile slot, no longer needed: the last step's rows are loaded)
        double* facs = ring + ((TMA && step < nsteps - 1) ? 0 : (step % STAGES)) * NC * FS;
#pragma unroll
        for (int cb = 0; cb < NC / B; ++cb) {
            const int c0 = cb * B;
            double fk[B];
            int mypiv = -1;
            double mypinv = 0.0;
            int planes[B];
            // ---- panel phase
#pragma unroll
            for (int k = 0; k < B; ++k) {
                const int c = c0 + k;
                const unsigned key = act ? (unsigned)__double2hiint(fabs(a[c])) : 0u;
                const unsigned mx = __reduce_max_sync(gmask, key);
                const unsigned cand = __ballot_sync(gmask, act && key == mx);
                if (mx == 0u) singular = true;
                const int plane = (cand != 0u) ? __ffs(cand) - 1 : (int)gbase;  // warp lane of the pivot row
                planes[k] = plane;
                const bool ispiv = act && key == mx && (cand & lt_mask) == 0u;
                double pinv = 0.0;
                if (ispiv) {
                    act = false;
                    mypiv = k;
                    pinv = rcp_pivot(a[c]);
                    mypinv = pinv;
                }
                pinv = __shfl_sync(gmask, pinv, plane);
                const double f = act ? a[c] * pinv : 0.0;
                if (lf != nullptr) {
                    const int prow = __shfl_sync(gmask, myrow, plane);
                    const double rec = (lane < C::ROWS) ? f : __hiloint2double(prow, plane - (int)gbase);
                    if (valid && lane <= C::ROWS) lf[((size_t)step * NC + c) * LS + lane] = rec;
                }
                if (act) a[c] = 0.0;
                fk[k] = f;
#pragma unroll
                for (int t = k + 1; t < B; ++t) {
                    const double u = __shfl_sync(gmask, a[c0 + t], plane);
                    a[c0 + t] = fma(-f, u, a[c0 + t]);
                }
            }
            // ---- composite multipliers M = f T, T = L11^-1 (L11[k][j] = multiplier of pivot lane k against pivot j)
            double Tm[B][B], Mj[B];
#pragma unroll
            for (int k = 1; k < B; ++k)
#pragma unroll
                for (int j = 0; j < k; ++j) Tm[k][j] = __shfl_sync(gmask, fk[j], planes[k]);  // L11[k][j] for now
#pragma unroll
            for (int j = 0; j < B; ++j) {  // column j of T by forward substitution (in place, rows ascending)
                double col[B];
#pragma unroll
                for (int i = j + 1; i < B; ++i) {
                    double t = -Tm[i][j];
#pragma unroll
                    for (int k = j + 1; k < i; ++k) t = fma(-Tm[i][k], col[k], t);
                    col[i] = t;
                }
                double m = fk[j];
#pragma unroll
                for (int k = j + 1; k < B; ++k) m = fma(fk[k], col[k], m);
                Mj[j] = m;
            }
            // ---- the block's pivot lanes publish their raw trailing entries
            double* myrowp = facs + (c0 + (mypiv >= 0 ? mypiv : 0)) * FS;
            if (mypiv >= 0) {
#pragma unroll
                for (int cc = c0 + B; cc < RL2; cc += 2) {
                    const double v0 = (cc < ROWLEN) ? a[cc < ROWLEN ? cc : 0] : 0.0;
                    const double v1 = (cc + 1 < ROWLEN) ? a[cc + 1 < ROWLEN ? cc + 1 : 0] : 0.0;
                    *reinterpret_cast<double2*>(myrowp + cc) = make_double2(v0, v1);
                }
            }
            __syncwarp(gmask);
            // ---- rank-B update of the trailing columns (pivot lanes included: their rows become rows of U)
#pragma unroll
            for (int cc = c0 + B; cc < ROWLEN; ++cc) {
                double sacc = a[cc];
#pragma unroll
                for (int j = 0; j < B; ++j) sacc = fma(-Mj[j], facs[(c0 + j) * FS + cc], sacc);
                a[cc] = sacc;
            }
            __syncwarp(gmask);
            if (mypiv >= 0) {  // factor row: panel entries | U trailing entries | 1 / pivot
#pragma unroll
                for (int cc = c0; cc < RL2; cc += 2) {
                    const double v0 = (cc < ROWLEN) ? a[cc < ROWLEN ? cc : 0] : 0.0;
                    const double v1 = (cc + 1 < ROWLEN) ? a[cc + 1 < ROWLEN ? cc + 1 : 0] : 0.0;
                    *reinterpret_cast<double2*>(myrowp + cc) = make_double2(v0, v1);
                }
                myrowp[RL2] = mypinv;
            }
        }
        __syncwarp(gmask);
        {
            double tmp[CPL];
#pragma unroll
            for (int i = 0; i < CPL; ++i) tmp[i] = facs[(lane + i * GL) < NC * FS ? lane + i * GL : 0];
            if (valid) {
                double* dst = fac + (size_t)step * NC * FS;
#pragma unroll
                for (int i = 0; i < CPL; ++i)
                    if (lane + i * GL < NC * FS) dst[lane + i * GL] = tmp[i];
            }
        }
        __syncwarp(gmask);
        if (step < nsteps - 1) {
#pragma unroll
            for (int j = 0; j < NC; ++j) {
                a[j] = a[NC + j];
                a[NC + j] = 0.0;
            }
        }
    }
    if (singular && valid) atomicOr(status, 4u);

    // ---- back substitution: identical to version 2 (lane c owns pivot row c of the block; x by shuffle)
    auto fetch_block = [&](int step) {
        if (step >= 0) {
            const double* src = fac + (size_t)step * NC * FS;
            double* dst = ring + (step % STAGES) * NC * FS;
#pragma unroll
            for (int i = 0; i < (NC * FS / 2 + GL - 1) / GL; ++i) {
                const int e = 2 * (lane + i * GL);
                if (e < NC * FS) {
                    const unsigned sa = (unsigned)__cvta_generic_to_shared(dst + e);
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(src + e) : "memory");
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    __syncwarp(gmask);
#pragma unroll
    for (int k = 2; k <= STAGES; ++k) fetch_block(nsteps - k);
    for (int step = nsteps - 1; step >= 0; --step) {
        const int nleft = prob.nleft(step);
        const int nright = prob.nright(step);
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        __syncwarp(gmask);
        const double* fb = ring + (step % STAGES) * NC * FS;
        const int row = lane < nleft ? lane : 0;
        const double* my = fb + row * FS;
        double acc = my[4 * N];
        for (int jx = 0; jx < nright; ++jx) acc = fma(-my[NC + jx], xs[jx], acc);
        const double pinv = my[RL2];
        double myx = 0.0;
#pragma unroll
        for (int cc = NC - 1; cc >= 0; --cc) {
            if (cc < nleft) {
                double xv = acc * pinv;
                xv = __shfl_sync(gmask, xv, (int)gbase + cc);
                if (lane < cc) acc = fma(-my[cc], xv, acc);
                if (lane == cc) myx = xv;
            }
        }
        __syncwarp(gmask);
        if (lane < nleft) {
            xs[lane] = myx;
            if (valid) prob.store(step, lane, 0, myx);
        }
        __syncwarp(gmask);
        fetch_block(step - STAGES);
    }
}

// -------------------------------------------------------------------------------------------------
// Version 3: the panel is distributed in two dimensions over one warp - 8 row lanes x 4 column groups, S = ceil(3N/8)
// row slots per lane, N panel columns (one window block) per group.  Column block gb of the band lives in group
// gb mod 4, so the window slides without moving data.  Per pivot the warp exchanges
//   1 pivot value + S multipliers + N row-segment entries (+ the right-hand-side entries)
// as 64-bit shuffles - 13 for N = 8 - instead of a 34-double row through shared memory (17 STS.128 from one lane at
// ~4 MIO cycles each + 17 LDS.128: tools/microbench/lu_patterns.cu, profiles/microbench_lu_r01.txt), all 32 lanes do
// the rank-1 update (S x N DFMA each), and the registers per thread drop from 128 to ~100.  Pivot rows stay in
// their registers until the end of the step and are then written to HBM in the layout of version 2
// (fac[step][c][FS]), so the back substitution is shared with version 2.
// -------------------------------------------------------------------------------------------------
template <int I, int E, class F>
__device__ __forceinline__ void static_for(F&& f) {
    if constexpr (I < E) {
        f(std::integral_constant<int, I>{});
        static_for<I + 1, E>(f);
    }
}

template <int N, int NRHS>
struct BvpCfg3 {
    static constexpr int NC = 2 * N;
    static constexpr int S = (3 * N + 7) / 8;          // row slots per lane
    static constexpr int RQ = (NRHS + 3) / 4;          // right-hand-side columns per group
    static constexpr int ROWLEN = 4 * N + NRHS;
    static constexpr int RL2 = (ROWLEN + 1) & ~1;
    static constexpr int FS = RL2 + 2;
    static constexpr int GL = 32;
    static constexpr int WARPS_PER_BLOCK = 4;
    static constexpr int GROUPS_PER_BLOCK = WARPS_PER_BLOCK;
    static constexpr int STAGES = 2;   // shared memory, not registers, would otherwise cap the resident warps
    static constexpr int SMEM_DOUBLES_PER_GROUP = STAGES * NC * FS + NC;
    static_assert(NC <= 32 && (NRHS == 1 || NRHS <= 32), "back substitution lanes");
};

// One pivot of the 2D elimination.  The four lanes that hold the pivot row (row lane ri_p, slot rp - both uniform
// over the warp) publish their N-column segments and right-hand-side entries in the broadcast buffer `bc`
// (4 x STS.128 warp-wide instead of 17 from one lane); everyone reads its own group's segment back (uniform per
// group), the multipliers travel from the owner group by shuffle.
//   bc layout: [4 groups][N] segments | [4 groups][RQ] right-hand sides
template <int N, int NRHS, int K>
__device__ __forceinline__ void pivot_update_2d(double (&a)[BvpCfg3<N, NRHS>::S][N], double (&rhs)[BvpCfg3<N, NRHS>::S][BvpCfg3<N, NRHS>::RQ],
                                                bool (&act)[BvpCfg3<N, NRHS>::S], int (&pidx)[BvpCfg3<N, NRHS>::S],
                                                double (&pinvs)[BvpCfg3<N, NRHS>::S], double* bc, int c, int ri, int g, int go,
                                                int ri_p, int rp) {
    using C = BvpCfg3<N, NRHS>;
    constexpr int S = C::S, RQ = C::RQ;
    const bool owner = (g == go);
    const bool mine = (ri == ri_p);
    double* seg_out = bc + g * N;
    double* rhs_out = bc + 4 * N + g * RQ;
    // rp is uniform over the warp: real branches (the barrier inside keeps the compiler from predicating the stores
    // of all S slots - a predicated-off STS.128 still costs its MIO cycles)
#pragma unroll
    for (int r = 0; r < S; ++r) {
        if (r == rp) {
            if (mine) {
                if (N % 2 == 0) {
#pragma unroll
                    for (int kk = 0; kk < N; kk += 2)
                        *reinterpret_cast<double2*>(seg_out + kk) = make_double2(a[r][kk], a[r][kk + 1 < N ? kk + 1 : kk]);
                } else {
#pragma unroll
                    for (int kk = 0; kk < N; ++kk) seg_out[kk] = a[r][kk];
                }
#pragma unroll
                for (int q = 0; q < RQ; ++q) rhs_out[q] = rhs[r][q];
                act[r] = false;
                pidx[r] = c;
            }
            __syncwarp();
        }
    }
    if (rp < 0 || rp >= S) __syncwarp();
    const double pinv = rcp_pivot(bc[go * N + K]);
    double f[S];
#pragma unroll
    for (int r = 0; r < S; ++r) {
        if (mine && r == rp) pinvs[r] = pinv;
        const double fr = (owner && act[r]) ? a[r][K] * pinv : 0.0;
        f[r] = __shfl_sync(FULL_MASK, fr, go * 8 + ri);
    }
    // unconditional updates: left of the pivot column the pivot row is already zero, at the pivot column the result
    // is overwritten below, inactive rows (and the pivot row itself) have f = 0
#pragma unroll
    for (int kk = 0; kk < N; ++kk) {
        const double sg = seg_out[kk];
#pragma unroll
        for (int r = 0; r < S; ++r) a[r][kk] = fma(-f[r], sg, a[r][kk]);
    }
    if (owner) {
#pragma unroll
        for (int r = 0; r < S; ++r)
            if (act[r]) a[r][K] = 0.0;
    }
#pragma unroll
    for (int q = 0; q < RQ; ++q) {
        const double sg = rhs_out[q];
#pragma unroll
        for (int r = 0; r < S; ++r) rhs[r][q] = fma(-f[r], sg, rhs[r][q]);
    }
}

template <int N, class Prob>
__device__ __forceinline__ void staircase_solve_2d(const Prob& prob, double* gs, double* fac, int lane, bool valid,
                                                   unsigned int* status) {
    constexpr int NRHS = Prob::NRHS;
    using C = BvpCfg3<N, NRHS>;
    constexpr int NC = C::NC, S = C::S, RQ = C::RQ, RL2 = C::RL2, FS = C::FS, STAGES = C::STAGES, GL = 32;
    const unsigned gmask = FULL_MASK;
    const unsigned gbase = 0;
    double* ring = gs;
    double* xs = gs + STAGES * NC * FS;
    const int g = lane >> 3, ri = lane & 7;

    double a[S][N], rhs[S][RQ], pinvs[S];
    bool act[S];
    int pidx[S];
#pragma unroll
    for (int r = 0; r < S; ++r) {
        act[r] = false;
        pidx[r] = -1;
        pinvs[r] = 0.0;
#pragma unroll
        for (int k = 0; k < N; ++k) a[r][k] = 0.0;
#pragma unroll
        for (int q = 0; q < RQ; ++q) rhs[r][q] = 0.0;
    }
    bool singular = false;
    const int nsteps = prob.nsteps();

    for (int step = 0; step < nsteps; ++step) {
        const int par2 = 2 * (step & 1);       // group of window block 0
        const int wb = (g - par2) & 3;          // window block held by this lane's group
        {   // new rows -> free (ri, slot) positions, slot-major order; every group keeps identical bookkeeping
            unsigned freeb[S];
#pragma unroll
            for (int r = 0; r < S; ++r) freeb[r] = __ballot_sync(FULL_MASK, !act[r]) & 0xffu;
            const int needed = prob.nnew(step);
            int before = 0;
#pragma unroll
            for (int r = 0; r < S; ++r) {
                const int rank = before + __popc(freeb[r] & ((1u << ri) - 1u));
                if (!act[r] && rank < needed) {
                    act[r] = true;
                    prob.load_seg(step, rank, wb, a[r]);
#pragma unroll
                    for (int q = 0; q < RQ; ++q) {
                        const int col = 4 * q + g;  // right-hand-side column held by (group g, local q)
                        rhs[r][q] = (col < NRHS) ? prob.load_rhs(step, rank, col) : 0.0;
                    }
                }
                before += __popc(freeb[r]);
            }
        }
        if (step + 1 < nsteps) prob.prefetch(step + 1, lane);
        const int nleft = prob.nleft(step);
        static_for<0, NC>([&](auto ic) {
            constexpr int c = decltype(ic)::value;
            if (c < nleft) {
                const int go = (par2 + c / N) & 3;
                const bool owner = (g == go);
                // candidate key: high word of |a| with the slot number in its two lowest bits (lower slots win
                // ties), so one REDUX.MAX + one ballot name the pivot row; resolution of the comparison 2^-18
                unsigned kmax = 0u;
#pragma unroll
                for (int r = 0; r < S; ++r) {
                    const unsigned kr = (owner && act[r]) ? (((unsigned)__double2hiint(fabs(a[r][c % N])) & ~3u) | (unsigned)(3 - r)) : 0u;
                    kmax = kr > kmax ? kr : kmax;
                }
                const unsigned mx = __reduce_max_sync(FULL_MASK, kmax);
                if ((mx & ~3u) == 0u) singular = true;
                const unsigned b = __ballot_sync(FULL_MASK, kmax == mx);
                const int rp = 3 - (int)(mx & 3u);
                const int ri_p = (__ffs(b) - 1) & 7;
                // rp is uniform over the warp: one specialised update per slot
                // two alternating broadcast buffers: one __syncwarp per pivot is enough
                pivot_update_2d<N, NRHS, c % N>(a, rhs, act, pidx, pinvs, ring + (c & 1) * (4 * N + 4 * RQ), c, ri, g, go,
                                                ri_p, rp);
            }
        });
        // pivot rows of this step -> HBM (version-2 layout), then release the two finished window blocks
#pragma unroll
        for (int r = 0; r < S; ++r) {
            if (pidx[r] >= 0) {
                if (valid) {
                    double* dst = fac + ((size_t)step * NC + pidx[r]) * FS;
                    if (N % 2 == 0) {
#pragma unroll
                        for (int k = 0; k < N; k += 2)
                            *reinterpret_cast<double2*>(dst + wb * N + k) = make_double2(a[r][k], a[r][k + 1 < N ? k + 1 : k]);
                    } else {
#pragma unroll
                        for (int k = 0; k < N; ++k) dst[wb * N + k] = a[r][k];
                    }
#pragma unroll
                    for (int q = 0; q < RQ; ++q)
                        if (4 * q + g < NRHS) dst[4 * N + 4 * q + g] = rhs[r][q];
                    if (g == 0) dst[RL2] = pinvs[r];
                }
                pidx[r] = -1;
            }
        }
        if (wb < 2) {
#pragma unroll
            for (int r = 0; r < S; ++r)
#pragma unroll
                for (int k = 0; k < N; ++k) a[r][k] = 0.0;
        }
    }
    if (singular && valid) atomicOr(status, 4u);
    __threadfence_block();
    __syncwarp();

    // ---- back substitution (as in version 2; every factor block comes through the cp.async ring)
    auto fetch_block = [&](int step) {
        if (step >= 0) {
            const double* src = fac + (size_t)step * NC * FS;
            double* dst = ring + (step % STAGES) * NC * FS;
#pragma unroll
            for (int i = 0; i < (NC * FS / 2 + GL - 1) / GL; ++i) {
                const int e = 2 * (lane + i * GL);
                if (e < NC * FS) {
                    const unsigned sa = (unsigned)__cvta_generic_to_shared(dst + e);
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(src + e) : "memory");
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
#pragma unroll
    for (int k = 1; k <= STAGES; ++k) fetch_block(nsteps - k);

    if (NRHS == 1) {
        for (int step = nsteps - 1; step >= 0; --step) {
            const int nleft = prob.nleft(step);
            const int nright = prob.nright(step);
            if (STAGES == 3)
                asm volatile("cp.async.wait_group 2;" ::: "memory");
            else
                asm volatile("cp.async.wait_group 1;" ::: "memory");
            __syncwarp(gmask);
            const double* facs = ring + (step % STAGES) * NC * FS;
            const int row = lane < nleft ? lane : 0;
            const double* my = facs + row * FS;
            double acc = my[4 * N];
            for (int jx = 0; jx < nright; ++jx) acc = fma(-my[NC + jx], xs[jx], acc);
            const double pinv = my[RL2];
            double myx = 0.0;
#pragma unroll
            for (int cc = NC - 1; cc >= 0; --cc) {
                if (cc < nleft) {
                    double xv = acc * pinv;
                    xv = __shfl_sync(gmask, xv, (int)gbase + cc);
                    if (lane < cc) acc = fma(-my[cc], xv, acc);
                    if (lane == cc) myx = xv;
                }
            }
            __syncwarp(gmask);
            if (lane < nleft) {
                xs[lane] = myx;
                if (valid) prob.store(step, lane, 0, myx);
            }
            __syncwarp(gmask);
            fetch_block(step - STAGES);
        }
    } else {
        const int r = lane < NRHS ? lane : 0;
        double xn[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) xn[c] = 0.0;
        for (int step = nsteps - 1; step >= 0; --step) {
            const int nleft = prob.nleft(step);
            const int nright = prob.nright(step);
            if (STAGES == 3)
                asm volatile("cp.async.wait_group 2;" ::: "memory");
            else
                asm volatile("cp.async.wait_group 1;" ::: "memory");
            __syncwarp(gmask);
            const double* facs = ring + (step % STAGES) * NC * FS;
            double acc[NC];
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                acc[c] = 0.0;
                if (c < nleft) {
                    const double* u = facs + c * FS;
                    double s = u[4 * N + r];
#pragma unroll
                    for (int jx = 0; jx < NC; ++jx)
                        if (jx < nright) s = fma(-u[NC + jx], xn[jx], s);
                    acc[c] = s;
                }
            }
#pragma unroll
            for (int cc = NC - 1; cc >= 0; --cc) {
                if (cc < nleft) {
                    const double xv = acc[cc] * facs[cc * FS + RL2];
                    acc[cc] = xv;
#pragma unroll
                    for (int c = 0; c < cc; ++c) acc[c] = fma(-facs[c * FS + cc], xv, acc[c]);
                }
            }
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                xn[c] = acc[c];
                if (valid && lane < NRHS && c < nleft) prob.store(step, c, r, acc[c]);
            }
            __syncwarp(gmask);
            fetch_block(step - STAGES);
        }
    }
}

template <int N>
__global__ void __launch_bounds__(BvpCfg3<N, 1>::WARPS_PER_BLOCK * 32, 4) k_bvp_v3(ChunkView V) {
    using C = BvpCfg3<N, 1>;
    extern __shared__ __align__(16) double smem[];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    long long prob = (long long)blockIdx.x * C::GROUPS_PER_BLOCK + wib;
    const long long nprob = (long long)V.nw * V.M;
    const bool valid = prob < nprob;
    if (!valid) prob = nprob - 1;
    const int w = (int)(prob / V.M), ms = (int)(prob % V.M);
    ForwardRows<N> rows(V, w, ms);
    double* fac = V.fac + (size_t)prob * V.fac_stride;
    staircase_solve_2d<N>(rows, smem + (size_t)wib * C::SMEM_DOUBLES_PER_GROUP, fac, lane, valid, V.status);
}

template <int N, int NRHS>
__global__ void __launch_bounds__(BvpCfg3<N, NRHS>::WARPS_PER_BLOCK * 32, 3) k_bvp_adjoint_v3(ChunkView V, int los0, int nbatch) {
    using C = BvpCfg3<N, NRHS>;
    extern __shared__ __align__(16) double smem[];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    long long gid = (long long)blockIdx.x * C::GROUPS_PER_BLOCK + wib;
    const long long ngroups = (long long)V.nw * V.M * nbatch;
    const bool valid = gid < ngroups;
    if (!valid) gid = ngroups - 1;
    const int batch = (int)(gid % nbatch);
    const long long prob = gid / nbatch;
    const int w = (int)(prob / V.M), ms = (int)(prob % V.M);
    AdjointRows<N, NRHS> rows(V, w, ms, los0 + batch * NRHS);
    double* fac = V.fac + (size_t)gid * V.fac_stride;
    staircase_solve_2d<N>(rows, smem + (size_t)wib * C::SMEM_DOUBLES_PER_GROUP, fac, lane, valid, V.status);
}

#ifndef DISCO_BVP4_MIN_BLOCKS
#define DISCO_BVP4_MIN_BLOCKS 4
#endif
template <int N, bool BLOCKED = false, bool TMA = false>
__global__ void __launch_bounds__(BvpCfg2<N, 1>::WARPS_PER_BLOCK * 32, BLOCKED ? DISCO_BVP4_MIN_BLOCKS : BvpCfg2<N, 1>::MIN_BLOCKS) k_bvp_v2(ChunkView V) {
    using C = BvpCfg2<N, 1>;
    extern __shared__ __align__(16) double smem[];
    const int lane_w = threadIdx.x & 31;
    const int gidx_in_block = threadIdx.x / C::GL;
    const int lane = threadIdx.x % C::GL;
    const unsigned gbase = (unsigned)((lane_w / C::GL) * C::GL);
    const unsigned gmask = (C::GL == 32) ? FULL_MASK : (((1u << C::GL) - 1u) << gbase);
    long long prob = (long long)blockIdx.x * C::GROUPS_PER_BLOCK + gidx_in_block;
    const long long nprob = (long long)V.nw * V.M;
    const bool valid = prob < nprob;
    if (!valid) prob = nprob - 1;
    const int w = (int)(prob / V.M), ms = (int)(prob % V.M);
    ForwardRows<N> rows(V, w, ms);
    double* fac = V.fac + (size_t)prob * V.fac_stride;
    double* lf = V.lfac ? V.lfac + (size_t)prob * V.lfac_stride : nullptr;
    if constexpr (BLOCKED)
        staircase_solve_v4<N, ForwardRows<N>, TMA>(rows, smem + (size_t)gidx_in_block * C::SMEM_DOUBLES_PER_GROUP, fac, lane, gbase, gmask,
                                                   valid, V.status, lf);
    else
        staircase_solve_v2<N>(rows, smem + (size_t)gidx_in_block * C::SMEM_DOUBLES_PER_GROUP, fac, lane, gbase, gmask, valid,
                              V.status, lf);
}

// -------------------------------------------------------------------------------------------------
// K3^T by factor reuse: A^T z = wvec(los) from the factors of the FORWARD elimination - what the reference does with
// dgbtrs('T') after dgbsv (sktran_do_rte.cpp:1793-1836) - instead of a second factorisation of A^T.
// The forward kernel (k_bvp_v2 with V.lfac set) leaves, per pivot t = (step, c):
//   fac [t][FS]: row t of U right of the diagonal (window columns c+1 .. 4N-1 of the step) and 1 / pivot;
//   lfac[t][LS]: the multiplier f_t[i] of every panel lane i (0 where the lane held no candidate row) and the
//                pivot's (lane, row of A).
// With E_t = I - f_t e_{p_t}^T the elimination reads E_n .. E_1 A = P^T U, so
//   phase 1 (t ascending):  U^T y = wvec        - y_t = w_t / u_tt, then w_cc -= u_{t,cc} y_t over the window;
//   phase 2 (t descending): z = E_1^T .. E_n^T P^T y - z[row_t] = y_t - sum_i f_t[i] v[i], v[lane_t] = z[row_t],
// where v[i] is the value of the row currently (in elimination order) held by panel lane i.
// Mapping: one lane per line of sight, `glt` lanes per (wavelength, order) and floor(32 / glt) problems per warp;
// every factor operand is a shared-memory load that is uniform over the problem's lanes (no shuffles), the blocks
// arrive through a two-slot cp.async ring.  4N + 2N + 3N doubles of state per lane.
// -------------------------------------------------------------------------------------------------
template <int N>
struct TsolveCfg {
    using F = BvpCfg2<N, 1>;
    static constexpr int NC = F::NC, FS = F::FS, LS = F::LS, ROWS = F::ROWS, RL2 = F::RL2;
    static constexpr int STAGES = 2;
    static constexpr int WARPS_PER_BLOCK = 2;
    // ring of two factor blocks + 4 doubles: the problems of a warp read their rings at the same offset in the same
    // instruction, a 32-byte skew per problem keeps those 16-byte broadcasts on different banks
    __host__ __device__ static constexpr int smem_doubles_per_group(int) { return STAGES * NC * FS + 4; }
    static_assert(LS <= FS, "multiplier blocks share the ring slots of the factor blocks");
};

// y and z are LOS-fastest in global memory ([t][nlos]) so that the lanes of a problem (one per LOS) touch one or two
// 128-byte lines per access instead of one line each; the wvec rows ([nlos][L][2N], LOS slowest - the layout K2c
// writes and K4 streams) are read per lane one step ahead.  (Staging them cooperatively through shared memory was
// measured slower: the extra 2.9 KB per problem costs a third of the resident problems.)
template <int N>
__global__ void __launch_bounds__(TsolveCfg<N>::WARPS_PER_BLOCK * 32) k_bvp_tsolve(ChunkView V, int glt, int gpw, int nbatch) {
    using C = TsolveCfg<N>;
    constexpr int NC = C::NC, FS = C::FS, LS = C::LS, ROWS = C::ROWS, RL2 = C::RL2, STAGES = C::STAGES;
    extern __shared__ __align__(16) double smem[];
    const int L = V.T.L, nlos = V.T.nlos;
    const int warp = threadIdx.x >> 5, lane_w = threadIdx.x & 31;
    const int gw_raw = lane_w / glt;
    const bool in_group = gw_raw < gpw;          // lanes past the last whole group idle along (no copies, no stores)
    const int gw = in_group ? gw_raw : gpw - 1;
    const int r = in_group ? lane_w - gw * glt : 0;
    const int gslot = warp * gpw + gw;
    long long gid = (long long)blockIdx.x * (C::WARPS_PER_BLOCK * gpw) + gslot;
    const long long ngroups = (long long)V.nw * V.M * nbatch;
    const bool valid_group = gid < ngroups;
    if (!valid_group) gid = ngroups - 1;
    const int batch = (int)(gid % nbatch);
    const long long prob = gid / nbatch;
    const int los0 = batch * glt;
    int los = los0 + r;
    const bool store_ok = valid_group && in_group && los < nlos;
    if (los >= nlos) los = nlos - 1;
    double* ring = smem + (size_t)gslot * C::smem_doubles_per_group(glt);
    const double* fac = V.fac + (size_t)prob * V.fac_stride;
    const double* lfac = V.lfac + (size_t)prob * V.lfac_stride;
    const size_t nrow = (size_t)2 * N * L;
    const double* wv = V.wvec + ((size_t)prob * nlos + los) * nrow;          // this lane's [L][2N]
    double* yb = V.yadj + (size_t)prob * nrow * nlos + los;                  // y[t * nlos]
    double* zb = V.zadj + (size_t)prob * nrow * nlos + los;                  // z[row * nlos]

    auto fetch = [&](const double* src, int ndoubles, int slot) {
        if (src != nullptr && in_group) {
            double* dst = ring + slot * NC * FS;
            for (int e = 2 * r; e < ndoubles; e += 2 * glt) {
                const unsigned sa = (unsigned)__cvta_generic_to_shared(dst + e);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(src + e) : "memory");
            }
        }
    };
    auto commit = [&]() { asm volatile("cp.async.commit_group;" ::: "memory"); };

    // ---- phase 1: U^T y = w, window of 4N right-hand-side entries per lane
    double win[4 * N];
    {
        const double2* w2 = reinterpret_cast<const double2*>(wv);
#pragma unroll
        for (int j = 0; j < N; ++j) {
            const double2 t = w2[j];
            win[2 * j] = t.x;
            win[2 * j + 1] = t.y;
        }
#pragma unroll
        for (int j = 0; j < N; ++j) {
            double2 t = make_double2(0.0, 0.0);
            if (L > 1) t = w2[N + j];
            win[NC + 2 * j] = t.x;
            win[NC + 2 * j + 1] = t.y;
        }
    }
    fetch(fac, NC * FS, 0);
    commit();
    for (int step = 0; step < L; ++step) {
        fetch(step + 1 < L ? fac + (size_t)(step + 1) * NC * FS : nullptr, NC * FS, (step + 1) % STAGES);
        commit();
        // right-hand-side entries that enter the window after this step: requested now, consumed at the end
        double2 nxt[N];
        {
            const bool more = step + 2 < L;
            const double2* w2 = reinterpret_cast<const double2*>(wv + (size_t)(more ? step + 2 : 0) * NC);
#pragma unroll
            for (int j = 0; j < N; ++j) {
                nxt[j] = make_double2(0.0, 0.0);
                if (more) nxt[j] = w2[j];
            }
        }
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        __syncwarp();
        const double* ub = ring + (step % STAGES) * NC * FS;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            const double* u = ub + c * FS;
            const double y = win[c] * u[RL2];
            win[c] = y;
#pragma unroll
            for (int cc = c + 1; cc < 4 * N; ++cc) win[cc] = fma(-u[cc], y, win[cc]);
        }
        if (store_ok) {
            double* yo = yb + (size_t)step * NC * nlos;
#pragma unroll
            for (int c = 0; c < NC; ++c) yo[(size_t)c * nlos] = win[c];
        }
#pragma unroll
        for (int j = 0; j < N; ++j) {
            win[2 * j] = win[NC + 2 * j];
            win[2 * j + 1] = win[NC + 2 * j + 1];
            win[NC + 2 * j] = nxt[j].x;
            win[NC + 2 * j + 1] = nxt[j].y;
        }
        __syncwarp();
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();

    // ---- phase 2: z = E_1^T .. E_n^T P^T y
    double v[ROWS];
#pragma unroll
    for (int i = 0; i < ROWS; ++i) v[i] = 0.0;
    fetch(lfac + (size_t)(L - 1) * NC * LS, NC * LS, (L - 1) % STAGES);
    commit();
    // y of a step: this lane's own phase-1 stores (same thread, same addresses), requested one step ahead; lanes
    // without a line of sight carry zeros
    auto load_y = [&](int step, double* dst) {
        const double* yi = yb + (size_t)(step >= 0 ? step : 0) * NC * nlos;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            dst[c] = 0.0;
            if (store_ok && step >= 0) dst[c] = yi[(size_t)c * nlos];
        }
    };
    double ycur[NC];
    load_y(L - 1, ycur);
    for (int step = L - 1; step >= 0; --step) {
        fetch(step > 0 ? lfac + (size_t)(step - 1) * NC * LS : nullptr, NC * LS, (step + STAGES - 1) % STAGES);
        commit();
        double ynext[NC];
        load_y(step - 1, ynext);
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        __syncwarp();
        const double* lb = ring + (step % STAGES) * NC * FS;
#pragma unroll
        for (int c = NC - 1; c >= 0; --c) {
            const double* f = lb + c * LS;
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
            for (int i = 0; i < ROWS; ++i) {
                const double t = f[i];
                if ((i & 3) == 0) s0 = fma(t, v[i], s0);
                else if ((i & 3) == 1) s1 = fma(t, v[i], s1);
                else if ((i & 3) == 2) s2 = fma(t, v[i], s2);
                else s3 = fma(t, v[i], s3);
            }
            const double z = ycur[c] - ((s0 + s1) + (s2 + s3));
            const double rec = f[ROWS];
            const int plane = __double2loint(rec), prow = __double2hiint(rec);
#pragma unroll
            for (int i = 0; i < ROWS; ++i) v[i] = (i == plane) ? z : v[i];
            if (store_ok) zb[(size_t)prow * nlos] = z;
        }
#pragma unroll
        for (int c = 0; c < NC; ++c) ycur[c] = ynext[c];
        __syncwarp();
    }
}

template <int N, int NRHS>
__global__ void __launch_bounds__(BvpCfg2<N, NRHS>::WARPS_PER_BLOCK * 32, BvpCfg2<N, NRHS>::MIN_BLOCKS) k_bvp_adjoint_v2(ChunkView V, int los0, int nbatch) {
    using C = BvpCfg2<N, NRHS>;
    extern __shared__ __align__(16) double smem[];
    const int lane_w = threadIdx.x & 31;
    const int gidx_in_block = threadIdx.x / C::GL;
    const int lane = threadIdx.x % C::GL;
    const unsigned gbase = (unsigned)((lane_w / C::GL) * C::GL);
    const unsigned gmask = (C::GL == 32) ? FULL_MASK : (((1u << C::GL) - 1u) << gbase);
    long long gid = (long long)blockIdx.x * C::GROUPS_PER_BLOCK + gidx_in_block;
    const long long ngroups = (long long)V.nw * V.M * nbatch;
    const bool valid = gid < ngroups;
    if (!valid) gid = ngroups - 1;
    const int batch = (int)(gid % nbatch);
    const long long prob = gid / nbatch;
    const int w = (int)(prob / V.M), ms = (int)(prob % V.M);
    AdjointRows<N, NRHS> rows(V, w, ms, los0 + batch * NRHS);
    double* fac = V.fac + (size_t)gid * V.fac_stride;
    staircase_solve_v2<N>(rows, smem + (size_t)gidx_in_block * C::SMEM_DOUBLES_PER_GROUP, fac, lane, gbase, gmask, valid,
                          V.status);
}

template <int N>
__global__ void __launch_bounds__(BvpCfg<N, 1>::WARPS_PER_BLOCK * 32) k_bvp(ChunkView V) {
    using C = BvpCfg<N, 1>;
    extern __shared__ double smem[];
    const int lane_w = threadIdx.x & 31;
    const int gidx_in_block = threadIdx.x / C::GL;
    const int lane = threadIdx.x % C::GL;
    const unsigned gbase = (unsigned)((lane_w / C::GL) * C::GL);
    const unsigned gmask = (C::GL == 32) ? FULL_MASK : (((1u << C::GL) - 1u) << gbase);
    long long prob = (long long)blockIdx.x * C::GROUPS_PER_BLOCK + gidx_in_block;
    const long long nprob = (long long)V.nw * V.M;
    const bool valid = prob < nprob;
    if (!valid) prob = nprob - 1;
    const int w = (int)(prob / V.M), ms = (int)(prob % V.M);
    ForwardRows<N> rows(V, w, ms);
    double* fac = V.fac + (size_t)prob * V.fac_stride;
    staircase_solve<N>(rows, smem + (size_t)gidx_in_block * C::SMEM_DOUBLES_PER_GROUP, fac, lane, gbase, gmask, valid,
                       V.status);
}

// one group per (w, m, batch of NRHS lines of sight)
template <int N, int NRHS>
__global__ void __launch_bounds__(BvpCfg<N, NRHS>::WARPS_PER_BLOCK * 32) k_bvp_adjoint(ChunkView V, int los0, int nbatch) {
    using C = BvpCfg<N, NRHS>;
    extern __shared__ double smem[];
    const int lane_w = threadIdx.x & 31;
    const int gidx_in_block = threadIdx.x / C::GL;
    const int lane = threadIdx.x % C::GL;
    const unsigned gbase = (unsigned)((lane_w / C::GL) * C::GL);
    const unsigned gmask = (C::GL == 32) ? FULL_MASK : (((1u << C::GL) - 1u) << gbase);
    long long gid = (long long)blockIdx.x * C::GROUPS_PER_BLOCK + gidx_in_block;
    const long long ngroups = (long long)V.nw * V.M * nbatch;
    const bool valid = gid < ngroups;
    if (!valid) gid = ngroups - 1;
    const int batch = (int)(gid % nbatch);
    const long long prob = gid / nbatch;
    const int w = (int)(prob / V.M), ms = (int)(prob % V.M);
    AdjointRows<N, NRHS> rows(V, w, ms, los0 + batch * NRHS);
    double* fac = V.fac + (size_t)gid * V.fac_stride;
    staircase_solve<N>(rows, smem + (size_t)gidx_in_block * C::SMEM_DOUBLES_PER_GROUP, fac, lane, gbase, gmask, valid,
                       V.status);
}

// Forward solve of NRHS solar geometries sharing one matrix (ForwardRowsMulti): the column-by-column elimination with
// NRHS right-hand-side columns, row-owner back substitution per right-hand side.
template <int N, int NRHS>
__global__ void __launch_bounds__(BvpCfg2<N, NRHS>::WARPS_PER_BLOCK * 32, BvpCfg2<N, NRHS>::MIN_BLOCKS) k_bvp_multi(ChunkView V) {
    using C = BvpCfg2<N, NRHS>;
    extern __shared__ __align__(16) double smem[];
    const int lane_w = threadIdx.x & 31;
    const int gidx_in_block = threadIdx.x / C::GL;
    const int lane = threadIdx.x % C::GL;
    const unsigned gbase = (unsigned)((lane_w / C::GL) * C::GL);
    const unsigned gmask = (C::GL == 32) ? FULL_MASK : (((1u << C::GL) - 1u) << gbase);
    long long prob = (long long)blockIdx.x * C::GROUPS_PER_BLOCK + gidx_in_block;
    const long long nprob = (long long)V.nw * V.M;
    const bool valid = prob < nprob;
    if (!valid) prob = nprob - 1;
    const int w = (int)(prob / V.M), ms = (int)(prob % V.M);
    ForwardRowsMulti<N, NRHS> rows(V, w, ms);
    double* fac = V.fac + (size_t)prob * V.fac_stride;
    staircase_solve_v2<N>(rows, smem + (size_t)gidx_in_block * C::SMEM_DOUBLES_PER_GROUP, fac, lane, gbase, gmask, valid, V.status);
}
template <int N, int NRHS>
static void launch_bvp_multi_nr(const ChunkView& V, cudaStream_t s) {
    using C = BvpCfg2<N, NRHS>;
    const long long nprob = (long long)V.nw * V.M;
    const size_t smem = (size_t)C::GROUPS_PER_BLOCK * C::SMEM_DOUBLES_PER_GROUP * sizeof(double);
    static DeviceOnce attr_set;
    if (attr_set.first()) cudaFuncSetAttribute(k_bvp_multi<N, NRHS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_bvp_multi<N, NRHS><<<(unsigned)((nprob + C::GROUPS_PER_BLOCK - 1) / C::GROUPS_PER_BLOCK), C::WARPS_PER_BLOCK * 32, smem, s>>>(V);
}
template <int N>
static void launch_bvp_multi_n(const ChunkView& V, cudaStream_t s) {
    if constexpr (3 * N <= 32) {
        switch (V.nsza) {
            case 2: launch_bvp_multi_nr<N, 2>(V, s); break;
            case 3: launch_bvp_multi_nr<N, 3>(V, s); break;
            case 4: launch_bvp_multi_nr<N, 4>(V, s); break;
            default: break;
        }
    }
}

// SK_B200_BVP=3 selects the 2D-distributed elimination (version 3) for N = 8.  Measured on B200 it is within 10 % of
// version 2 (bvp 10.5 vs 9.2, adjoint 16.1 vs 15.2 ms per 1000 wavelengths): version 2 is bound by shared-memory
// instruction throughput, version 3 by issue/latency of its ~110 instructions per pivot (profiles/README.md).
static bool bvp_use_2d() {
    static const bool v = [] {
        const char* e = std::getenv("SK_B200_BVP");
        return e && e[0] == '3';
    }();
    return v;
}

// SK_B200_BVP=2 selects the column-by-column elimination (version 2) instead of the blocked one (version 4)
static bool bvp_blocked() {
    static const bool v = [] {
        const char* e = std::getenv("SK_B200_BVP");
        return !(e && e[0] == '2');
    }();
    return v;
}

template <int N>
static void launch_bvp_n(const ChunkView& V, cudaStream_t s) {
    const long long nprob = (long long)V.nw * V.M;
    if (N == 8 && bvp_use_2d()) {
        using C = BvpCfg3<N == 8 ? N : 8, 1>;
        const size_t smem = (size_t)C::GROUPS_PER_BLOCK * C::SMEM_DOUBLES_PER_GROUP * sizeof(double);
        static DeviceOnce attr_set;
        if (attr_set.first()) {
            cudaFuncSetAttribute(k_bvp_v3<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        }
        k_bvp_v3<8><<<(unsigned)((nprob + C::GROUPS_PER_BLOCK - 1) / C::GROUPS_PER_BLOCK), C::WARPS_PER_BLOCK * 32, smem, s>>>(V);
    } else if constexpr (3 * N <= 32) {
        using C = BvpCfg2<N, 1>;
        const size_t smem = (size_t)C::GROUPS_PER_BLOCK * C::SMEM_DOUBLES_PER_GROUP * sizeof(double);
        static DeviceOnce attr_set;
        if (attr_set.first()) {
            cudaFuncSetAttribute(k_bvp_v2<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        }
        // SK_B200_BVP_TMA=1: the rows of every step come from layer tiles staged in shared memory by the TMA engine
        // (per-row cp.async.bulk into padded rows, mbarrier per tile slot) instead of L1-prefetched global loads.
        // Measured on B200 (C5 shape, 4000 wavelengths): 37.1 ms against 35.4 ms for the global loads - 18 small bulk
        // copies per layer and warp cost more than the 64-byte-per-lane row loads they replace - so it is opt-in.
        static const bool use_tma = [] {
            const char* e = std::getenv("SK_B200_BVP_TMA");
            return e && e[0] == '1';
        }();
        if (bvp_blocked() && use_tma && N >= 2 && C::GL >= 2 * N + 2) {
            static DeviceOnce attr5_set;
            if (attr5_set.first()) {
                cudaFuncSetAttribute(k_bvp_v2<N, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            }
            k_bvp_v2<N, true, true><<<(unsigned)((nprob + C::GROUPS_PER_BLOCK - 1) / C::GROUPS_PER_BLOCK), C::WARPS_PER_BLOCK * 32, smem, s>>>(V);
        } else if (bvp_blocked()) {
            static DeviceOnce attr4_set;
            if (attr4_set.first()) {
                cudaFuncSetAttribute(k_bvp_v2<N, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            }
            k_bvp_v2<N, true><<<(unsigned)((nprob + C::GROUPS_PER_BLOCK - 1) / C::GROUPS_PER_BLOCK), C::WARPS_PER_BLOCK * 32, smem, s>>>(V);
        } else {
            k_bvp_v2<N><<<(unsigned)((nprob + C::GROUPS_PER_BLOCK - 1) / C::GROUPS_PER_BLOCK), C::WARPS_PER_BLOCK * 32, smem, s>>>(V);
        }
    } else {
        using C = BvpCfg<N, 1>;
        const size_t smem = (size_t)C::GROUPS_PER_BLOCK * C::SMEM_DOUBLES_PER_GROUP * sizeof(double);
        static DeviceOnce attr_set;
        if (attr_set.first()) {
            cudaFuncSetAttribute(k_bvp<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        }
        k_bvp<N><<<(unsigned)((nprob + C::GROUPS_PER_BLOCK - 1) / C::GROUPS_PER_BLOCK), C::WARPS_PER_BLOCK * 32, smem, s>>>(V);
    }
}
template <int N, int NRHS>
static void launch_adj_batch(const ChunkView& V, int los0, int nbatch, cudaStream_t s) {
    const long long ngroups = (long long)V.nw * V.M * nbatch;
    if (N == 8 && bvp_use_2d()) {
        using C = BvpCfg3<N == 8 ? N : 8, NRHS>;
        const size_t smem = (size_t)C::GROUPS_PER_BLOCK * C::SMEM_DOUBLES_PER_GROUP * sizeof(double);
        static DeviceOnce attr_set;
        if (attr_set.first()) {
            cudaFuncSetAttribute(k_bvp_adjoint_v3<8, NRHS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        }
        k_bvp_adjoint_v3<8, NRHS><<<(unsigned)((ngroups + C::GROUPS_PER_BLOCK - 1) / C::GROUPS_PER_BLOCK),
                                    C::WARPS_PER_BLOCK * 32, smem, s>>>(V, los0, nbatch);
    } else if constexpr (3 * N <= 32) {
        using C = BvpCfg2<N, NRHS>;
        const size_t smem = (size_t)C::GROUPS_PER_BLOCK * C::SMEM_DOUBLES_PER_GROUP * sizeof(double);
        static DeviceOnce attr_set;
        if (attr_set.first()) {
            cudaFuncSetAttribute(k_bvp_adjoint_v2<N, NRHS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        }
        k_bvp_adjoint_v2<N, NRHS><<<(unsigned)((ngroups + C::GROUPS_PER_BLOCK - 1) / C::GROUPS_PER_BLOCK),
                                    C::WARPS_PER_BLOCK * 32, smem, s>>>(V, los0, nbatch);
    } else {
        using C = BvpCfg<N, NRHS>;
        const size_t smem = (size_t)C::GROUPS_PER_BLOCK * C::SMEM_DOUBLES_PER_GROUP * sizeof(double);
        static DeviceOnce attr_set;
        if (attr_set.first()) {
            cudaFuncSetAttribute(k_bvp_adjoint<N, NRHS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        }
        k_bvp_adjoint<N, NRHS><<<(unsigned)((ngroups + C::GROUPS_PER_BLOCK - 1) / C::GROUPS_PER_BLOCK),
                                 C::WARPS_PER_BLOCK * 32, smem, s>>>(V, los0, nbatch);
    }
}
// Lines of sight are solved in batches of NRHS right-hand sides per factorisation of A^T (a partially filled
// last batch carries zero columns): 4 when there are at most 4 lines of sight, else 10.
static int adj_rhs_for(int nlos) { return nlos <= 4 ? 4 : 10; }
template <int N>
static void launch_bvp_adjoint_n(const ChunkView& V, cudaStream_t s) {
    const int nlos = V.T.nlos;
    if constexpr (3 * N <= 32) {
        if (V.lfac != nullptr) {  // the forward solve kept its multipliers: transposed solve, no second factorisation
            using C = TsolveCfg<N>;
            const int glt = tsolve_lanes(nlos);
            const int gpw = tsolve_groups_per_warp(N, glt);
            const int nbatch = (nlos + glt - 1) / glt;
            const long long ngroups = (long long)V.nw * V.M * nbatch;
            const int gpb = C::WARPS_PER_BLOCK * gpw;
            const size_t smem = (size_t)gpb * C::smem_doubles_per_group(glt) * sizeof(double);
            static DeviceOnce attr_set;
            if (attr_set.first()) {
                cudaFuncSetAttribute(k_bvp_tsolve<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
            }
            k_bvp_tsolve<N><<<(unsigned)((ngroups + gpb - 1) / gpb), C::WARPS_PER_BLOCK * 32, smem, s>>>(V, glt, gpw, nbatch);
            return;
        }
    }
    const int nrhs = adj_rhs_for(nlos);
    const int nbatch = (nlos + nrhs - 1) / nrhs;
    if (nrhs == 4)
        launch_adj_batch<N, 4>(V, 0, nbatch, s);
    else
        launch_adj_batch<N, 10>(V, 0, nbatch, s);
}

}  // namespace disco
