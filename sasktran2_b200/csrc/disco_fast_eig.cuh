// Register-resident homogeneous solve (N <= 8 streams per hemisphere): thread per (wavelength, order, layer).
//
// The math is disco_core.h's layer_solve (what: sktran_do_rte.cpp:383-553, sktran_do_lpproduct.h:165-263) split
// into two kernels so that the live state of each thread stays below ~110 doubles and nothing spills:
//
//   k_eig_setup   S~+ and S~- (packed symmetric, accumulated one Legendre order at a time from a shared-memory
//                 table A[l][i] = sqrt(w_i / mu_i) P_l^m(mu_i) that is uniform over the block), Cholesky
//                 S~- = H H^T in place, C = H^T S~+ H one column at a time  ->  S~+, H, C as [element][problem]
//                 planes (coalesced for a thread-per-problem consumer)
//   k_eig_jacobi  cyclic Jacobi on C in a round-robin (tournament) ordering - N/2 independent rotations per round
//                 give the FP64 pipe instruction-level parallelism -, eigenvectors accumulated in registers,
//                 then X~ = H Z, S~+ X~ and W+- = D^-1 (X~ +- S~+ X~ / k) / 2 written in the [problem][i][j]
//                 layout every consumer reads.
//
// FP64 pipe facts these choices rest on (profiles/microbench_fp64_r01.txt): DFMA 36.5 TFLOP/s = 2 warp
// instructions / clk / SM; a 64-bit shuffle costs 4 DFMA slots; an LDS.128 broadcast 2.4.  Hence: no shuffles, all
// matrix operands in registers, table operands through uniform LDS.128.
#pragma once
#include "disco_kernels.cuh"

namespace disco {

__host__ __device__ constexpr int tri_idx(int i, int j) { return i >= j ? i * (i + 1) / 2 + j : j * (j + 1) / 2 + i; }

// plane index of problem (w, ms, p) in the [element][problem] staging arrays: all problems of one azimuth slot are
// contiguous so that a block (one slot) reads and writes coalesced
__device__ __forceinline__ size_t plane_index(const ChunkView& V, int ms, long long q) {
    return (size_t)ms * ((size_t)V.nw * V.T.L) + (size_t)q;
}

template <int N>
__global__ void __launch_bounds__(128) k_eig_setup(ChunkView V) {
    constexpr int NSTR = 2 * N, TS = N * (N + 1) / 2;
    __shared__ __align__(16) double sA[NSTR][N];
    __shared__ double sInvMu[N];
    const int ms = blockIdx.y;
    const int m = V.m_list[ms];
    for (int e = threadIdx.x; e < NSTR * N; e += blockDim.x) {
        const int l = e / N, i = e % N;
        sA[l][i] = sqrt(V.T.wt[i] / V.T.mu[i]) * V.T.lp_mu[((size_t)m * N + i) * NSTR + l];
    }
    if (threadIdx.x < N) sInvMu[threadIdx.x] = 1.0 / V.T.mu[threadIdx.x];
    __syncthreads();
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // w * L + p
    const long long nq = (long long)V.nw * V.T.L;
    if (q >= nq) return;
    const double ssa = V.lay_ssa[q];
    const double* __restrict__ beta = V.lay_beta + (size_t)q * NSTR;
    double Sp[TS], Sm[TS];
#pragma unroll
    for (int e = 0; e < TS; ++e) Sp[e] = Sm[e] = 0.0;
    for (int l = m; l < NSTR; l += 2) {
        {
            const double t = -ssa * beta[l];
            double a[N];
#pragma unroll
            for (int i = 0; i < N; ++i) a[i] = sA[l][i];
#pragma unroll
            for (int i = 0; i < N; ++i) {
                const double ta = t * a[i];
#pragma unroll
                for (int j = 0; j <= i; ++j) Sp[tri_idx(i, j)] = fma(ta, a[j], Sp[tri_idx(i, j)]);
            }
        }
        if (l + 1 < NSTR) {
            const double t = -ssa * beta[l + 1];
            double a[N];
#pragma unroll
            for (int i = 0; i < N; ++i) a[i] = sA[l + 1][i];
#pragma unroll
            for (int i = 0; i < N; ++i) {
                const double ta = t * a[i];
#pragma unroll
                for (int j = 0; j <= i; ++j) Sm[tri_idx(i, j)] = fma(ta, a[j], Sm[tri_idx(i, j)]);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < N; ++i) {
        Sp[tri_idx(i, i)] += sInvMu[i];
        Sm[tri_idx(i, i)] += sInvMu[i];
    }
    const size_t P = (size_t)V.nw * V.M * V.T.L;
    const size_t pi = plane_index(V, ms, q);
#pragma unroll
    for (int e = 0; e < TS; ++e) V.eigS[(size_t)e * P + pi] = Sp[e];
    // Cholesky S~- = H H^T, in place (lower triangle)
    bool bad = false;
#pragma unroll
    for (int j = 0; j < N; ++j) {
        double s = Sm[tri_idx(j, j)];
#pragma unroll
        for (int r = 0; r < j; ++r) s = fma(-Sm[tri_idx(j, r)], Sm[tri_idx(j, r)], s);
        if (!(s > 0.0)) {
            bad = true;
            s = 1e-300;
        }
        const double hjj = sqrt(s);
        const double inv = 1.0 / hjj;
        Sm[tri_idx(j, j)] = hjj;
#pragma unroll
        for (int i = j + 1; i < N; ++i) {
            double t = Sm[tri_idx(i, j)];
#pragma unroll
            for (int r = 0; r < j; ++r) t = fma(-Sm[tri_idx(i, r)], Sm[tri_idx(j, r)], t);
            Sm[tri_idx(i, j)] = t * inv;
        }
    }
    if (bad) atomicOr(V.status, 1u);
#pragma unroll
    for (int e = 0; e < TS; ++e) V.eigH[(size_t)e * P + pi] = Sm[e];
    // C = H^T (S~+ H), column by column; only the lower triangle is kept
#pragma unroll
    for (int j = 0; j < N; ++j) {
        double T[N];
#pragma unroll
        for (int r = 0; r < N; ++r) {
            double s = 0.0;
#pragma unroll
            for (int c = j; c < N; ++c) s = fma(Sp[tri_idx(r, c)], Sm[tri_idx(c, j)], s);
            T[r] = s;
        }
#pragma unroll
        for (int i = j; i < N; ++i) {
            double s = 0.0;
#pragma unroll
            for (int r = i; r < N; ++r) s = fma(Sm[tri_idx(r, i)], T[r], s);
            V.eigC[(size_t)tri_idx(i, j) * P + pi] = s;
        }
    }
}

// Branch-free reciprocal square root / reciprocal for positive normal arguments inside the float range (Jacobi
// rotation parameters): single-precision MUFU seed + two Newton steps in double.  The library sqrt / division carry
// slow-path calls that force register spills around every rotation.
__device__ __forceinline__ double rsqrt_nr(double x) {
    float seed;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(seed) : "f"((float)x));
    double y = (double)seed;
    double e = fma(-(x * y), y, 1.0);
    y = fma(0.5 * y, e, y);
    e = fma(-(x * y), y, 1.0);
    return fma(0.5 * y, e, y);
}
__device__ __forceinline__ double div_nr(double num, double den) {
    float seed;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(seed) : "f"((float)den));
    double y = (double)seed;
    double e = fma(-den, y, 1.0);
    y = fma(y, e, y);
    e = fma(-den, y, 1.0);
    y = fma(y, e, y);
    const double t = num * y;
    return fma(fma(-den, t, num), y, t);
}

// One Jacobi rotation of the pair (P_, Q_) on the packed symmetric C and the eigenvector accumulator Z.
// Returns true if the rotation was not negligible (same threshold as disco_core.h's jacobi_eig).
template <int N, int P_, int Q_>
__device__ __forceinline__ bool jacobi_rotate(double (&C)[N * (N + 1) / 2], double (&Z)[N * N]) {
    static_assert(P_ < Q_, "ordered pair");
    const double apq = C[tri_idx(Q_, P_)];
    const double app = C[tri_idx(P_, P_)], aqq = C[tri_idx(Q_, Q_)];
    const bool act = apq * apq > 1e-34 * fabs(app * aqq);
    const double d = aqq - app;
    const double h2 = fma(d, d, 4.0 * apq * apq);
    const double hy = rsqrt_nr(h2);
    double h = h2 * hy;
    h = fma(0.5 * hy, fma(-h, h, h2), h);  // sqrt(h2)
    double t = div_nr(2.0 * apq, d + copysign(h, d));
    if (!act) t = 0.0;  // also discards the NaN of h2 == 0
    const double c = rsqrt_nr(fma(t, t, 1.0));
    const double s = t * c;
    C[tri_idx(P_, P_)] = fma(-t, apq, app);
    C[tri_idx(Q_, Q_)] = fma(t, apq, aqq);
    C[tri_idx(Q_, P_)] = 0.0;
#pragma unroll
    for (int r = 0; r < N; ++r) {
        if (r != P_ && r != Q_) {
            const double arp = C[tri_idx(r, P_)], arq = C[tri_idx(r, Q_)];
            C[tri_idx(r, P_)] = fma(c, arp, -s * arq);
            C[tri_idx(r, Q_)] = fma(s, arp, c * arq);
        }
    }
#pragma unroll
    for (int r = 0; r < N; ++r) {
        const double v0 = Z[r * N + P_], v1 = Z[r * N + Q_];
        Z[r * N + P_] = fma(c, v0, -s * v1);
        Z[r * N + Q_] = fma(s, v0, c * v1);
    }
    return act;
}

// round-robin ("circle") tournament: round R pairs player (R + K) mod (N-1) with (R - K) mod (N-1); K = 0 pairs
// R with the fixed player N-1
template <int N, int R, int K>
struct RoundRobin {
    static constexpr int a = (K == 0) ? R : (R + K) % (N - 1);
    static constexpr int b = (K == 0) ? (N - 1) : (R - K + (N - 1)) % (N - 1);
    static constexpr int p = a < b ? a : b;
    static constexpr int q = a < b ? b : a;
};

template <int N, int R, int K>
struct JacobiPairs {
    __device__ __forceinline__ static bool run(double (&C)[N * (N + 1) / 2], double (&Z)[N * N]) {
        const bool r = jacobi_rotate<N, RoundRobin<N, R, K>::p, RoundRobin<N, R, K>::q>(C, Z);
        return JacobiPairs<N, R, K + 1>::run(C, Z) | r;
    }
};
template <int N, int R>
struct JacobiPairs<N, R, N / 2> {
    __device__ __forceinline__ static bool run(double (&)[N * (N + 1) / 2], double (&)[N * N]) { return false; }
};
template <int N, int R>
struct JacobiRounds {
    __device__ __forceinline__ static bool run(double (&C)[N * (N + 1) / 2], double (&Z)[N * N]) {
        const bool r = JacobiPairs<N, R, 0>::run(C, Z);
        return JacobiRounds<N, R + 1>::run(C, Z) | r;
    }
};
template <int N>
struct JacobiRounds<N, N - 1> {
    __device__ __forceinline__ static bool run(double (&)[N * (N + 1) / 2], double (&)[N * N]) { return false; }
};

// Relabelling between the rounds of a sweep: the circle tournament of round R + 1 is the one of round R with every
// player but the last moved down by one, so permuting C (rows and columns) and the columns of Z by
// pi(x) = x - 1 mod (N - 1), pi(N - 1) = N - 1 lets every round run the pairs of round 0.  All of C and Z is rewritten
// by the round's rotations anyway: the permutation is a renaming of their results (register moves at most), and the loop
// body is one round (a seventh of the 4 096 straight-line instructions whose instruction-cache misses were 2.2 of this
// kernel's 5 stall cycles per issue).  After N - 1 rounds (one sweep) the labels are back where they started.
template <int N>
__device__ __forceinline__ void jacobi_relabel(double (&C)[N * (N + 1) / 2], double (&Z)[N * N]) {
    constexpr int TS = N * (N + 1) / 2;
    double Cn[TS], Zn[N * N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const int pi = (i == N - 1) ? N - 1 : (i + N - 2) % (N - 1);
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            const int pj = (j == N - 1) ? N - 1 : (j + N - 2) % (N - 1);
            Cn[tri_idx(pi, pj)] = C[tri_idx(i, j)];
        }
#pragma unroll
        for (int r = 0; r < N; ++r) Zn[r * N + pi] = Z[r * N + i];
    }
#pragma unroll
    for (int e = 0; e < TS; ++e) C[e] = Cn[e];
#pragma unroll
    for (int e = 0; e < N * N; ++e) Z[e] = Zn[e];
}

template <int N, bool ROLLED>
__global__ void __launch_bounds__(128) k_eig_jacobi(ChunkView V) {
    constexpr int TS = N * (N + 1) / 2;
    const int ms = blockIdx.y;
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // w * L + p
    const long long nq = (long long)V.nw * V.T.L;
    const bool valid = q < nq;
    const long long qc = valid ? q : nq - 1;
    const size_t P = (size_t)V.nw * V.M * V.T.L;
    const size_t pi = plane_index(V, ms, qc);
    double C[TS], Z[N * N];
#pragma unroll
    for (int e = 0; e < TS; ++e) C[e] = V.eigC[(size_t)e * P + pi];
#pragma unroll
    for (int e = 0; e < N * N; ++e) Z[e] = (e / N == e % N) ? 1.0 : 0.0;
    if (N > 1) {
        for (int sweep = 0; sweep < 40; ++sweep) {
            bool any = false;
            if (ROLLED && N > 2) {
#pragma unroll 1
                for (int round = 0; round < N - 1; ++round) {
                    any |= JacobiPairs<N, 0, 0>::run(C, Z);
                    jacobi_relabel<N>(C, Z);
                }
            } else {
                any = JacobiRounds<N, 0>::run(C, Z);
            }
            if (!__any_sync(0xffffffffu, any)) break;
        }
    }
    double k[N], kinv_all[N];
    bool bad = false;
#pragma unroll
    for (int j = 0; j < N; ++j) {
        double ksq = C[tri_idx(j, j)];
        if (!(ksq > 0.0)) {
            bad = true;
            ksq = fabs(ksq) + 1e-300;
        }
        // sqrt and its reciprocal from one Newton-refined rsqrt (the IEEE sqrt / division slow paths cost ~60
        // instructions per eigenvalue); k is corrected once more so that k * k == ksq to rounding
        const double ry = rsqrt_nr(ksq);
        double kk = ksq * ry;
        kk = fma(0.5 * ry, fma(-kk, kk, ksq), kk);
        k[j] = kk;
        kinv_all[j] = ry;
    }
    if (bad && valid) atomicOr(V.status, 2u);
    // X~ = H Z in place (H lower triangular: row i only needs rows <= i, so go bottom-up)
    {
        double H[TS];
#pragma unroll
        for (int e = 0; e < TS; ++e) H[e] = V.eigH[(size_t)e * P + pi];
#pragma unroll
        for (int j = 0; j < N; ++j) {
#pragma unroll
            for (int i = N - 1; i >= 0; --i) {
                double s = 0.0;
#pragma unroll
                for (int r = 0; r <= i; ++r) s = fma(H[tri_idx(i, r)], Z[r * N + j], s);
                Z[i * N + j] = s;
            }
        }
    }
    if (!valid) return;
    const int w = (int)(q / V.T.L), p = (int)(q % V.T.L);
    const size_t idx = ((size_t)w * V.M + ms) * V.T.L + p;
    const double od = V.lay_od[q];
    double* __restrict__ kth = V.kth + idx * 2 * N;
#pragma unroll
    for (int j = 0; j < N; ++j) {
        kth[j] = k[j];
        kth[N + j] = exp(-k[j] * od);
    }
    double Sp[TS];
#pragma unroll
    for (int e = 0; e < TS; ++e) Sp[e] = V.eigS[(size_t)e * P + pi];
    double dinv[N];
#pragma unroll
    for (int i = 0; i < N; ++i) dinv[i] = 0.5 * rsqrt_nr(V.T.wt[i] * V.T.mu[i]);
    double* __restrict__ Wp = V.Wp + idx * N * N;
    double* __restrict__ Wm = V.Wm + idx * N * N;
    constexpr int JB = (N >= 2) ? 2 : 1;  // columns per pass (16-byte stores along j)
#pragma unroll
    for (int j0 = 0; j0 < N; j0 += JB) {
        double kinv[JB];
#pragma unroll
        for (int jj = 0; jj < JB; ++jj) kinv[jj] = kinv_all[j0 + jj];
#pragma unroll
        for (int i = 0; i < N; ++i) {
            double wp[JB], wm[JB];
#pragma unroll
            for (int jj = 0; jj < JB; ++jj) {
                double xm = 0.0;
#pragma unroll
                for (int r = 0; r < N; ++r) xm = fma(Sp[tri_idx(i, r)], Z[r * N + j0 + jj], xm);
                xm *= kinv[jj];
                const double x = Z[i * N + j0 + jj];
                wp[jj] = (x + xm) * dinv[i];
                wm[jj] = (x - xm) * dinv[i];
            }
            if (JB == 2) {
                *reinterpret_cast<double2*>(Wp + i * N + j0) = make_double2(wp[0], wp[JB - 1]);
                *reinterpret_cast<double2*>(Wm + i * N + j0) = make_double2(wm[0], wm[JB - 1]);
            } else {
                Wp[i * N + j0] = wp[0];
                Wm[i * N + j0] = wm[0];
            }
        }
    }
}

}  // namespace disco
