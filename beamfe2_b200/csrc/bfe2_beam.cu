// bfe2_beam.cu -- "one large structure" path: a chain of n_elem Hermitian beam
// elements (nodes numbered along the chain), far too large for one thread block.
//
//   * K and M are block tridiagonal with 3x3 blocks (one block row per node).  Supports are eliminated in
//     place: a supported DOF keeps its row/column with K_ii = 1, M_ii = 0 and zero couplings, so u_i = 0 and
//     the DOF numbering stays the reference's (pos = 3*node + k, Structure.py:379-393).
//   * Static solve: the sequential band Cholesky of the reference-equivalent CPU route
//     (scipy.linalg.solveh_banded) is a dependency chain of length n.  Here the chain is cut into P
//     partitions separated by single nodes (SPIKE / substructuring): every partition is block-Cholesky
//     factorised independently, its two "spikes" T^-1 E_left, T^-1 E_right give the Schur complement on the
//     P-1 separator nodes (again block tridiagonal, solved sequentially, P-1 steps), then the interiors are
//     corrected.  All right-hand sides are processed in parallel (thread per (partition, rhs)).
//   * Modal: shift-invert (sigma = 0) subspace iteration, driven from beamfe2_b200/beam.py:
//     Z = K^-1 (M X); A = Z^T (M X), B = Z^T (M Z); small dense generalized eigenproblem (the Jacobi kernel
//     of bfe2.cu); X <- Z Q.  The tall-skinny contractions Z^T Y ([q x n][n x q]) and Z Q ([n x q][q x q])
//     are the only dense GEMM-shaped work of the whole project and run on the FP64 tensor cores
//     (mma.sync.m8n8k4.f64 -> DMMA in SASS).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../include/bfe2.h"
#include "bfe2_device.cuh"

using namespace bfe2;

int bfe2_fail(int code, const char* fmt, ...);      // bfe2.cu

#define BEAM_CUDA(call)                                                                            \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) return bfe2_fail(-100, "%s failed: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

struct bfe2_beam {
    long long N = 0;            // nodes
    long long n = 0;            // 3 N
    int P = 1;                  // partitions
    std::vector<int> sep;       // separator node of partitions p | p+1, size P-1
    // device storage (one allocation)
    void* blob = nullptr;
    double *Kd, *Kl, *Md, *Ml;  // [N][9] diagonal / sub-diagonal blocks (block (k+1,k) in Kl[k])
    double *Sinv, *G;           // [N][9] partition-local block factor
    double *Vl, *Vr;            // [N][9] spikes
    double *RSinv, *RG, *Rl;    // [P][9] reduced system factor, Rl = block (j, j-1)
    int* d_sep;                 // [P+1] sep[-1] = -1 ... sep[P-1] = N, stored shifted by one
    int* d_part;                // [N] partition of an interior node, -1 for separator nodes
    int* d_info;                // [1]
    unsigned char* d_fixed;     // [3N]
};

namespace {

// ---- 3x3 helpers (row-major) ------------------------------------------------------------------
__device__ __forceinline__ void m3_load(double a[9], const double* p) {
#pragma unroll
    for (int i = 0; i < 9; ++i) a[i] = p[i];
}
__device__ __forceinline__ void m3_store(double* p, const double a[9]) {
#pragma unroll
    for (int i = 0; i < 9; ++i) p[i] = a[i];
}
// c = a * b
__device__ __forceinline__ void m3_mul(double c[9], const double a[9], const double b[9]) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) c[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
}
// c = a * b^T
__device__ __forceinline__ void m3_mul_bt(double c[9], const double a[9], const double b[9]) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            c[3 * i + j] = a[3 * i] * b[3 * j] + a[3 * i + 1] * b[3 * j + 1] + a[3 * i + 2] * b[3 * j + 2];
}
// c = a^T * b
__device__ __forceinline__ void m3_mul_at(double c[9], const double a[9], const double b[9]) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) c[3 * i + j] = a[i] * b[j] + a[3 + i] * b[3 + j] + a[6 + i] * b[6 + j];
}
// y = a x ; y = a^T x
__device__ __forceinline__ void m3_vec(double y[3], const double a[9], const double x[3]) {
#pragma unroll
    for (int i = 0; i < 3; ++i) y[i] = a[3 * i] * x[0] + a[3 * i + 1] * x[1] + a[3 * i + 2] * x[2];
}
__device__ __forceinline__ void m3_vec_t(double y[3], const double a[9], const double x[3]) {
#pragma unroll
    for (int i = 0; i < 3; ++i) y[i] = a[i] * x[0] + a[3 + i] * x[1] + a[6 + i] * x[2];
}
// inverse of a symmetric positive definite 3x3 through its Cholesky factor; returns false if not SPD
__device__ __forceinline__ bool m3_spd_inv(double inv[9], const double a[9]) {
    const double l00s = a[0];
    if (!(l00s > 0.0)) return false;
    const double l00 = sqrt(l00s), i00 = 1.0 / l00;
    const double l10 = a[3] * i00, l20 = a[6] * i00;
    const double d1 = a[4] - l10 * l10;
    if (!(d1 > 0.0)) return false;
    const double l11 = sqrt(d1), i11 = 1.0 / l11;
    const double l21 = (a[7] - l20 * l10) * i11;
    const double d2 = a[8] - l20 * l20 - l21 * l21;
    if (!(d2 > 0.0)) return false;
    const double l22 = sqrt(d2), i22 = 1.0 / l22;
    // M = L^-1 (lower): m00 = i00, m10 = -l10 i00 i11, m11 = i11, m20 = -(l20 m00 + l21 m10) i22, ...
    const double m00 = i00, m10 = -l10 * m00 * i11, m11 = i11;
    const double m20 = -(l20 * m00 + l21 * m10) * i22, m21 = -l21 * m11 * i22, m22 = i22;
    // A^-1 = M^T M
    inv[0] = m00 * m00 + m10 * m10 + m20 * m20;
    inv[1] = inv[3] = m10 * m11 + m20 * m21;
    inv[2] = inv[6] = m20 * m22;
    inv[4] = m11 * m11 + m21 * m21;
    inv[5] = inv[7] = m21 * m22;
    inv[8] = m22 * m22;
    return true;
}

// general 3x3 inverse (adjugate); used for the separator system, whose Schur-complement blocks are positive
// definite only up to the cancellation error of the condensation (cond(K) ~ n^4)
__device__ __forceinline__ bool m3_gen_inv(double inv[9], const double a[9]) {
    const double c00 = a[4] * a[8] - a[5] * a[7], c01 = a[5] * a[6] - a[3] * a[8], c02 = a[3] * a[7] - a[4] * a[6];
    const double det = a[0] * c00 + a[1] * c01 + a[2] * c02;
    if (det == 0.0 || !isfinite(det)) return false;
    const double r = 1.0 / det;
    inv[0] = c00 * r; inv[1] = (a[2] * a[7] - a[1] * a[8]) * r; inv[2] = (a[1] * a[5] - a[2] * a[4]) * r;
    inv[3] = c01 * r; inv[4] = (a[0] * a[8] - a[2] * a[6]) * r; inv[5] = (a[2] * a[3] - a[0] * a[5]) * r;
    inv[6] = c02 * r; inv[7] = (a[1] * a[6] - a[0] * a[7]) * r; inv[8] = (a[0] * a[4] - a[1] * a[3]) * r;
    return true;
}

// ---- assembly: thread per node ------------------------------------------------------------------
__device__ __forceinline__ void elem_global(const double* coords, const double* props, long long e, bool mass,
                                            double X[6][6]) {
    double L, c, s;
    beam_geometry(coords[2 * e], coords[2 * e + 1], coords[2 * e + 2], coords[2 * e + 3], L, c, s);
    if (mass) me_local(X, props[4 * e + 3], props[4 * e + 1], L, false);
    else ke_local(X, props[4 * e], props[4 * e + 1], props[4 * e + 2], L);
    rotate_to_global(X, c, s);
}

__global__ void k_beam_assemble(long long N, const double* __restrict__ coords, const double* __restrict__ props,
                                const unsigned char* __restrict__ fixed, double* Kd, double* Kl, double* Md, double* Ml) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= N) return;
    for (int pass = 0; pass < 2; ++pass) {
        const bool mass = pass == 1;
        double D[9], Lb[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) { D[i] = 0.0; Lb[i] = 0.0; }
        double X[6][6];
        if (k > 0) {                                   // element k-1 first: beam-list order (helpers.py:19-39)
            elem_global(coords, props, k - 1, mass, X);
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int b = 0; b < 3; ++b) D[3 * a + b] += X[3 + a][3 + b];
        }
        if (k < N - 1) {
            elem_global(coords, props, k, mass, X);
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int b = 0; b < 3; ++b) {
                    D[3 * a + b] += X[a][b];
                    Lb[3 * a + b] = X[3 + a][b];       // block (k+1, k)
                }
        }
        // supports: decouple the DOF, K_ii = 1, M_ii = 0
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const bool fa = fixed[3 * k + a] != 0;
#pragma unroll
            for (int b = 0; b < 3; ++b) {
                const bool fb = fixed[3 * k + b] != 0;
                if (fa || fb) D[3 * a + b] = (a == b && !mass) ? 1.0 : 0.0;
                if (k < N - 1) {
                    const bool fr = fixed[3 * (k + 1) + a] != 0;     // row DOF belongs to node k+1
                    if (fr || fb) Lb[3 * a + b] = 0.0;
                }
            }
        }
        m3_store((mass ? Md : Kd) + 9 * k, D);
        m3_store((mass ? Ml : Kl) + 9 * k, Lb);
    }
}

// ---- partition factorisation: thread per partition ---------------------------------------------
__global__ void k_beam_factor(int P, const int* __restrict__ sep, const double* __restrict__ Kd,
                              const double* __restrict__ Kl, double* Sinv, double* G, int* info) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const int a = sep[p] + 1, b = sep[p + 1] - 1;       // interior nodes a..b
    double S[9], Si[9], Lk[9], Gk[9], T[9];
    m3_load(S, Kd + 9 * (long long)a);
    for (int k = a; k <= b; ++k) {
        if (!m3_spd_inv(Si, S)) { atomicCAS(info, 0, k + 1); return; }
        m3_store(Sinv + 9 * (long long)k, Si);
        if (k < b) {
            m3_load(Lk, Kl + 9 * (long long)k);
            m3_mul(Gk, Lk, Si);                        // G_k = L_k S_k^-1
            m3_store(G + 9 * (long long)k, Gk);
            m3_mul_bt(T, Gk, Lk);                      // G_k L_k^T
            m3_load(S, Kd + 9 * (long long)(k + 1));
#pragma unroll
            for (int i = 0; i < 9; ++i) S[i] -= T[i];
        }
    }
}

// solve T y = w in place for one 3-vector column stored with `stride` doubles between DOFs of a node and
// `nstride` between nodes:  y(node k, dof d) at base[k*nstride + d*stride]
__device__ __forceinline__ void part_solve_inplace(double* base, long long nstride, long long stride, int a, int b,
                                                   const double* __restrict__ Sinv, const double* __restrict__ G) {
    double w[3], t[3], g[9];
    // forward: w_k = f_k - G_{k-1} w_{k-1}
    w[0] = base[a * nstride]; w[1] = base[a * nstride + stride]; w[2] = base[a * nstride + 2 * stride];
    for (int k = a + 1; k <= b; ++k) {
        m3_load(g, G + 9 * (long long)(k - 1));
        m3_vec(t, g, w);
        double* pk = base + k * nstride;
        w[0] = pk[0] - t[0]; w[1] = pk[stride] - t[1]; w[2] = pk[2 * stride] - t[2];
        pk[0] = w[0]; pk[stride] = w[1]; pk[2 * stride] = w[2];
    }
    // diagonal + backward: y_k = S_k^-1 w_k - G_k^T y_{k+1}
    double y[3] = {0.0, 0.0, 0.0};
    for (int k = b; k >= a; --k) {
        double* pk = base + k * nstride;
        w[0] = pk[0]; w[1] = pk[stride]; w[2] = pk[2 * stride];
        m3_load(g, Sinv + 9 * (long long)k);
        double v[3];
        m3_vec(v, g, w);
        if (k < b) {
            m3_load(g, G + 9 * (long long)k);
            m3_vec_t(t, g, y);
            v[0] -= t[0]; v[1] -= t[1]; v[2] -= t[2];
        }
        y[0] = v[0]; y[1] = v[1]; y[2] = v[2];
        pk[0] = y[0]; pk[stride] = y[1]; pk[2 * stride] = y[2];
    }
}

// spikes: thread per (partition, column c in 0..5); Vl/Vr hold 3x3 per node, column c of the spike
__global__ void k_beam_spikes(int P, long long N, const int* __restrict__ sep, const double* __restrict__ Kl,
                              const double* __restrict__ Sinv, const double* __restrict__ G, double* Vl, double* Vr) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= P * 6) return;
    const int p = t / 6, c = t % 6;
    const int sl = sep[p], sr = sep[p + 1];
    const int a = sl + 1, b = sr - 1;
    double* V = (c < 3) ? Vl : Vr;
    const int col = c % 3;
    for (int k = a; k <= b; ++k)
        for (int d = 0; d < 3; ++d) V[9 * (long long)k + 3 * d + col] = 0.0;
    if (c < 3) {
        if (sl < 0) return;                            // no left separator: spike is zero
        // E_l: block at first interior node a = L[sl] (block (a, sl)); column col
        for (int d = 0; d < 3; ++d) V[9 * (long long)a + 3 * d + col] = Kl[9 * (long long)sl + 3 * d + col];
    } else {
        if (sr >= N) return;
        // E_r: block at last interior node b = L[b]^T (L[b] = block (sr, b)); column col of L^T = row col of L
        for (int d = 0; d < 3; ++d) V[9 * (long long)b + 3 * d + col] = Kl[9 * (long long)b + 3 * col + d];
    }
    part_solve_inplace(V + col, 9, 3, a, b, Sinv, G);
}

// reduced (separator) system: build + sequential block factorisation, single thread (P-1 <= a few thousand)
__global__ void k_beam_reduced_factor(int P, const int* __restrict__ sep, const double* __restrict__ Kd,
                                      const double* __restrict__ Kl, const double* __restrict__ Vl,
                                      const double* __restrict__ Vr, double* RSinv, double* RG, double* Rl, int* info) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    double S[9], Si[9], L1[9], L2[9], T[9], A[9], Gp[9];
    for (int j = 0; j < P - 1; ++j) {
        const long long s = sep[j + 1];
        m3_load(S, Kd + 9 * s);
        m3_load(L1, Kl + 9 * (s - 1));                 // block (s, s-1)
        m3_load(L2, Kl + 9 * s);                       // block (s+1, s)
        m3_load(A, Vr + 9 * (s - 1));
        m3_mul(T, L1, A);                              // L[s-1] Vr[s-1]
#pragma unroll
        for (int i = 0; i < 9; ++i) S[i] -= T[i];
        m3_load(A, Vl + 9 * (s + 1));
        m3_mul_at(T, L2, A);                           // L[s]^T Vl[s+1]
#pragma unroll
        for (int i = 0; i < 9; ++i) S[i] -= T[i];
        double Rlj[9];
        m3_load(A, Vl + 9 * (s - 1));
        m3_mul(Rlj, L1, A);                            // block (j, j-1) = -L[s-1] Vl[s-1]
#pragma unroll
        for (int i = 0; i < 9; ++i) Rlj[i] = -Rlj[i];
        m3_store(Rl + 9 * j, Rlj);
        if (j > 0) {                                   // S_j -= Rl_j RSinv_{j-1} Rl_j^T
            m3_load(Si, RSinv + 9 * (j - 1));
            m3_mul(Gp, Rlj, Si);                       // RG_{j-1} = Rl_j RSinv_{j-1}
            m3_store(RG + 9 * (j - 1), Gp);
            m3_mul_bt(T, Gp, Rlj);
#pragma unroll
            for (int i = 0; i < 9; ++i) S[i] -= T[i];
        }
        // symmetrise (the two spike routes agree only to rounding)
        S[1] = S[3] = 0.5 * (S[1] + S[3]); S[2] = S[6] = 0.5 * (S[2] + S[6]); S[5] = S[7] = 0.5 * (S[5] + S[7]);
        if (!m3_spd_inv(Si, S) && !m3_gen_inv(Si, S)) { atomicCAS(info, 0, (int)s + 1); return; }
        m3_store(RSinv + 9 * j, Si);
    }
}

// ---- solve, stage 1: thread per (partition, rhs): y = T^-1 f on the interiors (in place in U) ---
__global__ void k_beam_solve_part(int P, int nrhs, const int* __restrict__ sep, const double* __restrict__ Sinv,
                                  const double* __restrict__ G, double* U) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)P * nrhs) return;
    const int p = (int)(t / nrhs), r = (int)(t % nrhs);
    const int a = sep[p] + 1, b = sep[p + 1] - 1;
    part_solve_inplace(U + r, 3LL * nrhs, nrhs, a, b, Sinv, G);
}

// stage 2: thread per rhs: reduced right-hand side, sequential separator solve, x_s into U
__global__ void k_beam_solve_reduced(int P, int nrhs, const int* __restrict__ sep, const double* __restrict__ Kl,
                                     const double* __restrict__ RSinv, const double* __restrict__ RG, double* U) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrhs || P < 2) return;
    const long long ns = 3LL * nrhs;
    double g[3], t[3], A[9], w[3] = {0, 0, 0};
    // forward over separators (reduced rhs built on the fly, stored in place at the separator rows)
    for (int j = 0; j < P - 1; ++j) {
        const long long s = sep[j + 1];
        double* ps = U + s * ns + r;
        g[0] = ps[0]; g[1] = ps[nrhs]; g[2] = ps[2 * nrhs];
        double y[3];
        const double* pm = U + (s - 1) * ns + r;
        y[0] = pm[0]; y[1] = pm[nrhs]; y[2] = pm[2 * nrhs];
        m3_load(A, Kl + 9 * (s - 1));
        m3_vec(t, A, y);                               // L[s-1] y[s-1]
        g[0] -= t[0]; g[1] -= t[1]; g[2] -= t[2];
        const double* pp = U + (s + 1) * ns + r;
        y[0] = pp[0]; y[1] = pp[nrhs]; y[2] = pp[2 * nrhs];
        m3_load(A, Kl + 9 * s);
        m3_vec_t(t, A, y);                             // L[s]^T y[s+1]
        g[0] -= t[0]; g[1] -= t[1]; g[2] -= t[2];
        if (j > 0) {
            m3_load(A, RG + 9 * (j - 1));
            m3_vec(t, A, w);                           // RG_{j-1} w_{j-1}
            g[0] -= t[0]; g[1] -= t[1]; g[2] -= t[2];
        }
        w[0] = g[0]; w[1] = g[1]; w[2] = g[2];
        ps[0] = w[0]; ps[nrhs] = w[1]; ps[2 * nrhs] = w[2];
    }
    double x[3] = {0, 0, 0};
    for (int j = P - 2; j >= 0; --j) {
        const long long s = sep[j + 1];
        double* ps = U + s * ns + r;
        w[0] = ps[0]; w[1] = ps[nrhs]; w[2] = ps[2 * nrhs];
        double v[3];
        m3_load(A, RSinv + 9 * j);
        m3_vec(v, A, w);
        if (j < P - 2) {
            m3_load(A, RG + 9 * j);
            m3_vec_t(t, A, x);
            v[0] -= t[0]; v[1] -= t[1]; v[2] -= t[2];
        }
        x[0] = v[0]; x[1] = v[1]; x[2] = v[2];
        ps[0] = x[0]; ps[nrhs] = x[1]; ps[2 * nrhs] = x[2];
    }
}

// stage 3: thread per (interior node, rhs): x = y - Vl x_left - Vr x_right
__global__ void k_beam_solve_fix(long long N, int P, int nrhs, const int* __restrict__ sep, const int* __restrict__ part_of,
                                 const double* __restrict__ Vl, const double* __restrict__ Vr, double* U) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= N * nrhs) return;
    const long long k = t / nrhs;
    const int r = (int)(t % nrhs);
    const int p = part_of[k];
    if (p < 0) return;                                 // separator node
    const long long ns = 3LL * nrhs;
    const long long sl = sep[p], sr = sep[p + 1];
    double* pk = U + k * ns + r;
    double y[3] = {pk[0], pk[nrhs], pk[2 * nrhs]};
    double A[9], x[3], tt[3];
    if (sl >= 0) {
        const double* ps = U + sl * ns + r;
        x[0] = ps[0]; x[1] = ps[nrhs]; x[2] = ps[2 * nrhs];
        m3_load(A, Vl + 9 * k);
        m3_vec(tt, A, x);
        y[0] -= tt[0]; y[1] -= tt[1]; y[2] -= tt[2];
    }
    if (sr < N) {
        const double* ps = U + sr * ns + r;
        x[0] = ps[0]; x[1] = ps[nrhs]; x[2] = ps[2 * nrhs];
        m3_load(A, Vr + 9 * k);
        m3_vec(tt, A, x);
        y[0] -= tt[0]; y[1] -= tt[1]; y[2] -= tt[2];
    }
    pk[0] = y[0]; pk[nrhs] = y[1]; pk[2 * nrhs] = y[2];
}

// ---- block tridiagonal matvec: thread per (node, rhs) -------------------------------------------
__global__ void k_beam_matvec(long long N, int nrhs, const double* __restrict__ D, const double* __restrict__ Lb,
                              const double* __restrict__ X, double* __restrict__ Y) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= N * nrhs) return;
    const long long k = t / nrhs;
    const int r = (int)(t % nrhs);
    const long long ns = 3LL * nrhs;
    double A[9], x[3], y[3], tt[3];
    const double* pk = X + k * ns + r;
    x[0] = pk[0]; x[1] = pk[nrhs]; x[2] = pk[2 * nrhs];
    m3_load(A, D + 9 * k);
    m3_vec(y, A, x);
    if (k > 0) {
        const double* pm = X + (k - 1) * ns + r;
        x[0] = pm[0]; x[1] = pm[nrhs]; x[2] = pm[2 * nrhs];
        m3_load(A, Lb + 9 * (k - 1));
        m3_vec(tt, A, x);
        y[0] += tt[0]; y[1] += tt[1]; y[2] += tt[2];
    }
    if (k < N - 1) {
        const double* pp = X + (k + 1) * ns + r;
        x[0] = pp[0]; x[1] = pp[nrhs]; x[2] = pp[2 * nrhs];
        m3_load(A, Lb + 9 * k);
        m3_vec_t(tt, A, x);
        y[0] += tt[0]; y[1] += tt[1]; y[2] += tt[2];
    }
    double* po = Y + k * ns + r;
    po[0] = y[0]; po[nrhs] = y[1]; po[2 * nrhs] = y[2];
}

// ---- FP64 tensor-core contractions (q = 64) -----------------------------------------------------
__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

constexpr int GQ = 64;            // block width of the subspace
constexpr int GS = 68;            // shared-memory row pitch (68 mod 16 = 4: conflict-free fragment loads)
constexpr int GCH = 32;           // rows of X, Y staged per chunk

// partial Gram: C_part[block] = X[rows]^T Y[rows]; 8 warps, warp w owns rows 8w..8w+7 of C (all 64 columns)
__global__ void __launch_bounds__(256) k_gram_partial(long long n, const double* __restrict__ X,
                                                      const double* __restrict__ Y, double* __restrict__ Cpart,
                                                      long long rows_per_block) {
    __shared__ double Xs[GCH * GS], Ys[GCH * GS];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int gid = lane >> 2, tig = lane & 3;
    double acc[8][2];
#pragma unroll
    for (int t = 0; t < 8; ++t) acc[t][0] = acc[t][1] = 0.0;
    const long long r0 = (long long)blockIdx.x * rows_per_block;
    const long long r1 = min(n, r0 + rows_per_block);
    for (long long base = r0; base < r1; base += GCH) {
        __syncthreads();
        for (int i = tid; i < GCH * GQ; i += 256) {
            const int rr = i / GQ, cc = i % GQ;
            const long long row = base + rr;
            const bool ok = row < r1;
            Xs[rr * GS + cc] = ok ? X[row * GQ + cc] : 0.0;
            Ys[rr * GS + cc] = ok ? Y[row * GQ + cc] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int k0 = 0; k0 < GCH; k0 += 4) {
            const double a = Xs[(k0 + tig) * GS + 8 * w + gid];        // A[m = gid][k = tig] = X[k][8w + m]
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                const double b = Ys[(k0 + tig) * GS + 8 * t + gid];    // B[k = tig][n = gid] = Y[k][8t + n]
                dmma_m8n8k4(acc[t][0], acc[t][1], a, b);
            }
        }
    }
    double* out = Cpart + (long long)blockIdx.x * GQ * GQ;
#pragma unroll
    for (int t = 0; t < 8; ++t) {                                       // C[row gid][cols 2 tig, 2 tig + 1]
        out[(8 * w + gid) * GQ + 8 * t + 2 * tig] = acc[t][0];
        out[(8 * w + gid) * GQ + 8 * t + 2 * tig + 1] = acc[t][1];
    }
}

__global__ void k_gram_reduce(int nparts, const double* __restrict__ Cpart, double* __restrict__ C) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= GQ * GQ) return;
    double s = 0.0;
    for (int p = 0; p < nparts; ++p) s += Cpart[(long long)p * GQ * GQ + i];   // fixed order: deterministic
    C[i] = s;
}

// Xout[rows] = Z[rows] * Q  (Q is 64 x 64 row-major); block = 64 rows, 8 warps x 8 rows
__global__ void __launch_bounds__(256) k_update(long long n, const double* __restrict__ Z, const double* __restrict__ Q,
                                                double* __restrict__ Xout) {
    extern __shared__ __align__(16) double upd_smem[];
    double* Zs = upd_smem;
    double* Qs = upd_smem + GQ * GS;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int gid = lane >> 2, tig = lane & 3;
    const long long base = (long long)blockIdx.x * GQ;
    for (int i = tid; i < GQ * GQ; i += 256) {
        const int rr = i / GQ, cc = i % GQ;
        const long long row = base + rr;
        Zs[rr * GS + cc] = row < n ? Z[row * GQ + cc] : 0.0;
        Qs[rr * GS + cc] = Q[i];
    }
    __syncthreads();
    double acc[8][2];
#pragma unroll
    for (int t = 0; t < 8; ++t) acc[t][0] = acc[t][1] = 0.0;
#pragma unroll
    for (int k0 = 0; k0 < GQ; k0 += 4) {
        const double a = Zs[(8 * w + gid) * GS + k0 + tig];             // A[m = gid][k = tig] = Z[8w + m][k]
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            const double b = Qs[(k0 + tig) * GS + 8 * t + gid];         // B[k][n] = Q[k][8t + n]
            dmma_m8n8k4(acc[t][0], acc[t][1], a, b);
        }
    }
    const long long row = base + 8 * w + gid;
    if (row < n) {
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            Xout[row * GQ + 8 * t + 2 * tig] = acc[t][0];
            Xout[row * GQ + 8 * t + 2 * tig + 1] = acc[t][1];
        }
    }
}

}  // namespace

// =================================================================================================
extern "C" {

int bfe2_beam_create(bfe2_beam** out, int64_t n_elem, const double* coords, const double* props,
                     const uint8_t* fixed_mask, int32_t n_partitions, void* stream) {
    if (!out || n_elem < 1 || !coords || !props || !fixed_mask) return bfe2_fail(-1, "bfe2_beam_create: bad argument");
    *out = nullptr;
    cudaStream_t st = (cudaStream_t)stream;
    bfe2_beam* h = new bfe2_beam();
    h->N = n_elem + 1;
    h->n = 3 * h->N;
    int P = n_partitions > 0 ? n_partitions : (int)std::min<long long>(8192, std::max<long long>(1, h->N / 32));
    while (P > 1 && h->N < 3LL * P) --P;               // every partition needs at least one interior node
    h->P = P;
    std::vector<int> sep(P + 1);
    sep[0] = -1;
    sep[P] = (int)h->N;
    for (int j = 1; j < P; ++j) sep[j] = (int)((long long)j * h->N / P);
    std::vector<int> part_of(h->N, 0);
    for (int p = 0; p < P; ++p) {
        for (int k = sep[p] + 1; k < sep[p + 1]; ++k) part_of[k] = p;
        if (p + 1 < P) part_of[sep[p + 1]] = -1;
    }
    const size_t nb = (size_t)h->N * 9 * sizeof(double);
    const size_t pb = (size_t)(P + 1) * 9 * sizeof(double);
    auto r64 = [](size_t b) { return (b + 63) & ~size_t(63); };
    const size_t total = 8 * r64(nb) + 3 * r64(pb) + r64((size_t)(P + 1) * sizeof(int)) +
                         r64((size_t)h->N * sizeof(int)) + r64(sizeof(int)) + r64((size_t)h->n) + 256;
    if (cudaMalloc(&h->blob, total) != cudaSuccess) {
        delete h;
        return bfe2_fail(-100, "cudaMalloc of %zu bytes failed", total);
    }
    char* q = (char*)h->blob;
    auto take = [&](size_t b) { char* r = q; q += (b + 63) & ~size_t(63); return r; };
    h->Kd = (double*)take(nb); h->Kl = (double*)take(nb); h->Md = (double*)take(nb); h->Ml = (double*)take(nb);
    h->Sinv = (double*)take(nb); h->G = (double*)take(nb); h->Vl = (double*)take(nb); h->Vr = (double*)take(nb);
    h->RSinv = (double*)take(pb); h->RG = (double*)take(pb); h->Rl = (double*)take(pb);
    h->d_sep = (int*)take((P + 1) * sizeof(int));
    h->d_part = (int*)take(h->N * sizeof(int));
    int* d_part = h->d_part;
    h->d_info = (int*)take(sizeof(int));
    h->d_fixed = (unsigned char*)take(h->n);
    BEAM_CUDA(cudaMemcpyAsync(h->d_sep, sep.data(), (P + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
    BEAM_CUDA(cudaMemcpyAsync(d_part, part_of.data(), h->N * sizeof(int), cudaMemcpyHostToDevice, st));
    BEAM_CUDA(cudaMemcpyAsync(h->d_fixed, fixed_mask, h->n, cudaMemcpyHostToDevice, st));
    BEAM_CUDA(cudaMemsetAsync(h->d_info, 0, sizeof(int), st));
    BEAM_CUDA(cudaStreamSynchronize(st));              // host vectors go out of scope
    h->sep.assign(sep.begin(), sep.end());
    const int T = 128;
    k_beam_assemble<<<(unsigned)((h->N + T - 1) / T), T, 0, st>>>(h->N, coords, props, h->d_fixed, h->Kd, h->Kl, h->Md, h->Ml);
    k_beam_factor<<<(P + T - 1) / T, T, 0, st>>>(P, h->d_sep, h->Kd, h->Kl, h->Sinv, h->G, h->d_info);
    k_beam_spikes<<<(P * 6 + T - 1) / T, T, 0, st>>>(P, h->N, h->d_sep, h->Kl, h->Sinv, h->G, h->Vl, h->Vr);
    k_beam_reduced_factor<<<1, 32, 0, st>>>(P, h->d_sep, h->Kd, h->Kl, h->Vl, h->Vr, h->RSinv, h->RG, h->Rl, h->d_info);
    BEAM_CUDA(cudaGetLastError());
    int info = 0;
    BEAM_CUDA(cudaMemcpyAsync(&info, h->d_info, sizeof(int), cudaMemcpyDeviceToHost, st));
    BEAM_CUDA(cudaStreamSynchronize(st));
    if (info != 0) {
        cudaFree(h->blob);
        delete h;
        return bfe2_fail(-4, "stiffness matrix is not positive definite (node %d)", info);
    }
    *out = h;
    return 0;
}

void bfe2_beam_destroy(bfe2_beam* h) {
    if (!h) return;
    if (h->blob) cudaFree(h->blob);
    delete h;
}

int bfe2_beam_dims(const bfe2_beam* h, int64_t* n_nodes, int64_t* n_dof, int32_t* n_partitions) {
    if (!h) return bfe2_fail(-1, "beam handle is NULL");
    if (n_nodes) *n_nodes = h->N;
    if (n_dof) *n_dof = h->n;
    if (n_partitions) *n_partitions = h->P;
    return 0;
}

int bfe2_beam_solve(bfe2_beam* h, int32_t nrhs, const double* F, double* U, void* stream) {
    if (!h || nrhs < 1 || !F || !U) return bfe2_fail(-1, "bfe2_beam_solve: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    if (F != U) BEAM_CUDA(cudaMemcpyAsync(U, F, (size_t)h->n * nrhs * sizeof(double), cudaMemcpyDeviceToDevice, st));
    const int T = 128;
    const int* d_part = h->d_part;
    const long long t1 = (long long)h->P * nrhs;
    k_beam_solve_part<<<(unsigned)((t1 + T - 1) / T), T, 0, st>>>(h->P, nrhs, h->d_sep, h->Sinv, h->G, U);
    if (h->P > 1) {
        k_beam_solve_reduced<<<(nrhs + 31) / 32, 32, 0, st>>>(h->P, nrhs, h->d_sep, h->Kl, h->RSinv, h->RG, U);
        const long long t3 = h->N * nrhs;
        k_beam_solve_fix<<<(unsigned)((t3 + T - 1) / T), T, 0, st>>>(h->N, h->P, nrhs, h->d_sep, d_part, h->Vl, h->Vr, U);
    }
    BEAM_CUDA(cudaGetLastError());
    return 0;
}

int bfe2_beam_matvec(bfe2_beam* h, int32_t which, int32_t nrhs, const double* X, double* Y, void* stream) {
    if (!h || nrhs < 1 || !X || !Y || X == Y || which < 0 || which > 1) return bfe2_fail(-1, "bfe2_beam_matvec: bad argument");
    const int T = 128;
    const long long t = h->N * nrhs;
    k_beam_matvec<<<(unsigned)((t + T - 1) / T), T, 0, (cudaStream_t)stream>>>(h->N, nrhs, which ? h->Md : h->Kd,
                                                                              which ? h->Ml : h->Kl, X, Y);
    BEAM_CUDA(cudaGetLastError());
    return 0;
}

size_t bfe2_gram_workspace_bytes(int64_t n) {
    const long long blocks = std::min<long long>(592, std::max<long long>(1, (n + GCH - 1) / GCH));
    return (size_t)blocks * GQ * GQ * sizeof(double);
}

int bfe2_gram64(int64_t n, const double* X, const double* Y, double* C, double* workspace, void* stream) {
    if (n < 1 || !X || !Y || !C || !workspace) return bfe2_fail(-1, "bfe2_gram64: bad argument");
    long long blocks = std::min<long long>(592, std::max<long long>(1, (n + GCH - 1) / GCH));
    long long rpb = ((n + blocks - 1) / blocks + GCH - 1) / GCH * GCH;
    blocks = (n + rpb - 1) / rpb;
    k_gram_partial<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(n, X, Y, workspace, rpb);
    k_gram_reduce<<<(GQ * GQ + 255) / 256, 256, 0, (cudaStream_t)stream>>>((int)blocks, workspace, C);
    BEAM_CUDA(cudaGetLastError());
    return 0;
}

int bfe2_update64(int64_t n, const double* Z, const double* Q, double* Xout, void* stream) {
    if (n < 1 || !Z || !Q || !Xout || Z == Xout) return bfe2_fail(-1, "bfe2_update64: bad argument");
    const size_t smem = 2 * GQ * GS * sizeof(double);
    BEAM_CUDA(cudaFuncSetAttribute(k_update, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_update<<<(unsigned)((n + GQ - 1) / GQ), 256, smem, (cudaStream_t)stream>>>(n, Z, Q, Xout);
    BEAM_CUDA(cudaGetLastError());
    return 0;
}

}  // extern "C"
