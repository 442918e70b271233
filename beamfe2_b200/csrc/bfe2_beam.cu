// bfe2_beam.cu -- "one large structure" path: a chain of n_elem Hermitian beam
// elements (nodes numbered along the chain), far too large for one thread block.
//
//   * K and M are block tridiagonal with 3x3 blocks (one block row per node).  Supports are eliminated in
//     place: a supported DOF keeps its row/column with K_ii = 1, M_ii = 0 and zero couplings, so u_i = 0 and
//     the DOF numbering stays the reference's (pos = 3*node + k, Structure.py:379-393).
//   * Precision.  cond(K) ~ 16 n^4 (1.6e21 at n = 1e5): assembled in FP64 the problem is no longer the
//     user's problem -- every FP64 solver, scipy included, is ~100 % off on the tip deflection and the first
//     frequencies.  The inputs are exact FP64 numbers, so by default the element matrices, the partitioned
//     factorisation and the substitutions run in DOUBLE-DOUBLE (bfe2_dd.cuh; eps * cond ~ 2e-11).  The same
//     kernels instantiated for plain FP64 (BFE2_BEAM_FP64) are kept for comparison.
//   * Static solve: the sequential band Cholesky of the reference-equivalent CPU route
//     (scipy.linalg.solveh_banded) is a dependency chain of length n.  Here the chain is cut into P
//     partitions separated by single nodes (SPIKE / substructuring): every partition is block-Cholesky
//     factorised independently, its two "spikes" T^-1 E_left, T^-1 E_right give the Schur complement on the
//     P-1 separator nodes (again block tridiagonal), then the interiors are corrected.  The separator system is
//     either eliminated sequentially (one level, P-1 steps) or -- two levels, automatic for long chains -- cut into
//     partitions once more, so that a solve walks three chains of ~cbrt(N) steps instead of two of ~sqrt(N).
//     All right-hand sides are processed in parallel (thread per (partition, rhs)).
//   * Modal: shift-invert (sigma = 0) subspace iteration, driven from beamfe2_b200/beam.py:
//     Z = K^-1 (M X); A = Z^T (M X), B = Z^T (M Z); small dense generalized eigenproblem (the Jacobi kernel
//     of bfe2.cu); X <- Z Q.  The tall-skinny contractions Z^T Y ([q x n][n x q]) and Z Q ([n x q][q x q])
//     are the only dense GEMM-shaped work of the whole project and run on the FP64 tensor cores
//     (mma.sync.m8n8k4.f64 -> DMMA in SASS).
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../include/bfe2.h"
#include "bfe2_beam_core.cuh"

using namespace bfe2;
using namespace bfe2::chain;

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
    int prec = BFE2_BEAM_DD;    // scalar type of the stored matrices / factor
    // device storage (one allocation); element type double or dd according to prec
    void* blob = nullptr;
    void *Kd, *Kl, *Md, *Ml;    // [N][9] diagonal / sub-diagonal blocks (block (k+1,k) in Kl[k])
    void *Sinv, *G;             // [N][9] partition-local block factor
    void *Vl, *Vr;              // [N][9] spikes
    void *RSinv, *RG, *Rl;      // [P][9] sequentially eliminated separator system of the LAST level, Rl = block (j, j-1)
    int* d_sep;                 // [P+1] sep[-1] = -1 ... sep[P-1] = N, stored shifted by one
    int* d_part;                // [N] partition of an interior node, -1 for separator nodes
    // second level (P2 > 0): the P - 1 separators of level 1 form a chain of their own, cut into P2 partitions
    int P2 = 0;
    long long N2 = 0;
    void *Kd2 = nullptr, *Kl2 = nullptr, *Sinv2 = nullptr, *G2 = nullptr, *Vl2 = nullptr, *Vr2 = nullptr;   // [N2][9]
    int *d_sep2 = nullptr, *d_part2 = nullptr;
    double* u2_ws = nullptr;    // [2][3 N2 * u2_cap] right-hand sides of the separator system (high, low words)
    long long u2_cap = 0;
    int* d_info;                // [1]
    unsigned char* d_fixed;     // [3N]
    double* lo_ws = nullptr;    // [n * lo_cap] low words of the running solution (dd only)
    long long lo_cap = 0;
    ~bfe2_beam() {
        if (blob) cudaFree(blob);
        if (lo_ws) cudaFree(lo_ws);
        if (u2_ws) cudaFree(u2_ws);
    }
};

namespace {

// ---- assembly: thread per node ------------------------------------------------------------------
template <class S>
__global__ void k_beam_assemble(long long N, const double* __restrict__ coords, const double* __restrict__ props,
                                const unsigned char* __restrict__ fixed, S* Kd, S* Kl, S* Md, S* Ml) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= N) return;
    assemble_node<S>(k, N, coords, props, fixed, false, Kd, Kl);
    assemble_node<S>(k, N, coords, props, fixed, true, Md, Ml);
}

// ---- partition factorisation: thread per partition ---------------------------------------------
template <class S>
__global__ void k_beam_factor(int P, const int* __restrict__ sep, const S* __restrict__ Kd,
                              const S* __restrict__ Kl, S* Sinv, S* G, int* info) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const int bad = factor_partition<S>(sep[p] + 1, sep[p + 1] - 1, Kd, Kl, Sinv, G);
    if (bad) atomicCAS(info, 0, bad);
}

// spikes: thread per (partition, column c in 0..5)
template <class S>
__global__ void k_beam_spikes(int P, long long N, const int* __restrict__ sep, const S* __restrict__ Kl,
                              const S* __restrict__ Sinv, const S* __restrict__ G, S* Vl, S* Vr) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= P * 6) return;
    spike_column<S>(t / 6, t % 6, N, sep, Kl, Sinv, G, Vl, Vr);
}

// reduced (separator) system: build + sequential block factorisation, single thread (P-1 <= a few thousand)
template <class S>
__global__ void k_beam_reduced_factor(int P, const int* __restrict__ sep, const S* __restrict__ Kd,
                                      const S* __restrict__ Kl, const S* __restrict__ Vl,
                                      const S* __restrict__ Vr, S* RSinv, S* RG, S* Rl, int* info) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    const int bad = reduced_factor<S>(P, sep, Kd, Kl, Vl, Vr, RSinv, RG, Rl);
    if (bad) atomicCAS(info, 0, bad);
}

// ---- solve, stage 0: U <- F (+ low words) -------------------------------------------------------
__global__ void k_beam_load_rhs(long long count, const double* __restrict__ F, const double* __restrict__ Flo,
                                double* U, double* Ulo) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    if (U != F) U[t] = F[t];
    if (Ulo) Ulo[t] = Flo ? Flo[t] : 0.0;
}

// ---- solve, stage 1: thread per (partition, rhs): y = T^-1 f on the interiors (in place in U) ---
template <class S>
__global__ void k_beam_solve_part(int P, const int* __restrict__ sep, const S* __restrict__ Sinv,
                                  const S* __restrict__ G, MVec<S> U) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)P * U.nrhs) return;
    const int p = (int)(t / U.nrhs), r = (int)(t % U.nrhs);
    part_solve_inplace<S>(U, r, 3 * U.nrhs, U.nrhs, sep[p] + 1, sep[p + 1] - 1, Sinv, G);
}

// stage 2: thread per rhs: reduced right-hand side, sequential separator solve, x_s into U
template <class S>
__global__ void k_beam_solve_reduced(int P, const int* __restrict__ sep, const S* __restrict__ Kl,
                                     const S* __restrict__ RSinv, const S* __restrict__ RG, MVec<S> U) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= U.nrhs || P < 2) return;
    solve_reduced_rhs<S>(P, r, sep, Kl, RSinv, RG, U);
}

// stage 3: thread per (interior node, rhs): x = y - Vl x_left - Vr x_right
template <class S>
__global__ void k_beam_solve_fix(long long N, const int* __restrict__ sep, const int* __restrict__ part_of,
                                 const S* __restrict__ Vl, const S* __restrict__ Vr, MVec<S> U) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= N * U.nrhs) return;
    const long long k = t / U.nrhs;
    const int p = part_of[k];
    if (p < 0) return;                                 // separator node
    solve_fix_node<S>(k, (int)(t % U.nrhs), p, N, sep, Vl, Vr, U);
}

// ---- two-level scheme: separator system as a chain of its own (thread per separator / per (separator, rhs)) ------
template <class S>
__global__ void k_beam_reduced_build(int N2, const int* __restrict__ sep, const S* __restrict__ Kd,
                                     const S* __restrict__ Kl, const S* __restrict__ Vl, const S* __restrict__ Vr,
                                     S* Kd2, S* Kl2) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= N2) return;
    reduced_build_node<S>(j, sep, Kd, Kl, Vl, Vr, Kd2, Kl2);
}

template <class S>
__global__ void k_beam_reduced_rhs(int N2, const int* __restrict__ sep, const S* __restrict__ Kl, MVec<S> U, MVec<S> U2) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)N2 * U.nrhs) return;
    reduced_rhs_node<S>((int)(t / U.nrhs), (int)(t % U.nrhs), sep, Kl, U, U2);
}

template <class S>
__global__ void k_beam_reduced_scatter(int N2, const int* __restrict__ sep, MVec<S> U2, MVec<S> U) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)N2 * U.nrhs) return;
    reduced_scatter_node<S>((int)(t / U.nrhs), (int)(t % U.nrhs), sep, U2, U);
}

// ---- block tridiagonal matvec / residual: thread per (node, rhs) ---------------------------------
template <class S>
__global__ void k_beam_matvec(long long N, const S* __restrict__ D, const S* __restrict__ Lb, MVec<S> X,
                              const double* __restrict__ F, double* __restrict__ Y) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= N * X.nrhs) return;
    matvec_node<S>(t / X.nrhs, (int)(t % X.nrhs), N, D, Lb, X, F, Y);
}

// first block row entries of the stored K / M rounded to FP64 (introspection for tests): out[k][18] = D, Lb
template <class S>
__global__ void k_beam_export(long long N, const S* __restrict__ D, const S* __restrict__ Lb, double* out) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= N * 18) return;
    const long long k = t / 18;
    const int i = (int)(t % 18);
    out[t] = sc<S>::hi(i < 9 ? D[9 * k + i] : Lb[9 * k + i - 9]);
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
    // blockIdx.y = structure of a batch (bfe2_gram64_batched): every structure has its own [n, 64] pair
    X += (long long)blockIdx.y * n * GQ;
    Y += (long long)blockIdx.y * n * GQ;
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
    double* out = Cpart + ((long long)blockIdx.y * gridDim.x + blockIdx.x) * GQ * GQ;
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
    Z += (long long)blockIdx.y * n * GQ;                                // blockIdx.y = structure of a batch
    Xout += (long long)blockIdx.y * n * GQ;
    Q += (long long)blockIdx.y * GQ * GQ;
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
namespace {

template <class S>
int beam_build(bfe2_beam* h, const double* coords, const double* props, cudaStream_t st) {
    const int T = 128, P = h->P;
    k_beam_assemble<S><<<(unsigned)((h->N + T - 1) / T), T, 0, st>>>(h->N, coords, props, h->d_fixed, (S*)h->Kd,
                                                                    (S*)h->Kl, (S*)h->Md, (S*)h->Ml);
    k_beam_factor<S><<<(P + 31) / 32, 32, 0, st>>>(P, h->d_sep, (const S*)h->Kd, (const S*)h->Kl, (S*)h->Sinv,
                                                   (S*)h->G, h->d_info);
    k_beam_spikes<S><<<(P * 6 + 31) / 32, 32, 0, st>>>(P, h->N, h->d_sep, (const S*)h->Kl, (const S*)h->Sinv,
                                                       (const S*)h->G, (S*)h->Vl, (S*)h->Vr);
    if (h->P2 == 0) {
        k_beam_reduced_factor<S><<<1, 32, 0, st>>>(P, h->d_sep, (const S*)h->Kd, (const S*)h->Kl, (const S*)h->Vl,
                                                   (const S*)h->Vr, (S*)h->RSinv, (S*)h->RG, (S*)h->Rl, h->d_info);
    } else {                                           // the separator system gets the same treatment once more
        const int N2 = (int)h->N2, P2 = h->P2;
        k_beam_reduced_build<S><<<(N2 + 63) / 64, 64, 0, st>>>(N2, h->d_sep, (const S*)h->Kd, (const S*)h->Kl,
                                                               (const S*)h->Vl, (const S*)h->Vr, (S*)h->Kd2, (S*)h->Kl2);
        k_beam_factor<S><<<(P2 + 31) / 32, 32, 0, st>>>(P2, h->d_sep2, (const S*)h->Kd2, (const S*)h->Kl2,
                                                        (S*)h->Sinv2, (S*)h->G2, h->d_info);
        k_beam_spikes<S><<<(P2 * 6 + 31) / 32, 32, 0, st>>>(P2, h->N2, h->d_sep2, (const S*)h->Kl2, (const S*)h->Sinv2,
                                                            (const S*)h->G2, (S*)h->Vl2, (S*)h->Vr2);
        k_beam_reduced_factor<S><<<1, 32, 0, st>>>(P2, h->d_sep2, (const S*)h->Kd2, (const S*)h->Kl2,
                                                   (const S*)h->Vl2, (const S*)h->Vr2, (S*)h->RSinv, (S*)h->RG,
                                                   (S*)h->Rl, h->d_info);
    }
    BEAM_CUDA(cudaGetLastError());
    return 0;
}

template <class S>
int beam_solve(bfe2_beam* h, int nrhs, double* U, double* Ulo, cudaStream_t st) {
    MVec<S> mv{U, Ulo, nrhs};
    const int T = 64;
    const long long t1 = (long long)h->P * nrhs;
    k_beam_solve_part<S><<<(unsigned)((t1 + T - 1) / T), T, 0, st>>>(h->P, h->d_sep, (const S*)h->Sinv,
                                                                     (const S*)h->G, mv);
    if (h->P > 1) {
        if (h->P2 == 0) {
            k_beam_solve_reduced<S><<<(nrhs + 31) / 32, 32, 0, st>>>(h->P, h->d_sep, (const S*)h->Kl,
                                                                     (const S*)h->RSinv, (const S*)h->RG, mv);
        } else {
            const int N2 = (int)h->N2, P2 = h->P2;
            MVec<S> mv2{h->u2_ws, Ulo ? h->u2_ws + 3 * h->N2 * h->u2_cap : nullptr, nrhs};
            const long long t2 = (long long)N2 * nrhs;
            k_beam_reduced_rhs<S><<<(unsigned)((t2 + 127) / 128), 128, 0, st>>>(N2, h->d_sep, (const S*)h->Kl, mv, mv2);
            const long long tp = (long long)P2 * nrhs;
            k_beam_solve_part<S><<<(unsigned)((tp + T - 1) / T), T, 0, st>>>(P2, h->d_sep2, (const S*)h->Sinv2,
                                                                             (const S*)h->G2, mv2);
            if (P2 > 1) {
                k_beam_solve_reduced<S><<<(nrhs + 31) / 32, 32, 0, st>>>(P2, h->d_sep2, (const S*)h->Kl2,
                                                                         (const S*)h->RSinv, (const S*)h->RG, mv2);
                k_beam_solve_fix<S><<<(unsigned)((t2 + 127) / 128), 128, 0, st>>>(h->N2, h->d_sep2, h->d_part2,
                                                                                  (const S*)h->Vl2, (const S*)h->Vr2, mv2);
            }
            k_beam_reduced_scatter<S><<<(unsigned)((t2 + 127) / 128), 128, 0, st>>>(N2, h->d_sep, mv2, mv);
        }
        const long long t3 = h->N * nrhs;
        k_beam_solve_fix<S><<<(unsigned)((t3 + 127) / 128), 128, 0, st>>>(h->N, h->d_sep, h->d_part,
                                                                          (const S*)h->Vl, (const S*)h->Vr, mv);
    }
    BEAM_CUDA(cudaGetLastError());
    return 0;
}

template <class S>
int beam_matvec(bfe2_beam* h, int which, int nrhs, const double* X, const double* Xlo, const double* F, double* Y,
                cudaStream_t st) {
    MVec<S> mv{const_cast<double*>(X), const_cast<double*>(Xlo), nrhs};
    const long long t = h->N * nrhs;
    k_beam_matvec<S><<<(unsigned)((t + 127) / 128), 128, 0, st>>>(h->N, (const S*)(which ? h->Md : h->Kd),
                                                                  (const S*)(which ? h->Ml : h->Kl), mv, F, Y);
    BEAM_CUDA(cudaGetLastError());
    return 0;
}

int beam_lo_workspace(bfe2_beam* h, int nrhs) {
    if (h->prec != BFE2_BEAM_DD || nrhs <= h->lo_cap) return 0;
    BEAM_CUDA(cudaDeviceSynchronize());                // a previous solve may still use the old buffer
    if (h->lo_ws) cudaFree(h->lo_ws);
    h->lo_ws = nullptr;
    h->lo_cap = 0;
    BEAM_CUDA(cudaMalloc(&h->lo_ws, (size_t)h->n * nrhs * sizeof(double)));
    h->lo_cap = nrhs;
    return 0;
}

// right-hand sides of the second-level chain (two-level handles only)
int beam_u2_workspace(bfe2_beam* h, int nrhs) {
    if (h->P2 == 0 || nrhs <= h->u2_cap) return 0;
    BEAM_CUDA(cudaDeviceSynchronize());
    if (h->u2_ws) cudaFree(h->u2_ws);
    h->u2_ws = nullptr;
    h->u2_cap = 0;
    BEAM_CUDA(cudaMalloc(&h->u2_ws, (size_t)2 * 3 * h->N2 * nrhs * sizeof(double)));
    h->u2_cap = nrhs;
    return 0;
}

}  // namespace

extern "C" {

int bfe2_beam_create_ex(bfe2_beam** out, int64_t n_elem, const double* coords, const double* props,
                        const uint8_t* fixed_mask, int32_t n_partitions, int32_t precision, void* stream) {
    if (!out || n_elem < 1 || !coords || !props || !fixed_mask) return bfe2_fail(-1, "bfe2_beam_create: bad argument");
    if (precision != BFE2_BEAM_FP64 && precision != BFE2_BEAM_DD) return bfe2_fail(-1, "bfe2_beam_create: bad precision");
    *out = nullptr;
    cudaStream_t st = (cudaStream_t)stream;
    bfe2_beam* h = new bfe2_beam();
    struct Guard {                                     // every error exit below frees the handle and its blob
        bfe2_beam* h;
        ~Guard() { delete h; }
    } guard{h};
    h->prec = precision;
    h->N = n_elem + 1;
    h->n = 3 * h->N;
    // one level: ~sqrt(N) partitions -- the interiors (length N/P, thread per partition x rhs) and the separator system
    // (length P, eliminated sequentially) are then equally long chains; measured at n = 1e5 in double-double: P = 256
    // 2.4 ms per solve, 1024 3.5 ms, 4096 12 ms (profiles/r2_config4_partitions.txt).  Two levels (automatic from
    // N = 4096 on): the separator system is cut again, three chains of ~cbrt(N) steps each.
    int P = 1, P2 = 0;
    choose_levels(h->N, n_partitions, P, P2);
    h->P = P;
    h->P2 = P2;
    h->N2 = P2 > 0 ? P - 1 : 0;
    std::vector<int> sep, part_of, sep2, part2;
    cut_chain(h->N, P, sep, part_of);
    if (P2 > 0) cut_chain(h->N2, P2, sep2, part2);
    const int PL = P2 > 0 ? P2 : P;                    // partitions of the last level: size of the sequential system
    const size_t es = precision == BFE2_BEAM_DD ? sizeof(dd) : sizeof(double);
    const size_t nb = (size_t)h->N * 9 * es;
    const size_t pb = (size_t)(PL + 1) * 9 * es;
    const size_t nb2 = (size_t)(h->N2 + 1) * 9 * es;
    auto r64 = [](size_t b) { return (b + 63) & ~size_t(63); };
    const size_t total = 8 * r64(nb) + 3 * r64(pb) + r64((size_t)(P + 1) * sizeof(int)) +
                         r64((size_t)h->N * sizeof(int)) + r64(sizeof(int)) + r64((size_t)h->n) + 256 +
                         (P2 > 0 ? 6 * r64(nb2) + r64((size_t)(P2 + 1) * sizeof(int)) + r64((size_t)h->N2 * sizeof(int)) : 0);
    if (cudaMalloc(&h->blob, total) != cudaSuccess) {
        h->blob = nullptr;
        return bfe2_fail(-100, "cudaMalloc of %zu bytes failed", total);
    }
    char* q = (char*)h->blob;
    auto take = [&](size_t b) { char* r = q; q += (b + 63) & ~size_t(63); return (void*)r; };
    h->Kd = take(nb); h->Kl = take(nb); h->Md = take(nb); h->Ml = take(nb);
    h->Sinv = take(nb); h->G = take(nb); h->Vl = take(nb); h->Vr = take(nb);
    h->RSinv = take(pb); h->RG = take(pb); h->Rl = take(pb);
    h->d_sep = (int*)take((P + 1) * sizeof(int));
    h->d_part = (int*)take(h->N * sizeof(int));
    h->d_info = (int*)take(sizeof(int));
    h->d_fixed = (unsigned char*)take(h->n);
    if (P2 > 0) {
        h->Kd2 = take(nb2); h->Kl2 = take(nb2); h->Sinv2 = take(nb2); h->G2 = take(nb2); h->Vl2 = take(nb2); h->Vr2 = take(nb2);
        h->d_sep2 = (int*)take((P2 + 1) * sizeof(int));
        h->d_part2 = (int*)take(h->N2 * sizeof(int));
        BEAM_CUDA(cudaMemcpyAsync(h->d_sep2, sep2.data(), (P2 + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
        BEAM_CUDA(cudaMemcpyAsync(h->d_part2, part2.data(), h->N2 * sizeof(int), cudaMemcpyHostToDevice, st));
    }
    BEAM_CUDA(cudaMemcpyAsync(h->d_sep, sep.data(), (P + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
    BEAM_CUDA(cudaMemcpyAsync(h->d_part, part_of.data(), h->N * sizeof(int), cudaMemcpyHostToDevice, st));
    BEAM_CUDA(cudaMemcpyAsync(h->d_fixed, fixed_mask, h->n, cudaMemcpyHostToDevice, st));
    BEAM_CUDA(cudaMemsetAsync(h->d_info, 0, sizeof(int), st));
    BEAM_CUDA(cudaStreamSynchronize(st));              // host vectors go out of scope
    const int rc = precision == BFE2_BEAM_DD ? beam_build<dd>(h, coords, props, st) : beam_build<double>(h, coords, props, st);
    if (rc) return rc;
    int info = 0;
    BEAM_CUDA(cudaMemcpyAsync(&info, h->d_info, sizeof(int), cudaMemcpyDeviceToHost, st));
    BEAM_CUDA(cudaStreamSynchronize(st));
    if (info != 0) return bfe2_fail(-4, "stiffness matrix is not positive definite (node %d)", info);
    guard.h = nullptr;
    *out = h;
    return 0;
}

int bfe2_beam_create(bfe2_beam** out, int64_t n_elem, const double* coords, const double* props,
                     const uint8_t* fixed_mask, int32_t n_partitions, void* stream) {
    return bfe2_beam_create_ex(out, n_elem, coords, props, fixed_mask, n_partitions, BFE2_BEAM_DD, stream);
}

void bfe2_beam_destroy(bfe2_beam* h) { delete h; }

int bfe2_beam_dims(const bfe2_beam* h, int64_t* n_nodes, int64_t* n_dof, int32_t* n_partitions) {
    if (!h) return bfe2_fail(-1, "beam handle is NULL");
    if (n_nodes) *n_nodes = h->N;
    if (n_dof) *n_dof = h->n;
    if (n_partitions) *n_partitions = h->P;
    return 0;
}

int bfe2_beam_solve_dd(bfe2_beam* h, int32_t nrhs, const double* F, const double* F_lo, double* U, double* U_lo,
                       void* stream) {
    if (!h || nrhs < 1 || !F || !U) return bfe2_fail(-1, "bfe2_beam_solve: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    double* lo = U_lo;
    if (h->prec == BFE2_BEAM_DD && !lo) {
        const int rc = beam_lo_workspace(h, nrhs);
        if (rc) return rc;
        lo = h->lo_ws;
    }
    if (h->prec != BFE2_BEAM_DD) lo = nullptr;
    if (const int rc = beam_u2_workspace(h, nrhs)) return rc;
    const long long count = h->n * (long long)nrhs;
    k_beam_load_rhs<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(count, F, F_lo, U, lo);
    const int rc = h->prec == BFE2_BEAM_DD ? beam_solve<dd>(h, nrhs, U, lo, st) : beam_solve<double>(h, nrhs, U, nullptr, st);
    if (rc) return rc;
    if (h->prec != BFE2_BEAM_DD && U_lo) BEAM_CUDA(cudaMemsetAsync(U_lo, 0, (size_t)count * sizeof(double), st));
    return 0;
}

int bfe2_beam_solve(bfe2_beam* h, int32_t nrhs, const double* F, double* U, void* stream) {
    return bfe2_beam_solve_dd(h, nrhs, F, nullptr, U, nullptr, stream);
}

int bfe2_beam_matvec(bfe2_beam* h, int32_t which, int32_t nrhs, const double* X, double* Y, void* stream) {
    if (!h || nrhs < 1 || !X || !Y || X == Y || which < 0 || which > 1) return bfe2_fail(-1, "bfe2_beam_matvec: bad argument");
    return h->prec == BFE2_BEAM_DD ? beam_matvec<dd>(h, which, nrhs, X, nullptr, nullptr, Y, (cudaStream_t)stream)
                                   : beam_matvec<double>(h, which, nrhs, X, nullptr, nullptr, Y, (cudaStream_t)stream);
}

int bfe2_beam_residual(bfe2_beam* h, int32_t nrhs, const double* F, const double* U, const double* U_lo, double* R,
                       void* stream) {
    if (!h || nrhs < 1 || !F || !U || !R || R == U || R == U_lo) return bfe2_fail(-1, "bfe2_beam_residual: bad argument");
    return h->prec == BFE2_BEAM_DD ? beam_matvec<dd>(h, 0, nrhs, U, U_lo, F, R, (cudaStream_t)stream)
                                   : beam_matvec<double>(h, 0, nrhs, U, nullptr, F, R, (cudaStream_t)stream);
}

int bfe2_beam_export_blocks(bfe2_beam* h, int32_t which, double* out, void* stream) {
    if (!h || !out || which < 0 || which > 1) return bfe2_fail(-1, "bfe2_beam_export_blocks: bad argument");
    const long long t = h->N * 18;
    const unsigned g = (unsigned)((t + 255) / 256);
    if (h->prec == BFE2_BEAM_DD)
        k_beam_export<dd><<<g, 256, 0, (cudaStream_t)stream>>>(h->N, (const dd*)(which ? h->Md : h->Kd),
                                                              (const dd*)(which ? h->Ml : h->Kl), out);
    else
        k_beam_export<double><<<g, 256, 0, (cudaStream_t)stream>>>(h->N, (const double*)(which ? h->Md : h->Kd),
                                                                  (const double*)(which ? h->Ml : h->Kl), out);
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

// batched forms for the first-k modal route of MANY mid-size frames (beamfe2_b200/frame_modes.BatchedFrameModes):
// X, Y, Z, Xout [B, n, 64], C, Q [B, 64, 64]; one block per structure walks all rows (n is a few hundred to a few thousand)
int bfe2_gram64_batched(int64_t n, int64_t B, const double* X, const double* Y, double* C, void* stream) {
    if (n < 1 || B < 1 || B > 65535 || !X || !Y || !C) return bfe2_fail(-1, "bfe2_gram64_batched: bad argument");
    const long long rpb = (n + GCH - 1) / GCH * GCH;
    k_gram_partial<<<dim3(1, (unsigned)B), 256, 0, (cudaStream_t)stream>>>(n, X, Y, C, rpb);
    BEAM_CUDA(cudaGetLastError());
    return 0;
}

int bfe2_update64_batched(int64_t n, int64_t B, const double* Z, const double* Q, double* Xout, void* stream) {
    if (n < 1 || B < 1 || B > 65535 || !Z || !Q || !Xout || Z == Xout) return bfe2_fail(-1, "bfe2_update64_batched: bad argument");
    const size_t smem = 2 * GQ * GS * sizeof(double);
    BEAM_CUDA(cudaFuncSetAttribute(k_update, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_update<<<dim3((unsigned)((n + GQ - 1) / GQ), (unsigned)B), 256, smem, (cudaStream_t)stream>>>(n, Z, Q, Xout);
    BEAM_CUDA(cudaGetLastError());
    return 0;
}

}  // extern "C"
