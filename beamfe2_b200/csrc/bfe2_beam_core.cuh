// bfe2_beam_core.cuh -- per-thread building blocks of the chain ("one large structure") path, templated over
// the scalar type S = double or S = dd (bfe2_dd.cuh).  __host__ __device__ so that the same arithmetic can be
// exercised by a host-side check (tools/beam_host_check.cu) where no GPU is available; the product path calls
// them only from the kernels of bfe2_beam.cu.
//
// A chain of n elements (beam k joins nodes k, k+1) gives block tridiagonal K, M with 3x3 blocks, one block
// row per node: D[k] = block (k,k), Lb[k] = block (k+1,k).  Supported DOFs keep their row/column with
// K_ii = 1, M_ii = 0 and zero couplings (DOF numbering stays the reference's, Structure.py:379-393).
#pragma once

#include <algorithm>
#include <cmath>
#include <vector>
#include "bfe2_dd.cuh"

namespace bfe2 {
namespace chain {

// ---- 3x3 helpers (row-major) ------------------------------------------------------------------
// software prefetch of a line into L1 (device only): the sequential chains below are bound by one memory round trip
// per step unless the operands of the steps ahead are already on their way
#ifdef __CUDA_ARCH__
#define BFE2_PREFETCH(ptr) asm volatile("prefetch.global.L1 [%0];" ::"l"(ptr))
#else
#define BFE2_PREFETCH(ptr) ((void)0)
#endif
constexpr int PF_AHEAD = 3;                       // steps of look-ahead
template <class S> BFE2_HD void m3_prefetch(const S* p) {      // a 3 x 3 block: 72 or 144 bytes, up to three lines
    BFE2_PREFETCH(p);
    BFE2_PREFETCH(p + 4);
    BFE2_PREFETCH(p + 8);
}
template <class S> BFE2_HD void m3_load(S a[9], const S* p) {
#pragma unroll
    for (int i = 0; i < 9; ++i) a[i] = p[i];
}
template <class S> BFE2_HD void m3_store(S* p, const S a[9]) {
#pragma unroll
    for (int i = 0; i < 9; ++i) p[i] = a[i];
}
template <class S> BFE2_HD void m3_mul(S c[9], const S a[9], const S b[9]) {            // c = a b
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) c[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
}
template <class S> BFE2_HD void m3_mul_bt(S c[9], const S a[9], const S b[9]) {         // c = a b^T
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            c[3 * i + j] = a[3 * i] * b[3 * j] + a[3 * i + 1] * b[3 * j + 1] + a[3 * i + 2] * b[3 * j + 2];
}
template <class S> BFE2_HD void m3_mul_at(S c[9], const S a[9], const S b[9]) {         // c = a^T b
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) c[3 * i + j] = a[i] * b[j] + a[3 + i] * b[3 + j] + a[6 + i] * b[6 + j];
}
template <class S> BFE2_HD void m3_vec(S y[3], const S a[9], const S x[3]) {            // y = a x
#pragma unroll
    for (int i = 0; i < 3; ++i) y[i] = a[3 * i] * x[0] + a[3 * i + 1] * x[1] + a[3 * i + 2] * x[2];
}
template <class S> BFE2_HD void m3_vec_t(S y[3], const S a[9], const S x[3]) {          // y = a^T x
#pragma unroll
    for (int i = 0; i < 3; ++i) y[i] = a[i] * x[0] + a[3 + i] * x[1] + a[6 + i] * x[2];
}

// inverse of a symmetric positive definite 3x3 through its Cholesky factor; false if not SPD
template <class S> BFE2_HD bool m3_spd_inv(S inv[9], const S a[9]) {
    typedef sc<S> T;
    const S one = T::make(1.0);
    if (!(T::hi(a[0]) > 0.0)) return false;
    const S l00 = T::root(a[0]), i00 = one / l00;
    const S l10 = a[3] * i00, l20 = a[6] * i00;
    const S d1 = a[4] - l10 * l10;
    if (!(T::hi(d1) > 0.0)) return false;
    const S l11 = T::root(d1), i11 = one / l11;
    const S l21 = (a[7] - l20 * l10) * i11;
    const S d2 = a[8] - l20 * l20 - l21 * l21;
    if (!(T::hi(d2) > 0.0)) return false;
    const S l22 = T::root(d2), i22 = one / l22;
    // W = L^-1 (lower);  A^-1 = W^T W
    const S m00 = i00, m10 = -(l10 * m00 * i11), m11 = i11;
    const S m20 = -((l20 * m00 + l21 * m10) * i22), m21 = -(l21 * m11 * i22), m22 = i22;
    inv[0] = m00 * m00 + m10 * m10 + m20 * m20;
    inv[1] = inv[3] = m10 * m11 + m20 * m21;
    inv[2] = inv[6] = m20 * m22;
    inv[4] = m11 * m11 + m21 * m21;
    inv[5] = inv[7] = m21 * m22;
    inv[8] = m22 * m22;
    return true;
}

// general 3x3 inverse (adjugate): fall-back for a separator block that lost definiteness to cancellation
template <class S> BFE2_HD bool m3_gen_inv(S inv[9], const S a[9]) {
    typedef sc<S> T;
    const S c00 = a[4] * a[8] - a[5] * a[7], c01 = a[5] * a[6] - a[3] * a[8], c02 = a[3] * a[7] - a[4] * a[6];
    const S det = a[0] * c00 + a[1] * c01 + a[2] * c02;
    const double dh = T::hi(det);
    if (dh == 0.0 || !(dh == dh) || dh - dh != 0.0) return false;
    const S r = T::make(1.0) / det;
    inv[0] = c00 * r; inv[1] = (a[2] * a[7] - a[1] * a[8]) * r; inv[2] = (a[1] * a[5] - a[2] * a[4]) * r;
    inv[3] = c01 * r; inv[4] = (a[0] * a[8] - a[2] * a[6]) * r; inv[5] = (a[2] * a[3] - a[0] * a[5]) * r;
    inv[6] = c02 * r; inv[7] = (a[1] * a[6] - a[0] * a[7]) * r; inv[8] = (a[0] * a[4] - a[1] * a[3]) * r;
    return true;
}

// ---- element matrices in global axes (HermitianBeam_2D.py:98-120, 233-249, 284-297; helpers.py:43-69) ----
// Evaluated in S from the FP64 inputs: coordinate differences and E*A, E*I are exact in dd.
template <class S>
BFE2_HD void elem_global(const double* coords, const double* props, long long e, bool mass, S X[6][6]) {
    typedef sc<S> T;
    const S dx = T::diff(coords[2 * e + 2], coords[2 * e]), dy = T::diff(coords[2 * e + 3], coords[2 * e + 1]);
    const S L = T::root(dx * dx + dy * dy);
    const S c = dx / L, s = dy / L;
    const S zero = T::make(0.0);
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int q = 0; q < 6; ++q) X[r][q] = zero;
    if (mass) {
        const S f = T::prod(props[4 * e + 3], props[4 * e + 1]) * L / T::make(420.0);
        const S fl = f * L, fll = fl * L;
        X[0][0] = X[3][3] = T::make(140.0) * f;
        X[0][3] = X[3][0] = T::make(70.0) * f;
        X[1][1] = X[4][4] = T::make(156.0) * f;
        X[1][2] = X[2][1] = T::make(22.0) * fl;
        X[1][4] = X[4][1] = T::make(54.0) * f;
        X[1][5] = X[5][1] = T::make(-13.0) * fl;
        X[2][2] = X[5][5] = T::make(4.0) * fll;
        X[2][4] = X[4][2] = T::make(13.0) * fl;
        X[2][5] = X[5][2] = T::make(-3.0) * fll;
        X[4][5] = X[5][4] = T::make(-22.0) * fl;
    } else {
        const S EA = T::prod(props[4 * e], props[4 * e + 1]), EI = T::prod(props[4 * e], props[4 * e + 2]);
        const S a = EA / L, eil = EI / L;
        const S b6 = T::make(6.0) * eil / L, b12 = T::make(2.0) * b6 / L;
        const S b4 = T::make(4.0) * eil, b2 = T::make(2.0) * eil;
        X[0][0] = X[3][3] = a;
        X[0][3] = X[3][0] = -a;
        X[1][1] = X[4][4] = b12;
        X[1][4] = X[4][1] = -b12;
        X[1][2] = X[2][1] = X[1][5] = X[5][1] = b6;
        X[2][4] = X[4][2] = X[4][5] = X[5][4] = -b6;
        X[2][2] = X[5][5] = b4;
        X[2][5] = X[5][2] = b2;
    }
    // X <- T X T^T, T = diag(R, R), R = [[c,-s,0],[s,c,0],[0,0,1]]
#pragma unroll
    for (int k = 0; k < 6; k += 3)
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            const S p = X[k][j], q = X[k + 1][j];
            X[k][j] = c * p - s * q;
            X[k + 1][j] = s * p + c * q;
        }
#pragma unroll
    for (int k = 0; k < 6; k += 3)
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            const S p = X[i][k], q = X[i][k + 1];
            X[i][k] = c * p - s * q;
            X[i][k + 1] = s * p + c * q;
        }
}

// block row k of K (mass = false) or M (mass = true): element k-1 first, then element k -- beam-list order
// (helpers.py:19-39); supports decoupled
template <class S>
BFE2_HD void assemble_node(long long k, long long N, const double* coords, const double* props,
                           const unsigned char* fixed, bool mass, S* Dout, S* Lout) {
    typedef sc<S> T;
    S D[9], Lb[9];
    const S zero = T::make(0.0);
#pragma unroll
    for (int i = 0; i < 9; ++i) { D[i] = zero; Lb[i] = zero; }
    S X[6][6];
    if (k > 0) {
        elem_global<S>(coords, props, k - 1, mass, X);
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b) D[3 * a + b] = X[3 + a][3 + b];
    }
    if (k < N - 1) {
        elem_global<S>(coords, props, k, mass, X);
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b) {
                D[3 * a + b] = D[3 * a + b] + X[a][b];
                Lb[3 * a + b] = X[3 + a][b];           // block (k+1, k)
            }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const bool fa = fixed[3 * k + a] != 0;
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            const bool fb = fixed[3 * k + b] != 0;
            if (fa || fb) D[3 * a + b] = T::make((a == b && !mass) ? 1.0 : 0.0);
            if (k < N - 1) {
                const bool fr = fixed[3 * (k + 1) + a] != 0;     // row DOF belongs to node k+1
                if (fr || fb) Lb[3 * a + b] = zero;
            }
        }
    }
    m3_store(Dout + 9 * k, D);
    m3_store(Lout + 9 * k, Lb);
}

// ---- partition factorisation (block Cholesky of the interior nodes a..b of one partition) --------
// S_a = D_a;  G_k = L_k S_k^-1;  S_{k+1} = D_{k+1} - G_k L_k^T.  Returns 0 or node + 1 of a non-SPD pivot block.
template <class S>
BFE2_HD int factor_partition(int a, int b, const S* Kd, const S* Kl, S* Sinv, S* G) {
    S Sk[9], Si[9], Lk[9], Gk[9], Tm[9];
    m3_load(Sk, Kd + 9 * (long long)a);
    for (int k = a; k <= b; ++k) {
        if (k + PF_AHEAD < b) {
            m3_prefetch(Kl + 9 * (long long)(k + PF_AHEAD));
            m3_prefetch(Kd + 9 * (long long)(k + PF_AHEAD + 1));
        }
        if (!m3_spd_inv(Si, Sk)) return k + 1;
        m3_store(Sinv + 9 * (long long)k, Si);
        if (k < b) {
            m3_load(Lk, Kl + 9 * (long long)k);
            m3_mul(Gk, Lk, Si);
            m3_store(G + 9 * (long long)k, Gk);
            m3_mul_bt(Tm, Gk, Lk);
            m3_load(Sk, Kd + 9 * (long long)(k + 1));
#pragma unroll
            for (int i = 0; i < 9; ++i) Sk[i] = Sk[i] - Tm[i];
        }
    }
    return 0;
}

// ---- vector accessors: element (node k, dof d, column r) of a multi-vector ------------------------
// FP64 or (hi, lo) pair of FP64 arrays, [n_dof, nrhs] row-major
template <class S> struct MVec;
template <> struct MVec<double> {
    double* hi; double* lo; long long nrhs;
    BFE2_HD double get(long long i) const { return hi[i]; }
    BFE2_HD void put(long long i, double v) const { hi[i] = v; }
    BFE2_HD void prefetch(long long i) const { BFE2_PREFETCH(hi + i); }
};
template <> struct MVec<dd> {
    double* hi; double* lo; long long nrhs;
    BFE2_HD dd get(long long i) const { return dd{hi[i], lo ? lo[i] : 0.0}; }
    BFE2_HD void put(long long i, dd v) const { hi[i] = v.hi; lo[i] = v.lo; }
    BFE2_HD void prefetch(long long i) const { BFE2_PREFETCH(hi + i); if (lo) BFE2_PREFETCH(lo + i); }
};
// one column of a spike block array: element (node k, dof d) at base[9 k + 3 d]
template <class S> struct SpikeCol {
    S* base;
    BFE2_HD S get(long long i) const { return base[i]; }
    BFE2_HD void put(long long i, S v) const { base[i] = v; }
    BFE2_HD void prefetch(long long i) const { BFE2_PREFETCH(base + i); }
};

// solve T y = w in place for one column: y(node k, dof d) at index k*nstride + d*stride + off of `acc`
template <class S, class Acc>
BFE2_HD void part_solve_inplace(const Acc& acc, long long off, long long nstride, long long stride, int a, int b,
                                const S* Sinv, const S* G) {
    S w[3], t[3], g[9];
    // forward: w_k = f_k - G_{k-1} w_{k-1}
    for (int d = 0; d < 3; ++d) w[d] = acc.get(off + a * nstride + d * stride);
    for (int k = a + 1; k <= b; ++k) {
        if (k + PF_AHEAD <= b) {
            m3_prefetch(G + 9 * (long long)(k + PF_AHEAD - 1));
            for (int d = 0; d < 3; ++d) acc.prefetch(off + (k + PF_AHEAD) * nstride + d * stride);
        }
        m3_load(g, G + 9 * (long long)(k - 1));
        const long long o = off + k * nstride;
        for (int d = 0; d < 3; ++d) {
            DotAcc<S> s3;
            s3.init(acc.get(o + d * stride));
            for (int j = 0; j < 3; ++j) s3.sub(g[3 * d + j], w[j]);
            t[d] = s3.get();
        }
        for (int d = 0; d < 3; ++d) {
            w[d] = t[d];
            acc.put(o + d * stride, w[d]);
        }
    }
    // diagonal + backward: y_k = S_k^-1 w_k - G_k^T y_{k+1}
    S y[3];
    for (int k = b; k >= a; --k) {
        if (k - PF_AHEAD >= a) {
            m3_prefetch(Sinv + 9 * (long long)(k - PF_AHEAD));
            m3_prefetch(G + 9 * (long long)(k - PF_AHEAD));
            for (int d = 0; d < 3; ++d) acc.prefetch(off + (k - PF_AHEAD) * nstride + d * stride);
        }
        const long long o = off + k * nstride;
        for (int d = 0; d < 3; ++d) w[d] = acc.get(o + d * stride);
        m3_load(g, Sinv + 9 * (long long)k);
        S v[3], gt[9];
        if (k < b) m3_load(gt, G + 9 * (long long)k);
        for (int d = 0; d < 3; ++d) {
            DotAcc<S> s3;
            s3.init(sc<S>::make(0.0));
            for (int j = 0; j < 3; ++j) s3.add(g[3 * d + j], w[j]);
            if (k < b)
                for (int j = 0; j < 3; ++j) s3.sub(gt[3 * j + d], y[j]);
            v[d] = s3.get();
        }
        for (int d = 0; d < 3; ++d) {
            y[d] = v[d];
            acc.put(o + d * stride, y[d]);
        }
    }
}

// spikes of partition p: column c in 0..5 (0..2: T^-1 E_left, 3..5: T^-1 E_right), 3x3 block per node
template <class S>
BFE2_HD void spike_column(int p, int c, long long N, const int* sep, const S* Kl, const S* Sinv, const S* G,
                          S* Vl, S* Vr) {
    typedef sc<S> T;
    const int sl = sep[p], sr = sep[p + 1];
    const int a = sl + 1, b = sr - 1;
    S* V = (c < 3) ? Vl : Vr;
    const int col = c % 3;
    for (int k = a; k <= b; ++k)
        for (int d = 0; d < 3; ++d) V[9 * (long long)k + 3 * d + col] = T::make(0.0);
    if (c < 3) {
        if (sl < 0) return;                            // no left separator: spike is zero
        for (int d = 0; d < 3; ++d) V[9 * (long long)a + 3 * d + col] = Kl[9 * (long long)sl + 3 * d + col];
    } else {
        if (sr >= N) return;
        for (int d = 0; d < 3; ++d) V[9 * (long long)b + 3 * d + col] = Kl[9 * (long long)b + 3 * col + d];
    }
    SpikeCol<S> acc{V};
    part_solve_inplace<S>(acc, col, 9, 3, a, b, Sinv, G);
}

// reduced (separator) system: build + sequential block factorisation.  Returns 0 or separator node + 1.
template <class S>
BFE2_HD int reduced_factor(int P, const int* sep, const S* Kd, const S* Kl, const S* Vl, const S* Vr,
                           S* RSinv, S* RG, S* Rl) {
    typedef sc<S> T;
    S Sk[9], Si[9], L1[9], L2[9], Tm[9], A[9], Gp[9], Rlj[9];
    const S half = T::make(0.5);
    for (int j = 0; j < P - 1; ++j) {
        const long long s = sep[j + 1];
        m3_load(Sk, Kd + 9 * s);
        m3_load(L1, Kl + 9 * (s - 1));                 // block (s, s-1)
        m3_load(L2, Kl + 9 * s);                       // block (s+1, s)
        m3_load(A, Vr + 9 * (s - 1));
        m3_mul(Tm, L1, A);                             // L[s-1] Vr[s-1]
#pragma unroll
        for (int i = 0; i < 9; ++i) Sk[i] = Sk[i] - Tm[i];
        m3_load(A, Vl + 9 * (s + 1));
        m3_mul_at(Tm, L2, A);                          // L[s]^T Vl[s+1]
#pragma unroll
        for (int i = 0; i < 9; ++i) Sk[i] = Sk[i] - Tm[i];
        m3_load(A, Vl + 9 * (s - 1));
        m3_mul(Rlj, L1, A);                            // block (j, j-1) = -L[s-1] Vl[s-1]
#pragma unroll
        for (int i = 0; i < 9; ++i) Rlj[i] = -Rlj[i];
        m3_store(Rl + 9 * j, Rlj);
        if (j > 0) {                                   // S_j -= Rl_j RSinv_{j-1} Rl_j^T
            m3_load(Si, RSinv + 9 * (j - 1));
            m3_mul(Gp, Rlj, Si);                       // RG_{j-1} = Rl_j RSinv_{j-1}
            m3_store(RG + 9 * (j - 1), Gp);
            m3_mul_bt(Tm, Gp, Rlj);
#pragma unroll
            for (int i = 0; i < 9; ++i) Sk[i] = Sk[i] - Tm[i];
        }
        // symmetrise (the two spike routes agree only to rounding)
        Sk[1] = Sk[3] = half * (Sk[1] + Sk[3]);
        Sk[2] = Sk[6] = half * (Sk[2] + Sk[6]);
        Sk[5] = Sk[7] = half * (Sk[5] + Sk[7]);
        if (!m3_spd_inv(Si, Sk) && !m3_gen_inv(Si, Sk)) return (int)s + 1;
        m3_store(RSinv + 9 * j, Si);
    }
    return 0;
}

// ---- two-level scheme: the separator system of level 1 is itself a block tridiagonal chain (node j = separator
//      sep[j + 1]); instead of eliminating it sequentially (P - 1 steps) it is handed to the same partitioned
//      machinery once more.  Kd2[j] = Schur diagonal S_j, Kl2[j] = block (j + 1, j).  Thread per separator.
template <class S>
BFE2_HD void reduced_build_node(int j, const int* sep, const S* Kd, const S* Kl, const S* Vl, const S* Vr, S* Kd2,
                                S* Kl2) {
    typedef sc<S> T;
    S Sk[9], L1[9], L2[9], A[9], B[9];
    const S half = T::make(0.5);
    const long long s = sep[j + 1];
    m3_load(Sk, Kd + 9 * s);
    m3_load(L1, Kl + 9 * (s - 1));                     // block (s, s-1)
    m3_load(L2, Kl + 9 * s);                           // block (s+1, s)
    m3_load(A, Vr + 9 * (s - 1));
    m3_load(B, Vl + 9 * (s + 1));
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) {
            DotAcc<S> acc;
            acc.init(Sk[3 * a + b]);
            for (int i = 0; i < 3; ++i) acc.sub(L1[3 * a + i], A[3 * i + b]);      // L[s-1] Vr[s-1]
            for (int i = 0; i < 3; ++i) acc.sub(L2[3 * i + a], B[3 * i + b]);      // L[s]^T Vl[s+1]
            Sk[3 * a + b] = acc.get();
        }
    Sk[1] = Sk[3] = half * (Sk[1] + Sk[3]);            // the two spike routes agree only to rounding
    Sk[2] = Sk[6] = half * (Sk[2] + Sk[6]);
    Sk[5] = Sk[7] = half * (Sk[5] + Sk[7]);
    m3_store(Kd2 + 9 * (long long)j, Sk);
    if (j > 0) {                                       // block (j, j-1) = -L[s-1] Vl[s-1]
        m3_load(A, Vl + 9 * (s - 1));
        S R[9];
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) {
                DotAcc<S> acc;
                acc.init(T::make(0.0));
                for (int i = 0; i < 3; ++i) acc.sub(L1[3 * a + i], A[3 * i + b]);
                R[3 * a + b] = acc.get();
            }
        m3_store(Kl2 + 9 * (long long)(j - 1), R);
    }
}

// right-hand side of the separator system: g_j = y_s - L[s-1] y[s-1] - L[s]^T y[s+1] (y = stage-1 result in U) -> U2
template <class S>
BFE2_HD void reduced_rhs_node(int j, int r, const int* sep, const S* Kl, const MVec<S>& U, const MVec<S>& U2) {
    const long long nrhs = U.nrhs, ns = 3 * nrhs;
    const long long s = sep[j + 1];
    S A[9], B[9], ym[3], yp[3];
    m3_load(A, Kl + 9 * (s - 1));
    m3_load(B, Kl + 9 * s);
    for (int d = 0; d < 3; ++d) {
        ym[d] = U.get((s - 1) * ns + r + d * nrhs);
        yp[d] = U.get((s + 1) * ns + r + d * nrhs);
    }
    for (int d = 0; d < 3; ++d) {
        DotAcc<S> acc;
        acc.init(U.get(s * ns + r + d * nrhs));
        for (int i = 0; i < 3; ++i) acc.sub(A[3 * d + i], ym[i]);
        for (int i = 0; i < 3; ++i) acc.sub(B[3 * i + d], yp[i]);
        U2.put((long long)j * ns + r + d * nrhs, acc.get());
    }
}

// separator values back into the full vector
template <class S>
BFE2_HD void reduced_scatter_node(int j, int r, const int* sep, const MVec<S>& U2, const MVec<S>& U) {
    const long long nrhs = U.nrhs, ns = 3 * nrhs;
    const long long s = sep[j + 1];
    for (int d = 0; d < 3; ++d) U.put(s * ns + r + d * nrhs, U2.get((long long)j * ns + r + d * nrhs));
}

// solve, stage 2 (one right-hand side r): reduced rhs, sequential separator solve, x_s into U
template <class S>
BFE2_HD void solve_reduced_rhs(int P, int r, const int* sep, const S* Kl, const S* RSinv, const S* RG,
                               const MVec<S>& U) {
    typedef sc<S> T;
    const long long nrhs = U.nrhs, ns = 3 * nrhs;
    S g[3], t[3], A[9], y[3], w[3];
    for (int d = 0; d < 3; ++d) w[d] = T::make(0.0);
    for (int j = 0; j < P - 1; ++j) {
        if (j + PF_AHEAD < P - 1) {
            const long long s2 = sep[j + PF_AHEAD + 1];
            for (int d = 0; d < 3; ++d) {
                U.prefetch(s2 * ns + r + d * nrhs);
                U.prefetch((s2 - 1) * ns + r + d * nrhs);
                U.prefetch((s2 + 1) * ns + r + d * nrhs);
            }
            m3_prefetch(Kl + 9 * (s2 - 1));
            m3_prefetch(Kl + 9 * s2);
            m3_prefetch(RG + 9 * (long long)(j + PF_AHEAD - 1));
        }
        const long long s = sep[j + 1];
        const long long os = s * ns + r, om = (s - 1) * ns + r, op = (s + 1) * ns + r;
        S yp[3], A2[9], A3[9];
        for (int d = 0; d < 3; ++d) { g[d] = U.get(os + d * nrhs); y[d] = U.get(om + d * nrhs); yp[d] = U.get(op + d * nrhs); }
        m3_load(A, Kl + 9 * (s - 1));                  // L[s-1] y[s-1]
        m3_load(A2, Kl + 9 * s);                       // L[s]^T y[s+1]
        if (j > 0) m3_load(A3, RG + 9 * (j - 1));      // RG_{j-1} w_{j-1}
        for (int d = 0; d < 3; ++d) {
            DotAcc<S> s3;
            s3.init(g[d]);
            for (int i = 0; i < 3; ++i) s3.sub(A[3 * d + i], y[i]);
            for (int i = 0; i < 3; ++i) s3.sub(A2[3 * i + d], yp[i]);
            if (j > 0)
                for (int i = 0; i < 3; ++i) s3.sub(A3[3 * d + i], w[i]);
            t[d] = s3.get();
        }
        for (int d = 0; d < 3; ++d) g[d] = t[d];
        for (int d = 0; d < 3; ++d) { w[d] = g[d]; U.put(os + d * nrhs, w[d]); }
    }
    S x[3];
    for (int d = 0; d < 3; ++d) x[d] = T::make(0.0);
    for (int j = P - 2; j >= 0; --j) {
        if (j - PF_AHEAD >= 0) {
            const long long s2 = sep[j - PF_AHEAD + 1];
            for (int d = 0; d < 3; ++d) U.prefetch(s2 * ns + r + d * nrhs);
            m3_prefetch(RSinv + 9 * (long long)(j - PF_AHEAD));
            m3_prefetch(RG + 9 * (long long)(j - PF_AHEAD));
        }
        const long long s = sep[j + 1];
        const long long os = s * ns + r;
        for (int d = 0; d < 3; ++d) w[d] = U.get(os + d * nrhs);
        S v[3], A2[9];
        m3_load(A, RSinv + 9 * j);
        if (j < P - 2) m3_load(A2, RG + 9 * j);
        for (int d = 0; d < 3; ++d) {
            DotAcc<S> s3;
            s3.init(T::make(0.0));
            for (int i = 0; i < 3; ++i) s3.add(A[3 * d + i], w[i]);
            if (j < P - 2)
                for (int i = 0; i < 3; ++i) s3.sub(A2[3 * i + d], x[i]);
            v[d] = s3.get();
        }
        for (int d = 0; d < 3; ++d) { x[d] = v[d]; U.put(os + d * nrhs, x[d]); }
    }
}

// solve, stage 3 (interior node k of partition p, right-hand side r): x = y - Vl x_left - Vr x_right
template <class S>
BFE2_HD void solve_fix_node(long long k, int r, int p, long long N, const int* sep, const S* Vl, const S* Vr,
                            const MVec<S>& U) {
    const long long nrhs = U.nrhs, ns = 3 * nrhs;
    const long long sl = sep[p], sr = sep[p + 1];
    const long long o = k * ns + r;
    S y[3], A[9], B[9], xl[3], xr[3];
    const bool left = sl >= 0, right = sr < N;
    for (int d = 0; d < 3; ++d) y[d] = U.get(o + d * nrhs);
    if (left) {
        for (int d = 0; d < 3; ++d) xl[d] = U.get(sl * ns + r + d * nrhs);
        m3_load(A, Vl + 9 * k);
    }
    if (right) {
        for (int d = 0; d < 3; ++d) xr[d] = U.get(sr * ns + r + d * nrhs);
        m3_load(B, Vr + 9 * k);
    }
    for (int d = 0; d < 3; ++d) {
        DotAcc<S> s3;
        s3.init(y[d]);
        if (left)
            for (int i = 0; i < 3; ++i) s3.sub(A[3 * d + i], xl[i]);
        if (right)
            for (int i = 0; i < 3; ++i) s3.sub(B[3 * d + i], xr[i]);
        y[d] = s3.get();
    }
    for (int d = 0; d < 3; ++d) U.put(o + d * nrhs, y[d]);
}

// block tridiagonal product, block row k, column r:  y = Lb[k-1] x[k-1] + D[k] x[k] + Lb[k]^T x[k+1]
// (accumulated in S, so with S = dd the FP64 result is the correctly rounded product);  if F != nullptr the
// result is F - y instead (residual).
template <class S>
BFE2_HD void matvec_node(long long k, int r, long long N, const S* D, const S* Lb, const MVec<S>& X,
                         const double* F, double* Y) {
    typedef sc<S> T;
    const long long nrhs = X.nrhs, ns = 3 * nrhs;
    S A[9], Am[9], Ap[9], x[3], xm[3], xp[3], y[3];
    const bool lower = k > 0, upper = k < N - 1;
    for (int d = 0; d < 3; ++d) x[d] = X.get(k * ns + r + d * nrhs);
    m3_load(A, D + 9 * k);
    if (lower) {
        for (int d = 0; d < 3; ++d) xm[d] = X.get((k - 1) * ns + r + d * nrhs);
        m3_load(Am, Lb + 9 * (k - 1));
    }
    if (upper) {
        for (int d = 0; d < 3; ++d) xp[d] = X.get((k + 1) * ns + r + d * nrhs);
        m3_load(Ap, Lb + 9 * k);
    }
    for (int d = 0; d < 3; ++d) {
        DotAcc<S> s3;
        s3.init(T::make(0.0));
        for (int i = 0; i < 3; ++i) s3.add(A[3 * d + i], x[i]);
        if (lower)
            for (int i = 0; i < 3; ++i) s3.add(Am[3 * d + i], xm[i]);
        if (upper)
            for (int i = 0; i < 3; ++i) s3.add(Ap[3 * i + d], xp[i]);
        y[d] = s3.get();
    }
    for (int d = 0; d < 3; ++d) {
        const long long o = k * ns + r + d * nrhs;
        Y[o] = F ? T::hi(T::make(F[o]) - y[d]) : T::hi(y[d]);
    }
}

// ---- host side: how a chain of N block rows is cut (shared by bfe2_beam.cu and tools/beam_host_check.cu) ------------
// P partitions separated by single nodes: sep[0] = -1, sep[P] = N, part[k] = partition of an interior node, -1 for a
// separator.
inline void cut_chain(long long N, int P, std::vector<int>& sep, std::vector<int>& part) {
    sep.assign(P + 1, 0);
    sep[0] = -1;
    sep[P] = (int)N;
    for (int j = 1; j < P; ++j) sep[j] = (int)((long long)j * N / P);
    part.assign(N, 0);
    for (int p = 0; p < P; ++p) {
        for (int k = sep[p] + 1; k < sep[p + 1]; ++k) part[k] = p;
        if (p + 1 < P) part[sep[p + 1]] = -1;
    }
}
// request > 0: ONE level with that many partitions; request < 0: TWO levels with -request partitions on the first;
// request = 0: automatic -- one level with ~sqrt(N) partitions for short chains, two levels from N = 4096 on with
// interiors, second-level interiors and the last sequential system all ~cbrt(N) long (the three chains a solve walks).
// P2 = 0 means one level.  Every partition keeps at least one interior node.
inline void choose_levels(long long N, int request, int& P1, int& P2) {
    auto fit = [](long long n, long long p) { while (p > 1 && n < 3 * p) --p; return (int)std::max<long long>(1, p); };
    const bool two = request < 0 || (request == 0 && N >= 4096);
    if (request > 0) P1 = fit(N, request);
    else if (request < 0) P1 = fit(N, -(long long)request);
    else if (!two) P1 = fit(N, std::min<long long>(4096, (long long)std::sqrt((double)N)));
    else P1 = fit(N, (long long)(N / std::max(4.0, std::cbrt((double)N))));
    P2 = 0;
    if (two && P1 - 1 >= 6) P2 = fit(P1 - 1, (long long)std::sqrt((double)(P1 - 1)));
}

}  // namespace chain
}  // namespace bfe2
