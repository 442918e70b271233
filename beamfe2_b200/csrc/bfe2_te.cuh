// bfe2_te.cuh -- "two-ended" dense modal engine for n_free <= 64.
//
// The Jacobi engine (bfe2_device.cuh, stage_jacobi_regs) needs ~7 sweeps of n^2/2 rotations: 7.6 MFLOP executed per
// P20 structure for 2.1 MFLOP of canonical work.  This engine follows the LAPACK route instead -- Householder
// tridiagonalisation, Sturm bisection, inverse iteration, back-transformation: ~1.3 MFLOP -- and repairs the one
// thing that route cannot do on its own: it is only ABSOLUTELY accurate (error ~ eps * lam_max), so on the P20 sweep
// (lam_max / lam_min up to 6e7) scipy's own small eigenvalues are off by up to 5e-10.  The repair is to solve the
// problem from BOTH ENDS:
//      C    = Lm^-1 K Lm^-T   = F F^T,        F = Lm^-1 Lk   -> eigenvalues lam,      accurate to eps * lam_max
//      C^-1 = Lm^T K^-1 Lm    = W^T W,        W = Lk^-1 Lm   -> eigenvalues 1 / lam,  accurate to eps / lam_min
// and to take every eigenvalue from the end it is closer to (split at sqrt(lam_min lam_max)): the relative error is
// then ~eps * sqrt(lam_max / lam_min) (1e-12 on the sweep; 2e-11 worst against the 40-digit truth fixture with
// LAPACK standing in, tools/te_host_check.py).  The wanted (lowest) mode shapes are eigenvectors of the LARGEST
// eigenvalues of C^-1: inverse iteration on its tridiagonal form, back-transformed, phi = Lm^-T u.
//
// Layout: one block of 64 threads per structure, thread i owns ROW i of the dense matrix in registers.
// The code is written in bulk-synchronous style against a tiny context (TE_SYNC, te_block_sum, te_quad_sum) so that
// tools/te_host_check.cu can run the very same source on the CPU (64 std::threads + std::barrier per block) where no
// GPU is available; the product instantiates it only inside k_solve_te (bfe2_solve_te.cu).
#pragma once

#include <math.h>
#include <stdint.h>

#ifdef BFE2_TE_HOST
#include <barrier>
#include <cstring>
#define TE_FN inline
struct TeCtx {
    int tid;
    std::barrier<>* bar;
    double* red;        // [2 * 64 + 16] scratch for reductions (double buffered)
    int flip;
};
#define TE_SYNC(ctx) (ctx).bar->arrive_and_wait()
#define TE_RSQRT(x) (1.0 / sqrt(x))
#define TE_RCP(x) (1.0 / (x))
#else
#define TE_FN __device__ __forceinline__
struct TeCtx {
    int tid;
    double* red;        // [8] scratch for cross-warp reductions (double buffered)
    int flip;
};
#define TE_SYNC(ctx) __syncthreads()
#define TE_RSQRT(x) bfe2::rsqrt_fast(x)
#define TE_RCP(x) bfe2::rcp_fast(x)
#endif

namespace bfe2 {
namespace te {

constexpr int NT = 64;            // threads per block = row capacity
constexpr int MAXM = 16;          // most mode shapes the engine returns (more: use the Jacobi engine)

// sum of x over the 64 threads, the same bits in every thread; contains ONE barrier (scratch is double buffered, so
// two consecutive calls need no barrier in between)
TE_FN double te_block_sum(TeCtx& c, double x) {
#ifdef BFE2_TE_HOST
    double* r = c.red + (c.flip ? NT : 0);
    c.flip ^= 1;
    r[c.tid] = x;
    TE_SYNC(c);
    double s = 0.0;
    for (int i = 0; i < NT; ++i) s += r[i];
    return s;
#else
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    double* r = c.red + (c.flip ? 2 : 0);
    c.flip ^= 1;
    if ((c.tid & 31) == 0) r[c.tid >> 5] = x;
    __syncthreads();
    return r[0] + r[1];
#endif
}

// sum over the 4 consecutive threads of a quad (tid / 4), same value in the four
TE_FN double te_quad_sum(TeCtx& c, double x) {
#ifdef BFE2_TE_HOST
    double* r = c.red + (c.flip ? NT : 0);
    c.flip ^= 1;
    r[c.tid] = x;
    TE_SYNC(c);
    const int b = c.tid & ~3;
    return (r[b] + r[b + 1]) + (r[b + 2] + r[b + 3]);
#else
    x += __shfl_xor_sync(0xffffffffu, x, 1);
    x += __shfl_xor_sync(0xffffffffu, x, 2);
    return x;
#endif
}

// ---- Householder tridiagonalisation of the symmetric matrix whose row `tid` is a[0..N) -------------------------
// Column k is annihilated below the sub-diagonal by H_k = I - beta v v^T (v[k+1..n)), A <- H_k A H_k.  All threads run
// full-width loops over the row: entries of v, q at positions <= k are zero, so no per-entry predicates are needed.
// Results: s_d[0..n), s_e[0..n-1) (sub-diagonal); if s_H != nullptr the vectors are kept for the back-transformation:
// s_H[k * ldh + i] = v_k[i], s_beta[k].
// s_v / s_q: [NT] each.  ~3 N fused multiply-adds per thread and step.  N <= NT.
template <int N>
TE_FN void tridiagonalize(TeCtx& c, double (&a)[N], int n, double* s_vq, double* s_x, double* s_d, double* s_e,
                          double* s_H, int ldh, double* s_beta) {
    // s_vq: [2 NT] interleaved (v_j, q_j) pairs, so that the rank-2 update reads both with one 16-byte load;
    // s_x:  [NT] column k + 1 of the CURRENT matrix, written by the update of step k (thread i holds A[i][k+1] in
    //       a[k+1], but k is a run-time index: the hand-off through shared memory avoids a select chain over the row)
    const int i = c.tid;
    s_vq[2 * i] = 0.0;
    s_vq[2 * i + 1] = 0.0;
    s_x[i] = a[0];
    TE_SYNC(c);
    for (int k = 0; k + 2 < n; ++k) {
        // x = column k below the diagonal (rows i > k); the diagonal entry itself is d[k]
        const double ak = s_x[i];
        const double xi = (i > k && i < n) ? ak : 0.0;
        if (i == k) s_d[k] = ak;
        s_vq[2 * i] = xi;
        const double sigma = te_block_sum(c, xi * xi);                   // barrier: s_vq visible
        const double x1 = s_vq[2 * (k + 1)];
        const double tail = sigma - x1 * x1;                             // what has to be annihilated
        double alpha, beta, v1;
        if (!(tail > 0.0)) {                                             // nothing to do: H = I
            alpha = x1; beta = 0.0; v1 = 0.0;
        } else {
            const double nrm = sqrt(sigma);
            alpha = x1 > 0.0 ? -nrm : nrm;
            v1 = x1 - alpha;
            beta = TE_RCP(sigma - x1 * alpha);                           // 2 / |v|^2
        }
        TE_SYNC(c);                                                      // everybody has read x1
        if (i == k + 1) s_vq[2 * i] = v1;
        if (i == 0) s_e[k] = alpha;
        TE_SYNC(c);
        const double vi = (i > k && i < n) ? s_vq[2 * i] : 0.0;
        if (s_H != nullptr && i < n) {
            s_H[k * ldh + i] = beta == 0.0 ? 0.0 : vi;
            if (i == 0) s_beta[k] = beta;
        }
        // p = beta A v on the active rows (four accumulators: the row is long, the FMA latency is not)
        double p0 = 0.0, p1 = 0.0, p2 = 0.0, p3 = 0.0;
#pragma unroll
        for (int j = 0; j + 3 < N; j += 4) {
            p0 = fma(a[j], s_vq[2 * j], p0);
            p1 = fma(a[j + 1], s_vq[2 * j + 2], p1);
            p2 = fma(a[j + 2], s_vq[2 * j + 4], p2);
            p3 = fma(a[j + 3], s_vq[2 * j + 6], p3);
        }
#pragma unroll
        for (int j = N & ~3; j < N; ++j) p0 = fma(a[j], s_vq[2 * j], p0);
        double p = (p0 + p1) + (p2 + p3);
        p = (i > k && i < n) ? p * beta : 0.0;
        const double kk = 0.5 * beta * te_block_sum(c, vi * p);          // barrier inside
        const double qi = p - kk * vi;
        s_vq[2 * i + 1] = qi;
        TE_SYNC(c);
        if (beta != 0.0) {
#pragma unroll
            for (int j = 0; j < N; ++j) a[j] = fma(-vi, s_vq[2 * j + 1], fma(-qi, s_vq[2 * j], a[j]));
        }
#pragma unroll
        for (int j = 0; j < N; ++j)
            if (j == k + 1) s_x[i] = a[j];                               // next column: predicated stores, no chain
        TE_SYNC(c);                                                      // s_vq is rewritten by the next step
    }
    // the trailing 2 x 2 block (or the whole matrix when n <= 2)
    if (i < n) {
        double dii = 0.0, sub = 0.0;
#pragma unroll
        for (int j = 0; j < N; ++j) { dii = (j == i) ? a[j] : dii; sub = (j + 1 == i) ? a[j] : sub; }
        if (i + 2 >= n) s_d[i] = dii;
        if (i == n - 1 && n >= 2) s_e[n - 2] = sub;
    }
    TE_SYNC(c);
}

// ---- number of eigenvalues of the tridiagonal (d, e) smaller than x: Sturm sequence in the division-free form
//      p_i = (d_i - x) p_{i-1} - e_{i-1}^2 p_{i-2}; a sign change between consecutive p's is one eigenvalue below x.
//      e2[i] = e_i^2.  Signs are read from the sign bit (an exact zero counts as positive: p_{i+1} = -e^2 p_{i-1} then
//      has the sign opposite to p_{i-1}, so the triple registers exactly one change either way).  NC recurrences run
//      side by side (instruction-level parallelism for a latency-bound thread).  The pairs (p, p_prev) are rescaled
//      by a power of two every 8 steps: with |T| <= 1 (normalise_tridiagonal) eight steps cannot leave the range.
TE_FN uint32_t te_hi(double x) {
#ifdef BFE2_TE_HOST
    uint64_t u;
    memcpy(&u, &x, 8);
    return (uint32_t)(u >> 32);
#else
    return (uint32_t)__double2hiint(x);
#endif
}
TE_FN double te_pow2_from_hi(uint32_t hi) {              // the double whose high word is hi and low word 0
#ifdef BFE2_TE_HOST
    const uint64_t u = (uint64_t)hi << 32;
    double x;
    memcpy(&x, &u, 8);
    return x;
#else
    return __hiloint2double((int)hi, 0);
#endif
}

template <int NC>
TE_FN void sturm_counts(const double* d, const double* e2, int n, const double (&x)[NC], int (&cnt)[NC]) {
    double pp[NC], p[NC];
    uint32_t sg[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) {
        pp[k] = 1.0;
        p[k] = d[0] - x[k];
        sg[k] = te_hi(p[k]);
        cnt[k] = (int)(sg[k] >> 31);
    }
    for (int i = 1; i < n; ++i) {
        const double di = d[i], ei = e2[i - 1];
#pragma unroll
        for (int k = 0; k < NC; ++k) {
            const double t = fma(di - x[k], p[k], -(ei * pp[k]));
            pp[k] = p[k];
            p[k] = t;
            const uint32_t h = te_hi(t);
            cnt[k] += (int)((h ^ sg[k]) >> 31);
            sg[k] = h;
        }
        if ((i & 7) == 7) {
#pragma unroll
            for (int k = 0; k < NC; ++k) {
                // s = 2^(1023 - exponent(p)) (exponent field clamped to [1, 2046]): |p s| ~ 1
                int ex = (int)((sg[k] >> 20) & 0x7ff);
                int sx = 2046 - ex;
                sx = sx < 1 ? 1 : (sx > 2046 ? 2046 : sx);
                const double sc = te_pow2_from_hi((uint32_t)sx << 20);
                p[k] *= sc;
                pp[k] *= sc;
            }
        }
    }
}

TE_FN int sturm_count(const double* d, const double* e2, int n, double x) {
    const double xs[1] = {x};
    int c[1];
    sturm_counts<1>(d, e2, n, xs, c);
    return c[0];
}

// eigenvalue number `idx` (0-based, ascending) of the tridiagonal in [lo, hi]: the interval is cut in FOUR per step
// (three interior Sturm counts evaluated side by side), then plain bisection finishes the last bits
TE_FN double bisect(const double* d, const double* e2, int n, int idx, double lo, double hi) {
    for (int it = 0; it < 26; ++it) {
        const double w = 0.25 * (hi - lo);
        const double xs[3] = {lo + w, lo + 2.0 * w, hi - w};
        if (!(xs[0] > lo && xs[1] > xs[0] && xs[2] > xs[1] && xs[2] < hi)) break;   // the interval is a few ulps wide
        int c[3];
        sturm_counts<3>(d, e2, n, xs, c);
        if (c[0] > idx) hi = xs[0];
        else if (c[1] > idx) { lo = xs[0]; hi = xs[1]; }
        else if (c[2] > idx) { lo = xs[1]; hi = xs[2]; }
        else lo = xs[2];
    }
    for (int it = 0; it < 8; ++it) {
        const double mid = 0.5 * (lo + hi);
        if (!(mid > lo && mid < hi)) break;                              // the interval is one ulp wide
        if (sturm_count(d, e2, n, mid) > idx) hi = mid; else lo = mid;
    }
    return 0.5 * (lo + hi);
}

// ---- eigenvector of the tridiagonal (d, e) for the eigenvalue mu by inverse iteration: (T - mu I) = P L U with
//      partial pivoting (the dlagtf / dlagts scheme), two solves from a fixed start vector.  Work arrays of length n
//      live in the caller's scratch (w: 5 n doubles); the result z[0..n) is normalised to |z| = 1.
TE_FN void tridiag_eigvec(const double* d, const double* e, int n, double mu, double eps_abs, int seed, double* z,
                          double* w) {
    double* a = w;            // diagonal of U
    double* b = w + n;        // first super-diagonal of U
    double* cc = w + 2 * n;   // second super-diagonal of U (from row interchanges)
    double* l = w + 3 * n;    // multipliers
    double* sw = w + 4 * n;   // 1.0 if rows i, i+1 were interchanged
    // factorisation
    double ai = d[0] - mu, bi = n > 1 ? e[0] : 0.0;
    for (int i = 0; i + 1 < n; ++i) {
        const double sub = e[i];
        const double dn = d[i + 1] - mu, en = (i + 2 < n) ? e[i + 1] : 0.0;
        if (fabs(ai) >= fabs(sub)) {                                     // no interchange
            if (ai == 0.0) ai = eps_abs;
            const double m = sub / ai;
            a[i] = ai; b[i] = bi; cc[i] = 0.0; l[i] = m; sw[i] = 0.0;
            ai = dn - m * bi;
            bi = en;
        } else {                                                         // interchange rows i and i + 1
            const double m = ai / sub;
            a[i] = sub; b[i] = dn; cc[i] = en; l[i] = m; sw[i] = 1.0;
            ai = bi - m * dn;
            bi = -m * en;
        }
    }
    if (fabs(ai) < eps_abs) ai = ai < 0.0 ? -eps_abs : eps_abs;
    a[n - 1] = ai; b[n - 1] = 0.0; cc[n - 1] = 0.0;
    for (int i = 0; i + 1 < n; ++i)
        if (fabs(a[i]) < eps_abs) a[i] = a[i] < 0.0 ? -eps_abs : eps_abs;
    // start vector: fixed pseudo-random signs / magnitudes (no dependence on thread scheduling)
    uint32_t s = 0x9e3779b9u * (uint32_t)(seed + 1);
    for (int i = 0; i < n; ++i) {
        s = s * 1664525u + 1013904223u;
        z[i] = 0.5 + (double)(s >> 8) * (1.0 / 16777216.0);              // in [0.5, 1.5): a U-solve start (EISPACK tinvit)
    }
    for (int pass = 0; pass < 3; ++pass) {
        if (pass > 0) {                                                  // forward: apply P L^-1
            for (int i = 0; i + 1 < n; ++i) {
                if (sw[i] != 0.0) { const double t = z[i]; z[i] = z[i + 1]; z[i + 1] = t - l[i] * z[i]; }
                else z[i + 1] -= l[i] * z[i];
            }
        }
        for (int i = n - 1; i >= 0; --i) {                               // back substitution with U
            double t = z[i];
            if (i + 1 < n) t -= b[i] * z[i + 1];
            if (i + 2 < n) t -= cc[i] * z[i + 2];
            z[i] = t / a[i];
        }
        double nrm = 0.0, big = 0.0;
        for (int i = 0; i < n; ++i) big = fmax(big, fabs(z[i]));
        const double ib = 1.0 / big;
        for (int i = 0; i < n; ++i) { z[i] *= ib; nrm = fma(z[i], z[i], nrm); }
        const double in = 1.0 / sqrt(nrm);
        for (int i = 0; i < n; ++i) z[i] *= in;
    }
}


// ---- band helpers (lower band, row-major: Ab[d * n + j] = A[j + d][j]; after factorisation the band holds L and
//      inv[j] = 1 / L[j][j]) -- plain bulk-synchronous versions; only the modal block below uses them ------------
// in-place band Cholesky by the threads of the block: per column, thread t < b (b + 1) / 2 owns one trailing-update
// pair.  Returns 0 or the 1-based column of a non-positive pivot (same value in every thread).
TE_FN int band_cholesky_block(TeCtx& c, double* Ab, double* inv, int n, int b) {
    int info = 0;
    for (int j = 0; j < n; ++j) {
        const double piv = Ab[j];
        if (info == 0 && !(piv > 0.0 && piv < INFINITY)) info = j + 1;
        const int m = b < n - 1 - j ? b : n - 1 - j;
        if (info == 0) {
            const double ip = 1.0 / piv;
            // pairs (p, q), 1 <= q <= p <= m:  A[j+p][j+q] -= A[j+p][j] A[j+q][j] / piv
            for (int t = c.tid; t < m * (m + 1) / 2; t += NT) {
                int p = 1;
                while (p * (p + 1) / 2 <= t) ++p;
                const int q = t - p * (p - 1) / 2 + 1;
                Ab[(p - q) * n + j + q] = fma(-(Ab[p * n + j] * ip), Ab[q * n + j], Ab[(p - q) * n + j + q]);
            }
        }
        TE_SYNC(c);
    }
    if (info == 0) {
        for (int j = c.tid; j < n; j += NT) {
            const double r = 1.0 / sqrt(Ab[j]);
            inv[j] = r;
        }
        TE_SYNC(c);
        for (int t = c.tid; t < (b + 1) * n; t += NT) {
            const int d = t / n, j = t % n;
            if (j + d < n) Ab[t] *= inv[j];
        }
    }
    TE_SYNC(c);
    return info;
}

// x[first..n) <- L^-1 x (forward substitution with the band factor), entries before `first` are known zeros
TE_FN void band_forward(const double* Lb, const double* inv, double* x, int stride, int n, int b, int first) {
    for (int r = first; r < n; ++r) {
        double acc = x[r * stride];
        const int k0 = r - b > first ? r - b : first;
        for (int k = k0; k < r; ++k) acc = fma(-Lb[(r - k) * n + k], x[k * stride], acc);
        x[r * stride] = acc * inv[r];
    }
}
// x <- L^-T x
TE_FN void band_backward(const double* Lb, const double* inv, double* x, int n, int b) {
    for (int j = n - 1; j >= 0; --j) {
        double acc = x[j];
        const int m = b < n - 1 - j ? b : n - 1 - j;
        for (int d = 1; d <= m; ++d) acc = fma(-Lb[d * n + j], x[j + d], acc);
        x[j] = acc * inv[j];
    }
}

// sum_{r >= r0} ci[r] cj[r]: kept out of line on the device -- it is used from a fully unrolled loop over the row
#ifndef BFE2_TE_HOST
__device__ __noinline__
#else
inline
#endif
double column_dot(const double* ci, const double* cj, int r0, int n) {
    double a0 = 0.0, a1 = 0.0;
    int r = r0;
    for (; r + 1 < n; r += 2) { a0 = fma(ci[r], cj[r], a0); a1 = fma(ci[r + 1], cj[r + 1], a1); }
    if (r < n) a0 = fma(ci[r], cj[r], a0);
    return a0 + a1;
}

// scale a tridiagonal by a power of two so that its Gershgorin radius is <= 1; returns the factor g (T_true = g T)
TE_FN double normalise_tridiagonal(TeCtx& c, double* d, double* e, double* e2, int n, double* s_tmp) {
    if (c.tid == 0) {
        double g = 0.0;
        for (int i = 0; i < n; ++i) {
            const double r = fabs(d[i]) + (i > 0 ? fabs(e[i - 1]) : 0.0) + (i + 1 < n ? fabs(e[i]) : 0.0);
            g = fmax(g, r);
        }
        int ex = 0;
        frexp(g, &ex);                                                   // g = m 2^ex, m in [0.5, 1)
        s_tmp[0] = ldexp(1.0, ex);
    }
    TE_SYNC(c);
    const double g = s_tmp[0], ig = 1.0 / g;
    for (int i = c.tid; i < n; i += NT) {
        d[i] *= ig;
        if (i + 1 < n) { e[i] *= ig; e2[i] = e[i] * e[i]; }
    }
    TE_SYNC(c);
    return g;
}

// Shared-memory carve-up of the modal block (offsets in doubles from `base`)
struct Smem {
    int W, v, q, d0, e0, f0, d1, e1, f1, beta, lam, mu, Z, red, tmp, total;
};
#ifdef BFE2_TE_HOST
inline
#else
__host__ __device__ inline
#endif
Smem smem_layout(int n, int LD, int n_modes) {
    Smem s;
    int o = 0;
    s.W = o;    o += n * LD;           // C, then W = Lk^-1 Lm, then the Householder vectors of C^-1
    s.v = o;    o += 2 * NT;       // (v, q) pairs of the Householder step
    s.q = o;    o += NT;           // next column hand-off
    s.d0 = o;   o += NT;   s.e0 = o;  o += NT;   s.f0 = o;  o += NT;     // tridiagonal of C: d, e, e^2
    s.d1 = o;   o += NT;   s.e1 = o;  o += NT;   s.f1 = o;  o += NT;     // tridiagonal of C^-1
    s.beta = o; o += NT;
    s.lam = o;  o += NT;
    s.mu = o;   o += NT;
    s.Z = o;    o += (n_modes > 0 ? n_modes : 1) * NT;   // eigenvectors / shapes
#ifdef BFE2_TE_HOST
    s.red = o;  o += 2 * NT + 16;
#else
    s.red = o;  o += 8;
#endif
    s.tmp = o;  o += 8;
    s.total = o;
    return s;
}

// ---- the modal block: K phi = lam M phi, all n eigenvalues (ascending) and the first n_modes shapes ------------
// in : s_Kb, s_Mb  UNFACTORED lower bands [(b+1) n] of the condensed K and M (both overwritten by their Cholesky
//      factors; s_invK / s_invM receive the reciprocal diagonals -- the caller's static stage uses the factor of K)
// out: s.lam [n] eigenvalues; s.Z [n_modes][NT] shapes on the free DOFs, M-normalised, largest entry positive
// returns 0, k > 0 (K not positive definite at column k), -1 (M not positive definite), -3 (this engine declines:
//      run the Jacobi engine on the structure)
template <int N>
TE_FN int modal_block(TeCtx& c, int n, int b, double* s_Kb, double* s_Mb, double* s_invK, double* s_invM,
                      double* base, int LD, int n_modes, int* info_k) {
    const Smem S = smem_layout(n, LD, n_modes);
    double* W = base + S.W;
    const int i = c.tid;
    // 1. M = Lm Lm^T
    const int infoM = band_cholesky_block(c, s_Mb, s_invM, n, b);
    // 2. C = Lm^-1 K Lm^-T, column-wise then row-wise forward substitutions (W[j * LD + i] = C[i][j])
    for (int t = i; t < n * LD; t += NT) W[t] = 0.0;
    TE_SYNC(c);
    if (i < n) {
        double* col = W + i * LD;
        const int lo = i - b > 0 ? i - b : 0, hi = i + b < n - 1 ? i + b : n - 1;
        for (int r = lo; r < i; ++r) col[r] = s_Kb[(i - r) * n + r];
        for (int r = i; r <= hi; ++r) col[r] = s_Kb[(r - i) * n + i];
        if (infoM == 0) band_forward(s_Mb, s_invM, col, 1, n, b, lo);
    }
    TE_SYNC(c);
    if (i < n && infoM == 0) band_forward(s_Mb, s_invM, W + i, LD, n, b, 0);
    TE_SYNC(c);
    // 3. K = Lk Lk^T (the static stage needs it even when the modal part fails)
    *info_k = band_cholesky_block(c, s_Kb, s_invK, n, b);
    if (*info_k != 0) return *info_k;
    if (infoM != 0) return -1;
    // 4. + 5. the two ends, one after the other through ONE instance of the tridiagonalisation (code size):
    //    which = 0: rows of C into registers, eigenvalues only (no vectors kept);
    //    which = 1: W = Lk^-1 Lm (lower triangular, column j starts at row j), rows of C^-1 = W^T W, vectors kept in W
    double a[N];
    double g[2] = {1.0, 1.0};
    for (int which = 0; which < 2; ++which) {
        if (which == 1) {
            if (i < n) {
                double* col = W + i * LD;
                for (int r = 0; r < n; ++r) col[r] = 0.0;
                col[i] = 1.0 / s_invM[i];                                // Lm[i][i]
                for (int d = 1; d <= b && i + d < n; ++d) col[i + d] = s_Mb[d * n + i];
                band_forward(s_Kb, s_invK, col, 1, n, b, i);
            }
            TE_SYNC(c);
        }
#pragma unroll
        for (int j = 0; j < N; ++j) {
            double v = 0.0;
            if (i < n && j < n) v = which == 0 ? 0.5 * (W[j * LD + i] + W[i * LD + j]) : column_dot(W + i * LD, W + j * LD, i > j ? i : j, n);
            a[j] = v;
        }
        TE_SYNC(c);                                                      // W is about to receive the Householder vectors
        tridiagonalize<N>(c, a, n, base + S.v, base + S.q, which ? base + S.d1 : base + S.d0,
                          which ? base + S.e1 : base + S.e0, which ? W : nullptr, LD, base + S.beta);
        g[which] = normalise_tridiagonal(c, which ? base + S.d1 : base + S.d0, which ? base + S.e1 : base + S.e0,
                                         which ? base + S.f1 : base + S.f0, n, base + S.tmp);
    }
    const double g0 = g[0], g1 = g[1];
    // 6. eigenvalues: every one from the end of the spectrum it is closer to.  theta in (0, 1] are the eigenvalues of
    //    T0 = C / g0, mu in (0, 1] those of T1 = C^-1 / g1;  lam = g0 theta = 1 / (g1 mu).  Split at the geometric
    //    mean of the two Gershgorin bounds: lam_s = sqrt(g0 / g1).
    const double theta_s = 1.0 / sqrt(g0 * g1);
    const int ns = sturm_count(base + S.d0, base + S.f0, n, theta_s);    // eigenvalues taken from C^-1
    const int nmu = ns > n_modes ? ns : n_modes;                         // mu's needed (shapes come from C^-1)
    // thread i < ns takes its eigenvalue from T1 (the n-1-i-th mu), the others from T0: ONE call with selected
    // pointers, so that a warp does not run the two cases one after the other
    double lam = 0.0, mu = 0.0;
    if (i < n) {
        const bool inv = i < ns;
        const double r = bisect(inv ? base + S.d1 : base + S.d0, inv ? base + S.f1 : base + S.f0, n, inv ? n - 1 - i : i,
                                0.0, 1.0);
        if (inv) { mu = r; lam = 1.0 / (g1 * r); }
        else lam = g0 * r;
    }
    if (nmu > ns) {                                                      // shapes wanted beyond the split: their mu as well
        if (i >= ns && i < nmu) mu = bisect(base + S.d1, base + S.f1, n, n - 1 - i, 0.0, 1.0);
    }
    if (i < n) {
        base[S.lam + i] = lam;
        base[S.mu + i] = mu;
    }
    TE_SYNC(c);
    int bad = 0;
    // shapes are eigenvectors of C^-1, accurate to eps * mu_max / gap: good for the modes near ITS top.  A structure
    // that wants shapes more than 1e5 below mu_max (tiny models asked for every mode) is declined: Jacobi takes it.
    if (n_modes > 0 && !(base[S.mu + n_modes - 1] > 1e-5 * base[S.mu])) bad = 1;
    if (i < n) {
        if (!(lam > 0.0 && lam < INFINITY)) bad = 1;
        if (i > 0 && !(lam >= base[S.lam + i - 1])) bad = 1;             // the two ends must interleave consistently
    }
    // 7. shapes: eigenvectors of T1 for the n_modes largest mu (inverse iteration), close ones re-orthogonalised
    double* Z = base + S.Z;
    if (i < n_modes) {
        double w[5 * NT];
        tridiag_eigvec(base + S.d1, base + S.e1, n, mu, 1e-300 + 2.3e-16, i, Z + i * NT, w);
        for (int r = n; r < NT; ++r) Z[i * NT + r] = 0.0;
        // residual of the tridiagonal eigenpair (|T1| <= 1): a failed inverse iteration declines the engine
        double res = 0.0;
        const double* d1 = base + S.d1;
        const double* e1 = base + S.e1;
        const double* z = Z + i * NT;
        for (int r = 0; r < n; ++r) {
            double t = (d1[r] - mu) * z[r];
            if (r > 0) t = fma(e1[r - 1], z[r - 1], t);
            if (r + 1 < n) t = fma(e1[r], z[r + 1], t);
            res = fmax(res, fabs(t));
        }
        if (!(res < 1e-11)) bad = 1;
    }
    TE_SYNC(c);
    for (int m = 1; m < n_modes; ++m) {
        // modified Gram-Schmidt against the previous vectors whose mu is within 1e-3 (relative): inverse iteration
        // leaves a close pair orthogonal only to eps / gap
        const double mum = base[S.mu + m];
        for (int p = m - 1; p >= 0; --p) {
            const double mup = base[S.mu + p];
            if (!(fabs(mup - mum) < 1e-3 * fabs(mup))) break;            // mu descends with the mode index: stop at the first far one
            const double dot = te_block_sum(c, Z[p * NT + i] * Z[m * NT + i]);
            Z[m * NT + i] -= dot * Z[p * NT + i];
            TE_SYNC(c);
            const double nn = te_block_sum(c, Z[m * NT + i] * Z[m * NT + i]);
            Z[m * NT + i] *= 1.0 / sqrt(nn);
            TE_SYNC(c);
        }
    }
    // 8. back-transformation u = H_0 ... H_{n-3} z: four threads per vector, rows r = part, part + 4, ...
    {
        const int m = i >> 2, part = i & 3;
        double z[NT / 4];
#pragma unroll
        for (int k = 0; k < NT / 4; ++k) z[k] = (m < n_modes) ? Z[m * NT + part + 4 * k] : 0.0;
        for (int k = n - 3; k >= 0; --k) {
            const double* v = W + k * LD;
            double dot = 0.0;
#pragma unroll
            for (int r = 0; r < NT / 4; ++r) dot = fma((part + 4 * r < n) ? v[part + 4 * r] : 0.0, z[r], dot);
            dot = te_quad_sum(c, dot) * base[S.beta + k];
#pragma unroll
            for (int r = 0; r < NT / 4; ++r) z[r] = fma(-dot, (part + 4 * r < n) ? v[part + 4 * r] : 0.0, z[r]);
        }
        TE_SYNC(c);
        if (m < n_modes) {
#pragma unroll
            for (int k = 0; k < NT / 4; ++k) Z[m * NT + part + 4 * k] = z[k];
        }
    }
    TE_SYNC(c);
    // 9. phi = Lm^-T u, |u| = 1  =>  phi^T M phi = 1;  sign: largest entry positive
    if (i < n_modes) {
        double* u = Z + i * NT;
        double nn = 0.0;
        for (int r = 0; r < n; ++r) nn = fma(u[r], u[r], nn);
        const double in = 1.0 / sqrt(nn);
        for (int r = 0; r < n; ++r) u[r] *= in;
        band_backward(s_Mb, s_invM, u, n, b);
        double best = 0.0, sign = 1.0;
        for (int r = 0; r < n; ++r) {
            const double av = fabs(u[r]);
            if (av > best) { best = av; sign = u[r] < 0.0 ? -1.0 : 1.0; }
        }
        if (!(best > 0.0 && best < INFINITY)) bad = 1;
        for (int r = 0; r < n; ++r) u[r] *= sign;
    }
    const double nbad = te_block_sum(c, (double)bad);
    TE_SYNC(c);
    return nbad > 0.0 ? -3 : 0;
}

}  // namespace te
}  // namespace bfe2
