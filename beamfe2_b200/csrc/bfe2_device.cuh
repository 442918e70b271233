// bfe2_device.cuh -- per-structure device routines shared by the unfused (K1..K4) and fused kernels.
//
// One thread block owns one structure variant; everything between the input loads and the output
// stores lives in shared memory / registers.  Reference citations are file:line in jkbgbr/BeamFE2.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <limits.h>

namespace bfe2 {

// ------------------------------------------------------------------------------------------------
// Device view of the topology plan (integer maps, built on the host in bfe2_plan.cpp-part of bfe2.cu)
// ------------------------------------------------------------------------------------------------
struct PlanDev {
    int nN, nE, sumdof, n, hbw;   // nodes, beams, 3*nN, free DOFs, half bandwidth (free numbering)
    int nslots;                   // (hbw+1)*n band slots
    int npad, LD;                 // Jacobi: even slot count >= n, leading dimension (rows, padded)
    const int32_t* conn;          // [nE*2]
    const int32_t* free_of_pos;   // [sumdof]
    const int32_t* pos_of_free;   // [n]
    const int32_t* slot_ptr;      // [nslots+1]
    const int32_t* contrib;       // [ncontrib]  e*36 + r*6 + c
    const int32_t* contrib_k;     // [ncontrib]  e*K7_PITCH + K number, sign in bit 31 (closed-form beam matrices)
    const int32_t* contrib_m;     // [ncontrib]  e*M12_PITCH + M number, sign in bit 31
    const int32_t* node_ptr;      // [nN+1]
    const int32_t* node_inc;      // [2*nE]      e*2 + end
    const uint8_t* lumped;        // [nE]
};

// Shared-memory carve-up (offsets in doubles from the dynamic smem base).  Region U is a union:
//   phase "elements/assembly": Ke[36 nE], Me[36 nE]
//   phase "static":            f[C n], ufull[C sumdof], ef[C nE 6]
//   phase "modal":             W[npad * LD]
struct SmemLayout {
    int coords, props, mass, geo, Kb, Mb, invK, invM, U, lam, order, total;
};

__host__ __device__ inline int imax3(int a, int b, int c) { return a > b ? (a > c ? a : c) : (b > c ? b : c); }

__host__ __device__ inline SmemLayout make_layout(int nN, int nE, int n, int hbw, int npad, int LD, int C) {
    SmemLayout L;
    int o = 0;
    L.coords = o; o += 2 * nN;
    L.props = o;  o += 4 * nE;
    L.mass = o;   o += 2 * nN;
    L.geo = o;    o += 3 * nE;
    L.Kb = o;     o += (hbw + 1) * n;
    L.Mb = o;     o += (hbw + 1) * n;
    L.invK = o;   o += n;
    L.invM = o;   o += n;
    o = (o + 1) & ~1;                                   // 16-byte align the big region
    L.U = o;
    // static phase: f, ufull, ef + room for the dense expansion of compact loads (f_nodal, r_local)
    o += imax3(72 * nE, C * (n + 6 * nN + 12 * nE), (npad > 0 && npad * LD < 640) ? 640 : npad * LD);
    L.lam = o;    o += npad;
    L.order = o;  o += (npad + 1) / 2 + 1;              // npad int32 + 2 flags, stored in double slots
    L.total = o;
    return L;
}

// Branch-free reciprocal / reciprocal square root: hardware seed (MUFU.RCP64H / RSQ64H, ~20 bits) and
// two Newton steps (rcp) or one third-order step (rsqrt).  Arguments here are finite, normal and positive-or-signed (rcp) -- no special cases.
__device__ __forceinline__ double rcp_fast(double a) {
    double x;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(a));
    double e = fma(-a, x, 1.0);
    x = fma(x, e, x);
    e = fma(-a, x, 1.0);
    return fma(x, e, x);
}
__device__ __forceinline__ double rsqrt_fast(double a) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    // one third-order step: e = 1 - a y^2,  y <- y (1 + e/2 + 3 e^2 / 8); seed error 2^-20 -> below 2^-58
    const double e = fma(-a * y, y, 1.0);
    const double p = fma(0.375, e, 0.5), ye = y * e;  // independent of each other: one level less than y (p e)
    return fma(ye, p, y);
}

// ------------------------------------------------------------------------------------------------
// Element level
// ------------------------------------------------------------------------------------------------

// Local stiffness, HermitianBeam_2D.py:284-297.  DOF order [ux_i, uy_i, rz_i, ux_j, uy_j, rz_j].
__device__ __forceinline__ void ke_local(double X[6][6], double E, double A, double I, double L) {
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int c = 0; c < 6; ++c) X[r][c] = 0.0;
    const double EA = E * A, EI = E * I;
    const double a = EA / L;
    const double b12 = 12.0 * EI / (L * L * L);
    const double b6 = 6.0 * EI / (L * L);
    const double b4 = 4.0 * EI / L;
    const double b2 = 2.0 * EI / L;
    X[0][0] = X[3][3] = a;
    X[0][3] = X[3][0] = -a;
    X[1][1] = X[4][4] = b12;
    X[1][4] = X[4][1] = -b12;
    X[1][2] = X[2][1] = X[1][5] = X[5][1] = b6;
    X[2][4] = X[4][2] = X[4][5] = X[5][4] = -b6;
    X[2][2] = X[5][5] = b4;
    X[2][5] = X[5][2] = b2;
}

// Local mass: consistent HermitianBeam_2D.py:233-249, lumped :225-231 (rotary terms m * 1e-10).
__device__ __forceinline__ void me_local(double X[6][6], double rho, double A, double L, bool lumped) {
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int c = 0; c < 6; ++c) X[r][c] = 0.0;
    if (lumped) {
        const double m = 0.5 * A * L * rho;
        X[0][0] = X[1][1] = X[3][3] = X[4][4] = m;
        X[2][2] = X[5][5] = m * 1e-10;
        return;
    }
    const double f = rho * A * L / 420.0;
    X[0][0] = X[3][3] = f * 140.0;
    X[0][3] = X[3][0] = f * 70.0;
    X[1][1] = X[4][4] = f * 156.0;
    X[1][2] = X[2][1] = f * (22.0 * L);
    X[1][4] = X[4][1] = f * 54.0;
    X[1][5] = X[5][1] = f * (-13.0 * L);
    X[2][2] = X[5][5] = f * (4.0 * L * L);
    X[2][4] = X[4][2] = f * (13.0 * L);
    X[2][5] = X[5][2] = f * (-3.0 * L * L);
    X[4][5] = X[5][4] = f * (-22.0 * L);
}

// X <- T X T^T with T = diag(R, R), R = [[c,-s,0],[s,c,0],[0,0,1]]  (helpers.py:43-69,
// HermitianBeam_2D.py:173-175).  Same left-to-right evaluation order as the reference: rows first.
__device__ __forceinline__ void rotate_to_global(double X[6][6], double c, double s) {
#pragma unroll
    for (int k = 0; k < 6; k += 3) {
#pragma unroll
        for (int j = 0; j < 6; ++j) {               // Y = T X  (rows k, k+1)
            const double a = X[k][j], b = X[k + 1][j];
            X[k][j] = c * a - s * b;
            X[k + 1][j] = s * a + c * b;
        }
    }
#pragma unroll
    for (int k = 0; k < 6; k += 3) {
#pragma unroll
        for (int i = 0; i < 6; ++i) {               // Z = Y T^T (columns k, k+1)
            const double a = X[i][k], b = X[i][k + 1];
            X[i][k] = c * a - s * b;
            X[i][k + 1] = s * a + c * b;
        }
    }
}

// Geometry of one beam: length HermitianBeam_2D.py:98-100; direction cosines of :114-120
// (the reference takes cos/sin of arctan2(dy,dx); dx/L, dy/L are the same numbers to 1 ulp).
__device__ __forceinline__ void beam_geometry(double xi, double yi, double xj, double yj,
                                              double& L, double& c, double& s) {
    const double dx = xj - xi, dy = yj - yi;
    L = sqrt((xi - xj) * (xi - xj) + (yi - yj) * (yi - yj));
    c = dx / L;
    s = dy / L;
}

// ------------------------------------------------------------------------------------------------
// Block-level stages.  `tid`/`nt` = thread index / count inside the block.
// ------------------------------------------------------------------------------------------------

// Stages E + A: beam matrices in global axes and collision-free assembly of the condensed band matrices.  Every
// band slot (d, j) is owned by one thread which adds its contributions in beam-list order -- the summation order
// of helpers.py:19-39 -- so no atomics and a deterministic result.  Supported rows/columns are simply not in
// the plan (Structure.condense, Structure.py:183-191); nodal masses go on the diagonal (Structure.py:198-209;
// mx->ux, my->uy, no rotary term :289-313).  The beam matrices are used in closed form: the rotated
// stiffness T k T^T of a beam has 7 distinct numbers, the rotated consistent mass 12,
//   K: p = a c^2 + b12 s^2, q = (a - b12) c s, r = a s^2 + b12 c^2, t = b6 s, w = b6 c, b4, b2
//   M: f = rho A L / 420;  ii/jj block: f (140 c^2 + 156 s^2), -16 f c s, f (140 s^2 + 156 c^2), 22 L f s, 22 L f c,
//      4 L^2 f;  ij block: f (70 c^2 + 54 s^2), 16 f c s, f (70 s^2 + 54 c^2), 13 L f s, 13 L f c, -3 L^2 f
// and the plan says which of them (and with which sign) every band contribution is (contrib_k / contrib_m).
// Pitch (doubles) of a beam's 7 stiffness / 12 mass numbers in shared memory: odd, so that the plan-driven gathers
// of neighbouring beams fall into different banks (a pitch of 8 puts every second beam on the same one).
constexpr int K7_PITCH = 7, M12_PITCH = 13;

__device__ __forceinline__ void beam_numbers(double xi, double yi, double xj, double yj, double E, double Ar, double I,
                                             double rho, bool want_mass, bool lumped, double* k7, double* m12,
                                             double& L, double& cs, double& sn) {
    const double dx = xj - xi, dy = yj - yi;
    const double d2 = fma(dx, dx, dy * dy);
    const double iL = rsqrt_fast(d2);
    L = d2 * iL;
    cs = dx * iL;
    sn = dy * iL;
    const double a = (E * Ar) * iL, eil = (E * I) * iL;
    const double b6 = 6.0 * eil * iL, b12 = 2.0 * b6 * iL;
    k7[0] = fma(a * cs, cs, (b12 * sn) * sn);
    k7[1] = ((a - b12) * cs) * sn;
    k7[2] = fma(a * sn, sn, (b12 * cs) * cs);
    k7[3] = b6 * sn;
    k7[4] = b6 * cs;
    k7[5] = 4.0 * eil;
    k7[6] = 2.0 * eil;
    if (!want_mass) return;
    if (lumped) {                                     // HermitianBeam_2D.py:225-231: m on translations, m 1e-10 on rotations
        const double m = 0.5 * Ar * L * rho;
        m12[0] = m; m12[1] = 0.0; m12[2] = m; m12[3] = 0.0; m12[4] = 0.0; m12[5] = m * 1e-10;
#pragma unroll
        for (int k = 6; k < 12; ++k) m12[k] = 0.0;
    } else {
        const double f = rho * Ar * L / 420.0, fl = f * L;
        const double cc = cs * cs, ss = sn * sn, fcs = (f * cs) * sn;
        m12[0] = f * fma(140.0, cc, 156.0 * ss);
        m12[1] = -16.0 * fcs;
        m12[2] = f * fma(140.0, ss, 156.0 * cc);
        m12[3] = (22.0 * fl) * sn;
        m12[4] = (22.0 * fl) * cs;
        m12[5] = (4.0 * fl) * L;
        m12[6] = f * fma(70.0, cc, 54.0 * ss);
        m12[7] = 16.0 * fcs;
        m12[8] = f * fma(70.0, ss, 54.0 * cc);
        m12[9] = (13.0 * fl) * sn;
        m12[10] = (13.0 * fl) * cs;
        m12[11] = (-3.0 * fl) * L;
    }
}

__device__ __forceinline__ void stage_elements_cf(const PlanDev& P, const double* s_coords, const double* s_props,
                                                  double* s_geo, double* s_k7, double* s_m12, bool want_mass,
                                                  int tid, int nt) {
    for (int e = tid; e < P.nE; e += nt) {
        const int ni = P.conn[2 * e], nj = P.conn[2 * e + 1];
        double L, cs, sn;
        beam_numbers(s_coords[2 * ni], s_coords[2 * ni + 1], s_coords[2 * nj], s_coords[2 * nj + 1], s_props[4 * e],
                     s_props[4 * e + 1], s_props[4 * e + 2], s_props[4 * e + 3], want_mass, P.lumped[e] != 0,
                     s_k7 + K7_PITCH * e, s_m12 + M12_PITCH * e, L, cs, sn);
        s_geo[3 * e] = L;
        s_geo[3 * e + 1] = cs;
        s_geo[3 * e + 2] = sn;
    }
}

__device__ __forceinline__ void stage_assemble_cf(const PlanDev& P, const double* s_k7, const double* s_m12,
                                                  const double* s_mass, double* s_Kb, double* s_Mb, bool want_mass,
                                                  int tid, int nt) {
    for (int slot = tid; slot < P.nslots; slot += nt) {
        const int p0 = P.slot_ptr[slot], p1 = P.slot_ptr[slot + 1];
        double k = 0.0, m = 0.0;
        for (int p = p0; p < p1; ++p) {
            const int ck = P.contrib_k[p];
            const double vk = s_k7[ck & 0x7fffffff];
            k += (ck < 0) ? -vk : vk;
            if (want_mass) {
                const int cm = P.contrib_m[p];
                const double vm = s_m12[cm & 0x7fffffff];
                m += (cm < 0) ? -vm : vm;
            }
        }
        s_Kb[slot] = k;
        if (want_mass) {
            if (slot < P.n && s_mass != nullptr) {       // d == 0: diagonal
                const int pos = P.pos_of_free[slot];
                const int dof = pos % 3;
                if (dof < 2) m += s_mass[2 * (pos / 3) + dof];
            }
            s_Mb[slot] = m;
        }
    }
}

// Stage E': geometry only (the static kernel needs L, c, s for end forces but not the matrices).
__device__ __forceinline__ void stage_geometry(const PlanDev& P, const double* s_coords, double* s_geo, int tid, int nt) {
    for (int e = tid; e < P.nE; e += nt) {
        const int ni = P.conn[2 * e], nj = P.conn[2 * e + 1];
        double L, c, s;
        beam_geometry(s_coords[2 * ni], s_coords[2 * ni + 1], s_coords[2 * nj], s_coords[2 * nj + 1], L, c, s);
        s_geo[3 * e] = L;
        s_geo[3 * e + 1] = c;
        s_geo[3 * e + 2] = s;
    }
}

// In-place banded Cholesky by ONE WARP, any band width.  Ab[d*n + j] = A[j+d][j], d = 0..b.  Returns 0 or the
// 1-based column of the first non-positive pivot (uniform across the warp).  Also writes 1/L[j][j].
//   Root-free elimination first: in column j lane q (1 <= q <= m) owns column j+q of the trailing block and
//   subtracts A[j+p][j] A[j+q][j] / pivot from its rows p = q..m (consecutive lanes touch consecutive addresses);
//   column j itself is only read, so ONE __syncwarp per column suffices.  One lane-parallel pass then scales the
//   columns into the Cholesky factor.
__device__ __forceinline__ int band_cholesky_warp(double* Ab, double* invdiag, int n, int b, int lane) {
    // invdiag[j] holds the ORIGINAL diagonal until the scaling pass: a pivot that has lost all its digits
    // (<= 64 eps * original) is reported as non-positive, i.e. numerically singular.
    for (int j = lane; j < n; j += 32) invdiag[j] = Ab[j];
    __syncwarp();
    for (int j = 0; j < n; ++j) {
        const double piv = Ab[j];
        if (!(piv > 1.4210854715202004e-14 * invdiag[j] && piv < INFINITY)) return j + 1;
        const double ip = rcp_fast(piv);
        const int m = min(b, n - 1 - j);
        for (int q = 1 + lane; q <= m; q += 32) {
            const double* c = Ab + q * n + j;        // c[d n] = A[j+q+d][j]
            double* t = Ab + j + q;                  // t[d n] = A[j+q+d][j+q]
            const double cq = c[0] * ip;
            for (int d = 0; d <= m - q; ++d) t[d * n] = fma(-c[d * n], cq, t[d * n]);
        }
        __syncwarp();
    }
    for (int j = lane; j < n; j += 32) {
        const double piv = Ab[j];
        const double r = rsqrt_fast(piv);
        invdiag[j] = r;
        Ab[j] = piv * r;
    }
    __syncwarp();
    for (int d = 1; d <= b; ++d)
        for (int j = lane; j < n - d; j += 32) Ab[d * n + j] *= invdiag[j];
    __syncwarp();
    return 0;
}

// ---- narrow bands (hbw <= HB <= 8): substitutions by ONE THREAD per right-hand side with the active window in
//      registers: per row one fma + one mul on the critical path, no shared-memory round trip, no sync.
// x <- (L L^T)^-1 x (FWD && BWD) or x <- L^-T x (BWD only); xin/xout may alias; stride = distance of rows
// EXACT: the band is exactly HB wide and its slots outside the matrix hold finite numbers (zeros from the
// assembly): no bounds tests -- such entries only ever multiply the zeros the window starts with.
template <int HB, bool FWD, bool BWD, bool EXACT = false>
__device__ __forceinline__ void band_subst_thread(const double* Lb, const double* invdiag, double* x, int stride,
                                                  int n, int b, int first = 0) {
    if (FWD) {
        double win[HB];                              // win[d-1] = x[j-d]
#pragma unroll
        for (int d = 0; d < HB; ++d) win[d] = 0.0;
        for (int j = first; j < n; ++j) {
            double far = x[j * stride];
#pragma unroll
            for (int d = 2; d <= HB; ++d) {
                const double lc = (EXACT || (d <= b && j - d >= first)) ? Lb[d * n + (j - d)] : 0.0;
                far = fma(-lc, win[d - 1], far);
            }
            const double l1 = (EXACT || (1 <= b && j - 1 >= first)) ? Lb[n + (j - 1)] : 0.0;
            const double v = fma(-l1, win[0], far) * invdiag[j];
#pragma unroll
            for (int d = HB - 1; d >= 1; --d) win[d] = win[d - 1];
            win[0] = v;
            x[j * stride] = v;
        }
    }
    if (BWD) {
        double win[HB];                              // win[d-1] = x[j+d]
#pragma unroll
        for (int d = 0; d < HB; ++d) win[d] = 0.0;
        for (int j = n - 1; j >= first; --j) {
            double far = x[j * stride];
#pragma unroll
            for (int d = 2; d <= HB; ++d) {
                const double lc = (EXACT || (d <= b && j + d < n)) ? Lb[d * n + j] : 0.0;
                far = fma(-lc, win[d - 1], far);
            }
            const double l1 = (EXACT || (1 <= b && j + 1 < n)) ? Lb[n + j] : 0.0;
            const double v = fma(-l1, win[0], far) * invdiag[j];
#pragma unroll
            for (int d = HB - 1; d >= 1; --d) win[d] = win[d - 1];
            win[0] = v;
            x[j * stride] = v;
        }
    }
}

// Dispatch on the band width; `clean` = the caller guarantees zeros in the band slots outside the matrix.
template <bool FWD, bool BWD>
__device__ __forceinline__ void band_subst(const double* Lb, const double* invdiag, double* x, int stride, int n, int b,
                                           int first, bool clean) {
    if (b == 5 && clean) band_subst_thread<5, FWD, BWD, true>(Lb, invdiag, x, stride, n, b, first);
    else if (b <= 5) band_subst_thread<5, FWD, BWD>(Lb, invdiag, x, stride, n, b, first);
    else if (b == 8 && clean) band_subst_thread<8, FWD, BWD, true>(Lb, invdiag, x, stride, n, b, first);
    else band_subst_thread<8, FWD, BWD>(Lb, invdiag, x, stride, n, b, first);
}

// Row-major band ([d][n], the layout of the fused kernel): the factorisation by the G lanes of a group / warp.
// Root-free elimination first: per column only the trailing update  A[j+p][j+q] -= A[j+p][j] A[j+q][j] / pivot
// (lane t owns pair (p, q) number t) and ONE __syncwarp -- no scaled column has to be stored and re-read inside
// the column loop; then one lane-parallel pass turns the unscaled columns into the Cholesky factor
// (L[i][j] = A[i][j] / sqrt(pivot_j), invdiag[j] = 1 / L[j][j]).  invdiag holds the ORIGINAL diagonal until that
// pass (the relative pivot test).  Returns 0 or the 1-based column of the first bad pivot, uniform per group; a
// failed group keeps running in lockstep without storing.
template <int G, int HB>
__device__ __forceinline__ int band_cholesky_group(double* Ab, double* invdiag, int n, int b, int g) {
    for (int j = g; j < n; j += G) invdiag[j] = Ab[j];
    constexpr int NP = HB * (HB + 1) / 2, ROUNDS = (NP + G - 1) / G;
    int off_t[ROUNDS], off_p[ROUNDS], off_q[ROUNDS], need[ROUNDS];
#pragma unroll
    for (int k = 0; k < ROUNDS; ++k) {
        const int t = g + k * G;
        int p = 1;
        while (p * (p + 1) / 2 <= t) ++p;            // t = p (p - 1) / 2 + (q - 1), 1 <= q <= p
        const int q = t - p * (p - 1) / 2 + 1;
        need[k] = (t < NP) ? p : HB + 1;             // the pair exists in a column with m >= p
        off_p[k] = p * n;
        off_q[k] = q * n;
        off_t[k] = (p - q) * n + q;
    }
    __syncwarp();
    int info = 0;
    for (int col = 0; col < n; ++col) {
        const int m = min(b, n - 1 - col);
        const double piv = Ab[col];
        const double orig = invdiag[col];
        if (info == 0 && !(piv > 1.4210854715202004e-14 * orig && piv < INFINITY)) info = col + 1;
        const double ip = rcp_fast(piv);
        if (info == 0) {
#pragma unroll
            for (int k = 0; k < ROUNDS; ++k)
                if (need[k] <= m) {
                    double* t = Ab + off_t[k] + col;
                    *t = fma(-(Ab[off_p[k] + col] * ip), Ab[off_q[k] + col], *t);
                }
        }
        __syncwarp();
    }
    if (info == 0) {
        for (int j = g; j < n; j += G) {
            const double piv = Ab[j];
            const double r = rsqrt_fast(piv);
            invdiag[j] = r;
            Ab[j] = piv * r;
        }
    }
    __syncwarp();
    if (info == 0) {
        for (int d = 1; d <= b; ++d)
            for (int j = g; j < n - d; j += G) Ab[d * n + j] *= invdiag[j];
    }
    __syncwarp();
    return info;
}

// ---- column-major band for the group kernel: column j is W = HB + 1 (rounded up to even) consecutive doubles
//      Cc[j * W + d] = A[j + d][j], 16-byte aligned, entries outside the band or the matrix are zero.  After the
//      factorisation slot d = 0 holds 1 / L[j][j] and d >= 1 holds L[j + d][j]: one substitution row is W / 2
//      16-byte loads (all lanes of a group read the same column: broadcasts).
template <int HB>
struct BandCm { static constexpr int W = (HB + 2) & ~1; };

// Root-free elimination by the G lanes of a group (lane t owns trailing-update pair (p, q) number t), one
// __syncwarp per column, then one lane-per-column pass that scales the column and leaves 1 / L[j][j] in slot 0.
// orig[j] receives the original diagonal (relative pivot test).  Returns 0 or the 1-based bad column.
template <int G, int HB>
__device__ __forceinline__ int band_cholesky_group_cm(double* Cc, double* orig, int n, int g) {
    constexpr int W = BandCm<HB>::W;
    for (int j = g; j < n; j += G) orig[j] = Cc[j * W];
    constexpr int NP = HB * (HB + 1) / 2, ROUNDS = (NP + G - 1) / G;
    int off_t[ROUNDS], off_p[ROUNDS], off_q[ROUNDS], need[ROUNDS];
#pragma unroll
    for (int k = 0; k < ROUNDS; ++k) {
        const int t = g + k * G;
        int p = 1;
        while (p * (p + 1) / 2 <= t) ++p;            // t = p (p - 1) / 2 + (q - 1), 1 <= q <= p
        const int q = t - p * (p - 1) / 2 + 1;
        need[k] = (t < NP) ? p : HB + 1;             // the pair exists in a column with m >= p
        off_p[k] = p;
        off_q[k] = q;
        off_t[k] = q * W + (p - q);                  // A[j+p][j+q] = Cc[(j+q) W + (p-q)]
    }
    __syncwarp();
    int info = 0;
    for (int col = 0; col < n; ++col) {
        const int m = min(HB, n - 1 - col);
        double* cj = Cc + col * W;
        const double piv = cj[0];
        if (info == 0 && !(piv > 1.4210854715202004e-14 * orig[col] && piv < INFINITY)) info = col + 1;
        const double ip = rcp_fast(piv);
        if (info == 0) {
#pragma unroll
            for (int k = 0; k < ROUNDS; ++k)
                if (need[k] <= m) cj[off_t[k]] = fma(-(cj[off_p[k]] * ip), cj[off_q[k]], cj[off_t[k]]);
        }
        __syncwarp();
    }
    if (info == 0) {
        for (int j = g; j < n; j += G) {
            double2* c2 = reinterpret_cast<double2*>(Cc + j * W);
            double2 v = c2[0];
            const double r = rsqrt_fast(v.x);
            v.x = r;
            v.y *= r;
            c2[0] = v;
#pragma unroll
            for (int h = 1; h < W / 2; ++h) {
                double2 w = c2[h];
                w.x *= r;
                w.y *= r;
                c2[h] = w;
            }
        }
    }
    __syncwarp();
    return info;
}

// x <- (L L^T)^-1 x with the column-major factor, one thread per right-hand side.  Both passes walk the columns:
// forward right-looking (y_j = x_j / L_jj, then x_{j+d} -= L[j+d][j] y_j in a register window), backward
// left-looking.  Reads of x behind row n - 1 only ever meet zero entries of L.
template <int HB>
__device__ __forceinline__ void band_subst_cm(const double* Cc, double* x, int n) {
    constexpr int W = BandCm<HB>::W;
    double win[HB + 1];
#pragma unroll
    for (int d = 0; d <= HB; ++d) win[d] = x[d];
    for (int j = 0; j < n; ++j) {
        double l[W];
        const double2* c2 = reinterpret_cast<const double2*>(Cc + j * W);
#pragma unroll
        for (int h = 0; h < W / 2; ++h) { const double2 v = c2[h]; l[2 * h] = v.x; l[2 * h + 1] = v.y; }
        const double y = win[0] * l[0];
        x[j] = y;
#pragma unroll
        for (int d = 1; d <= HB; ++d) win[d - 1] = fma(-l[d], y, win[d]);
        win[HB] = x[j + HB + 1];
    }
#pragma unroll
    for (int d = 0; d <= HB; ++d) win[d] = 0.0;      // win[d - 1] = u[j + d]
    for (int j = n - 1; j >= 0; --j) {
        double l[W];
        const double2* c2 = reinterpret_cast<const double2*>(Cc + j * W);
#pragma unroll
        for (int h = 0; h < W / 2; ++h) { const double2 v = c2[h]; l[2 * h] = v.x; l[2 * h + 1] = v.y; }
        double acc = x[j];
#pragma unroll
        for (int d = HB; d >= 1; --d) acc = fma(-l[d], win[d - 1], acc);
        const double u = acc * l[0];
#pragma unroll
        for (int d = HB - 1; d >= 1; --d) win[d] = win[d - 1];
        win[0] = u;
        x[j] = u;
    }
}

// Forward + back substitution with the band factor, one thread per right-hand side (in place).
__device__ __forceinline__ void band_solve_thread(const double* Lb, const double* invdiag, double* x, int n, int b) {
    for (int j = 0; j < n; ++j) {                    // L y = f
        double acc = x[j];
        const int m = min(b, j);
        for (int d = 1; d <= m; ++d) acc -= Lb[d * n + (j - d)] * x[j - d];
        x[j] = acc * invdiag[j];
    }
    for (int j = n - 1; j >= 0; --j) {               // L^T u = y
        double acc = x[j];
        const int m = min(b, n - 1 - j);
        for (int d = 1; d <= m; ++d) acc -= Lb[d * n + j] * x[j + d];
        x[j] = acc * invdiag[j];
    }
}

// Back substitution only: L^T x = w (in place), one thread.
__device__ __forceinline__ void band_backsolve_thread(const double* Lb, const double* invdiag, double* x, int n, int b) {
    for (int j = n - 1; j >= 0; --j) {
        double acc = x[j];
        const int m = min(b, n - 1 - j);
        for (int d = 1; d <= m; ++d) acc -= Lb[d * n + j] * x[j + d];
        x[j] = acc * invdiag[j];
    }
}

// Stage S: static solve for C load cases + end forces + reactions, results to global memory.
//   f = f_nodal + sum_e T(alpha_e) r_local_e        (Structure.py:244-287, Loads.py:100-112)
//   u = K^-1 f on the free DOFs, 0 at supports      (Solver.py:76-87 with condensation)
//   endforces = Ke u_local - r_local                (Results.py:48-79, HermitianBeam_2D.py:336-351)
//   reactions = sum_e R(alpha_e) endforces - f_nodal (Results.py:92-127)
template <bool WARP_SCOPE>
__device__ __forceinline__ void scope_sync() {
    if (WARP_SCOPE) __syncwarp();
    else __syncthreads();
}

template <bool WARP_SCOPE = false>
__device__ __forceinline__ void stage_static(const PlanDev& P, int C, const double* s_props, const double* s_geo,
                                             const double* s_Kb, const double* s_invK,
                                             double* s_f, double* s_ufull, double* s_ef,
                                             const double* g_f_nodal, const double* g_r_local,
                                             double* g_u, double* g_ef, double* g_react, int tid, int nt,
                                             bool clean = false, const int32_t* fixed_pos = nullptr, int n_fixed = 0) {
    // fixed_pos != nullptr: compact outputs -- u on the free DOFs only ([C, n]), reactions at the supported
    // positions only ([C, n_fixed]; Results.py:129-143 reads them nowhere else)
    const int n = P.n, sumdof = P.sumdof, nE = P.nE;
    // 1. right-hand sides in free numbering
    for (int t = tid; t < C * n; t += nt) {
        const int c = t / n, j = t % n;
        const int pos = P.pos_of_free[j];
        double v = g_f_nodal[c * sumdof + pos];
        if (g_r_local != nullptr) {
            const int node = pos / 3, k = pos % 3;
            for (int q = P.node_ptr[node]; q < P.node_ptr[node + 1]; ++q) {
                const int inc = P.node_inc[q];
                const int e = inc >> 1, end = inc & 1;
                const double* r = g_r_local + ((size_t)c * nE + e) * 6 + 3 * end;
                const double cs = s_geo[3 * e + 1], sn = s_geo[3 * e + 2];
                if (k == 0) v += cs * r[0] - sn * r[1];
                else if (k == 1) v += sn * r[0] + cs * r[1];
                else v += r[2];
            }
        }
        s_f[c * n + j] = v;
    }
    scope_sync<WARP_SCOPE>();
    // 2. one thread per load case: L L^T u = f
    for (int c = tid; c < C; c += nt) {
        if (P.hbw <= 8) band_subst<true, true>(s_Kb, s_invK, s_f + c * n, 1, n, P.hbw, 0, clean);
        else band_solve_thread(s_Kb, s_invK, s_f + c * n, n, P.hbw);
    }
    scope_sync<WARP_SCOPE>();
    // 3. expand to all DOFs (exact zeros at supports), write u
    for (int t = tid; t < C * sumdof; t += nt) {
        const int c = t / sumdof, pos = t % sumdof;
        const int j = P.free_of_pos[pos];
        const double v = (j >= 0) ? s_f[c * n + j] : 0.0;
        s_ufull[t] = v;
        if (fixed_pos == nullptr) g_u[t] = v;
    }
    if (fixed_pos != nullptr)
        for (int t = tid; t < C * n; t += nt) g_u[t] = s_f[t];
    scope_sync<WARP_SCOPE>();
    if (g_ef == nullptr && g_react == nullptr) return;
    // 4. local end forces per (load case, beam)
    for (int t = tid; t < C * nE; t += nt) {
        const int c = t / nE, e = t % nE;
        const int ni = P.conn[2 * e], nj = P.conn[2 * e + 1];
        const double L = s_geo[3 * e], cs = s_geo[3 * e + 1], sn = s_geo[3 * e + 2];
        const double* ui = s_ufull + c * sumdof + 3 * ni;
        const double* uj = s_ufull + c * sumdof + 3 * nj;
        // u_local = R(-alpha) u  (Results.py:63-71)
        const double d0 = cs * ui[0] + sn * ui[1], d1 = -sn * ui[0] + cs * ui[1], d2 = ui[2];
        const double d3 = cs * uj[0] + sn * uj[1], d4 = -sn * uj[0] + cs * uj[1], d5 = uj[2];
        const double E = s_props[4 * e], A = s_props[4 * e + 1], I = s_props[4 * e + 2];
        const double EA = E * A, EI = E * I;
        const double a = EA / L, b12 = 12.0 * EI / (L * L * L), b6 = 6.0 * EI / (L * L);
        const double b4 = 4.0 * EI / L, b2 = 2.0 * EI / L;
        double q[6];
        q[0] = a * d0 - a * d3;
        q[1] = b12 * d1 + b6 * d2 - b12 * d4 + b6 * d5;
        q[2] = b6 * d1 + b4 * d2 - b6 * d4 + b2 * d5;
        q[3] = -a * d0 + a * d3;
        q[4] = -b12 * d1 - b6 * d2 + b12 * d4 - b6 * d5;
        q[5] = b6 * d1 + b2 * d2 - b6 * d4 + b4 * d5;
        if (g_r_local != nullptr) {
            const double* r = g_r_local + ((size_t)c * nE + e) * 6;
#pragma unroll
            for (int k = 0; k < 6; ++k) q[k] -= r[k];
        }
#pragma unroll
        for (int k = 0; k < 6; ++k) s_ef[t * 6 + k] = q[k];
        if (g_ef != nullptr) {
#pragma unroll
            for (int k = 0; k < 6; ++k) g_ef[(size_t)t * 6 + k] = q[k];
        }
    }
    if (g_react == nullptr) return;
    scope_sync<WARP_SCOPE>();
    // 5. global reactions per (load case, position)
    const int nr = fixed_pos ? n_fixed : sumdof;
    for (int t = tid; t < C * nr; t += nt) {
        const int c = t / nr, pos = fixed_pos ? fixed_pos[t % nr] : t % nr;
        const int node = pos / 3, k = pos % 3;
        double v = 0.0;
        for (int q = P.node_ptr[node]; q < P.node_ptr[node + 1]; ++q) {
            const int inc = P.node_inc[q];
            const int e = inc >> 1, end = inc & 1;
            const double* f = s_ef + ((size_t)c * nE + e) * 6 + 3 * end;
            const double cs = s_geo[3 * e + 1], sn = s_geo[3 * e + 2];
            if (k == 0) v += cs * f[0] - sn * f[1];
            else if (k == 1) v += sn * f[0] + cs * f[1];
            else v += f[2];
        }
        g_react[t] = v - g_f_nodal[c * sumdof + pos];
    }
}

// ------------------------------------------------------------------------------------------------
// Stage M: generalized eigenproblem K phi = lam M phi  (scipy.linalg.eigh(K, M), Solver.py:39).
//
//   M = Lm Lm^T (band Cholesky, done by the caller); standard form C = Lm^-1 K Lm^-T (stage_form_a);
//   diagonally pivoted Cholesky C = Z Z^T (Demmel-Veselic / Drmac preconditioner); the cyclic Jacobi method is
//   then applied ONE-SIDED to the factor: Z V = U Sigma, rotating column pairs of Z until they are mutually
//   orthogonal (Veselic/Hari: the small eigenvalues keep their relative accuracy).  At convergence
//   lam_k = |z_k|^2, u_k = z_k / |z_k| is an eigenvector of C and  phi_k = Lm^-T u_k  satisfies
//   K phi = lam M phi  with  phi^T M phi = 1  (scipy's normalisation).
//
//   Ordering: odd-even transposition on a line of npad (even) positions; when n is odd one position holds a
//   zero dummy column.  n <= 64: columns in registers, two per lane, the moving column travels by warp shuffle
//   (stage_jacobi_regs); 64 < n <= 116: LPP lanes per column pair, the moving column travels through a
//   double-buffered exchange area in shared memory (stage_jacobi_oe).
// ------------------------------------------------------------------------------------------------
#define BFE2_JACOBI_TOL 1.0e-15
#define BFE2_JACOBI_TAU_WIDE 2.0e-9        // 2 n tau^2 < tol for n <= 116 (the shared-memory variant)
#define BFE2_JACOBI_TAU 2.5e-9             // 2 n tau^2 < tol for n <= 64 (the register path)

// ------------------------------------------------------------------------------------------------
// Stage M, register-resident variant (n <= 64): the same one-sided Jacobi, but the columns never
// leave the register file during the sweeps.
//
//   Layout: TWO warps per structure; warp h owns rows [h*RPL, (h+1)*RPL) of EVERY column, lane k owns
//   the columns at line positions 2k (A) and 2k+1 (B).  Ordering: odd-even transposition on a line of
//   npad (even) positions -- step s pairs positions (i, i+1), i = s mod 2, rotates them and
//   interchanges them; after npad steps every pair has met exactly once.  With the interchange folded
//   into the neighbour exchange the per-step data movement is ONE cyclic warp shuffle of one column:
//       even step: rotate (A, B) in place;  B <- B of lane k+1 (cyclic)
//       odd  step: rotate (A, B) in place;  A <- A of lane k-1 (cyclic);  the last lane, which holds
//                  the two idle end positions, exchanges them ("rotation" c = 0, s = -1)
//   Column norms travel with the columns and are updated by a' = a - t g, b' = b + t g (refreshed from
//   the data at the start of every sweep), so a step costs one dot product, not three.  The two
//   row-halves of the dot product meet through 2 x 8 bytes of shared memory per lane and one
//   __syncthreads per step; both warps then compute bit-identical rotation parameters.
// ------------------------------------------------------------------------------------------------
// shared-memory access through 32-bit shared-window addresses (saves the generic->shared conversion that
// the compiler otherwise re-materialises around every barrier)
__device__ __forceinline__ void sts_f64(unsigned addr, double v) {
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}
__device__ __forceinline__ double lds_f64(unsigned addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
    return v;
}

// One step of the odd-even ordering, software-pipelined: the rotation of the resident pair, the cyclic
// shuffle of the moving column and the dot product of the NEXT pairing are fused in one straight-line loop
// over the rows, so that a single warp keeps the FP64 pipe (rotation + dot FMAs) and the shuffle path busy
// at the same time.  On entry g is the dot product of the resident pair; on exit the partial dot product of
// the next pairing has been stored to shared memory (pg_mine).
template <int RPL, bool ODD>
__device__ __forceinline__ void jacobi_regs_step(double (&A)[RPL], double (&B)[RPL], double& nA, double& nB,
                                                 int& unsettled, double g, unsigned pg_mine, bool act, bool last, int src) {
    // rotate when |g| > tol * sqrt(nA nB)  <=>  g^2 > tol^2 nA nB
    const double gg = g * g, pab = nA * nB;
    const bool live = act && !(ODD && last);
    const bool rot = live && (gg > (BFE2_JACOBI_TOL * BFE2_JACOBI_TOL) * pab);
    // A sweep whose cosines all start below tau and whose rotations all have |sin| below sigma leaves cosines
    // below tol + 2 n tau sigma (each later rotation of a column moves a settled cosine by at most |sin| * cos):
    // with 2 n tau sigma < tol it is the last one, no empty checking sweep needed (the quadratic-convergence
    // exit of LAPACK's dgesvj, stated on both quantities).
    if (live && gg > (BFE2_JACOBI_TAU * BFE2_JACOBI_TAU) * pab) unsettled = 1;
    double c = 1.0, s = 0.0;
    if (ODD && last && act) {                         // exchange the two idle end columns (the sign flip is
        c = 0.0; s = -1.0;                            // immaterial: columns are defined up to sign)
        const double tmp = nA; nA = nB; nB = tmp;
    }
    if (rot) {
        // tan(2 theta) = h / d, h = 2g, d = nB - nA, |theta| <= pi/4:
        //   cos(2 theta) = |d| / r, sin(2 theta) = sign(d) h / r, r = sqrt(d^2 + h^2)
        //   c^2 = (1 + cos 2theta) / 2,  s = sin(2 theta) / (2 c),  t = s / c
        const double d = nB - nA, hh = 2.0 * g;
        const double ir = rsqrt_fast(fma(d, d, hh * hh));
        const double hs = (d < 0.0) ? -hh : hh;       // sign(d) h
        const double c2 = fma(0.5 * fabs(d), ir, 0.5);
        const double ic = rsqrt_fast(c2);             // 1 / c
        c = c2 * ic;
        s = (0.5 * hs * ir) * ic;
        const double tg = (s * ic) * g;               // t g
        nA -= tg;
        nB += tg;
        if (fabs(s) > BFE2_JACOBI_TAU) unsettled = 1;
    }
    constexpr int NACC = 2;      // accumulators of the fused dot product (measured: 2 > 4 > 8, register pressure)
    double acc[NACC];
#pragma unroll
    for (int k = 0; k < NACC; ++k) acc[k] = 0.0;
    if (__any_sync(0xffffffffu, rot || (ODD && last && act))) {
#pragma unroll
        for (int r = 0; r < RPL; ++r) {
            const double sb = s * B[r], cb = c * B[r];
            const double nb = fma(s, A[r], cb);
            const double na = fma(c, A[r], -sb);
            if (!ODD) { A[r] = na; B[r] = __shfl_sync(0xffffffffu, nb, src); }
            else { B[r] = nb; A[r] = __shfl_sync(0xffffffffu, na, src); }
            acc[r & (NACC - 1)] = fma(A[r], B[r], acc[r & (NACC - 1)]);
        }
    } else {                                          // nobody in this warp rotates: move and multiply only
#pragma unroll
        for (int r = 0; r < RPL; ++r) {
            if (!ODD) B[r] = __shfl_sync(0xffffffffu, B[r], src);
            else A[r] = __shfl_sync(0xffffffffu, A[r], src);
            acc[r & (NACC - 1)] = fma(A[r], B[r], acc[r & (NACC - 1)]);
        }
    }
    if (!ODD) nB = __shfl_sync(0xffffffffu, nB, src);
    else nA = __shfl_sync(0xffffffffu, nA, src);
    sts_f64(pg_mine, acc[0] + acc[1]);
}

// Diagonally pivoted Cholesky of the symmetric matrix held column-wise in the Jacobi register layout
// (lane k: columns 2k, 2k+1; warp h: rows h RPL .. h RPL + RPL - 1):  C = Z Z^T, column j of Z replacing column j
// of C in place, rows kept in their original order.  Z is the Demmel-Veselic / Drmac preconditioner of the
// one-sided Jacobi method: Z^T Z is far closer to diagonal than C, which saves about a fifth of the sweeps, and
// the method keeps its relative accuracy (Cholesky of a symmetrically scaled matrix is as accurate as the
// scaled matrix is well conditioned).
//   Per step: every warp finds the largest remaining diagonal entry from its own copy of the running diagonal
//   (20-bit key, redux + ballot: the pivot only has to be large, and identical in all warps); the lane that owns
//   the pivot column publishes its part of the column, as it is, to shared memory; after ONE block barrier (the
//   buffer alternates) every lane subtracts c[r] * c[col] / pivot from its two columns.  A column that has been
//   a pivot is frozen (multiplier 0).  Scaling by 1 / sqrt(pivot) and the exact zeros of the rows eliminated
//   before a column became the pivot are applied once at the end, from the elimination step of every row and
//   column, instead of in every step.
// Returns false when a pivot is not positive (matrix not positive definite).
template <int RPL, int NW>
__device__ __forceinline__ bool pivoted_cholesky_regs(double (&A)[RPL], double (&B)[RPL], double dA, double dB,
                                                      int n, double* lbuf, int lane, int h, bool act) {
    constexpr int RPLP = (RPL + 2) & ~1;              // rows of one warp in the buffer, padded to 16 bytes
    const double NEG = -INFINITY;
    const int cA = 2 * lane, cB = 2 * lane + 1;
    const int iA = (cA / RPL) * RPLP + cA % RPL, iB = (cB / RPL) * RPLP + cB % RPL;
    int* step_of = reinterpret_cast<int*>(lbuf + 2 * NW * RPLP);   // [NW * RPL] elimination step of a row / column
    double rsA = 0.0, rsB = 0.0;                      // 1 / sqrt(pivot) of my columns
    int kA = 0, kB = 0;                               // their elimination steps
    for (int k = 0; k < n; ++k) {
        const double dm = fmax(dA, dB);
        const unsigned key = dm > 0.0 ? (unsigned)__double2hiint(dm) : 0u;
        const unsigned best = __reduce_max_sync(0xffffffffu, key);
        const int src = __ffs(__ballot_sync(0xffffffffu, key == best)) - 1;
        const int p = __shfl_sync(0xffffffffu, dA >= dB ? cA : cB, src);
        const double dp = __shfl_sync(0xffffffffu, dm, src);
        if (!(dp > 0.0) || !isfinite(dp)) return false;          // uniform over the block
        const double ip = rcp_fast(dp);
        double* lb = lbuf + (k & 1) * (NW * RPLP);
        if (lane == (p >> 1)) {
            double* dst = lb + h * RPLP;
            if (p & 1) {
#pragma unroll
                for (int r = 0; r < RPL; ++r) dst[r] = B[r];
            } else {
#pragma unroll
                for (int r = 0; r < RPL; ++r) dst[r] = A[r];
            }
            if (h == 0) step_of[p] = k;
        }
        __syncthreads();
        double mA = (act && dA != NEG) ? lb[iA] : 0.0, mB = (act && dB != NEG) ? lb[iB] : 0.0;
        if (cA == p) { mA = 0.0; dA = NEG; kA = k; rsA = rsqrt_fast(dp); }
        if (cB == p) { mB = 0.0; dB = NEG; kB = k; rsB = rsqrt_fast(dp); }
        dA = fma(-mA * ip, mA, dA);                   // stays -inf for frozen columns
        dB = fma(-mB * ip, mB, dB);
        mA *= ip;
        mB *= ip;
        const double2* l2 = reinterpret_cast<const double2*>(lb + h * RPLP);
#pragma unroll
        for (int r = 0; r < RPL; r += 2) {
            const double2 l = l2[r >> 1];
            A[r] = fma(-l.x, mA, A[r]);
            B[r] = fma(-l.x, mB, B[r]);
            if (r + 1 < RPL) {
                A[r + 1] = fma(-l.y, mA, A[r + 1]);
                B[r + 1] = fma(-l.y, mB, B[r + 1]);
            }
        }
    }
    __syncthreads();
    // finish: scale, and exact zeros where the row left before the column did (rows that never left: padding)
    const int row0 = h * RPL;
#pragma unroll
    for (int r = 0; r < RPL; ++r) {
        const int tr = (row0 + r < n) ? step_of[row0 + r] : -1;
        A[r] = (tr < kA) ? 0.0 : A[r] * rsA;
        B[r] = (tr < kB) ? 0.0 : B[r] * rsB;
    }
    __syncthreads();                                  // the buffer is scratch of the sweeps from here on
    return true;
}

// Sweep engine for 64 < n <= 116: the odd-even ordering of the register variant with LPP lanes per column pair.
// Pair p holds the columns at line positions 2p (A) and 2p+1 (B) in registers, lane h of the pair rows
// [h RPL, (h+1) RPL).  A step rotates the resident pair, hands the moving column (B on even steps, A on odd
// ones) and its cached norm to the neighbour pair through a double-buffered exchange area in shared memory -- ONE
// block barrier and one column written + one column read per pair and step (the round-robin variant it replaces
// moved both columns and needed two barriers) -- and forms the dot product of the new pairing (shuffle reduction
// over the pair's lanes).  The last pair, which holds the two idle end positions on odd steps, exchanges them
// (c = 0, s = -1).  s_W: slot j = column j on entry and on exit (column order is irrelevant to the caller); in
// between its memory is the exchange area [2][npairs][LD].
template <int RPL, int LPP>
__device__ __forceinline__ int stage_jacobi_oe(const PlanDev& P, double* s_W, double* s_nrm, int max_sweeps, int tid) {
    const int npad = P.npad, LD = P.LD, NP = npad >> 1;
    const int pr = tid / LPP, h = tid % LPP;
    const bool act = pr < NP;
    const bool last = pr == NP - 1;
    const int p = act ? pr : 0;
    const int dn = (p + 1 == NP) ? 0 : p + 1;            // even steps: B comes from the pair below
    const int up = (p == 0) ? NP - 1 : p - 1;            // odd steps: A comes from the pair above
    double A[RPL], B[RPL];
    {
        const double* pa = s_W + (2 * p) * LD + h * RPL;
        const double* pb = pa + LD;
#pragma unroll
        for (int r = 0; r < RPL; ++r) {
            A[r] = act ? pa[r] : 0.0;
            B[r] = act ? pb[r] : 0.0;
        }
    }
    __syncthreads();                                      // the columns are in registers: s_W becomes the exchange area
    double* mine0 = s_W + (size_t)p * LD + h * RPL;       // buffer 0; buffer 1 is NP * LD doubles further
    const double* from_dn0 = s_W + (size_t)dn * LD + h * RPL;
    const double* from_up0 = s_W + (size_t)up * LD + h * RPL;
    const int half = NP * LD;
    auto pair_sum = [&](double v) {
#pragma unroll
        for (int m = 1; m < LPP; m <<= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
        return v;
    };
    int sweeps = 0;
    bool converged = false;
    int cur = 0;
    for (; sweeps < max_sweeps; ++sweeps) {
        double nA = 0.0, nB = 0.0, g = 0.0;
#pragma unroll
        for (int r = 0; r < RPL; ++r) {
            nA = fma(A[r], A[r], nA);
            nB = fma(B[r], B[r], nB);
            g = fma(A[r], B[r], g);
        }
        nA = pair_sum(nA);
        nB = pair_sum(nB);
        g = pair_sum(g);
        int unsettled = 0;
        for (int step = 0; step < npad; ++step) {
            const bool odd = step & 1;
            const bool live = act && !(odd && last);
            const double gg = g * g, pab = nA * nB;
            double c = 1.0, s = 0.0;
            if (live && gg > (BFE2_JACOBI_TAU_WIDE * BFE2_JACOBI_TAU_WIDE) * pab) unsettled = 1;
            if (live && gg > (BFE2_JACOBI_TOL * BFE2_JACOBI_TOL) * pab) {
                const double d = nB - nA, hh = 2.0 * g;
                const double ir = rsqrt_fast(fma(d, d, hh * hh));
                const double hs = (d < 0.0) ? -hh : hh;
                const double c2 = fma(0.5 * fabs(d), ir, 0.5);
                const double ic = rsqrt_fast(c2);
                c = c2 * ic;
                s = (0.5 * hs * ir) * ic;
                const double tg = (s * ic) * g;
                nA -= tg;
                nB += tg;
                if (fabs(s) > BFE2_JACOBI_TAU_WIDE) unsettled = 1;
            }
            if (odd && last && act) {                     // exchange the two idle end columns
                c = 0.0; s = -1.0;
                const double t = nA; nA = nB; nB = t;
            }
            double* out = mine0 + cur * half;
#pragma unroll
            for (int r = 0; r < RPL; ++r) {
                const double sb = s * B[r], cb = c * B[r];
                const double nb = fma(s, A[r], cb);
                const double na = fma(c, A[r], -sb);
                A[r] = na;
                B[r] = nb;
                if (act) out[r] = odd ? na : nb;
            }
            if (act && h == 0) s_nrm[cur * NP + p] = odd ? nA : nB;
            __syncthreads();
            const double* in = (odd ? from_up0 : from_dn0) + cur * half;
            double acc = 0.0;
            if (odd) {
#pragma unroll
                for (int r = 0; r < RPL; ++r) { A[r] = act ? in[r] : 0.0; acc = fma(A[r], B[r], acc); }
                nA = s_nrm[cur * NP + up];
            } else {
#pragma unroll
                for (int r = 0; r < RPL; ++r) { B[r] = act ? in[r] : 0.0; acc = fma(A[r], B[r], acc); }
                nB = s_nrm[cur * NP + dn];
            }
            g = pair_sum(acc);
            cur ^= 1;
        }
        if (!__syncthreads_or(unsettled)) { ++sweeps; converged = true; break; }
    }
    __syncthreads();                                      // exchange area no longer read: the columns come back
    if (act) {
        double* pa = s_W + (2 * p) * LD + h * RPL;
        double* pb = pa + LD;
#pragma unroll
        for (int r = 0; r < RPL; ++r) { pa[r] = A[r]; pb[r] = B[r]; }
    }
    __syncthreads();
    return converged ? sweeps : -sweeps;
}

// The same preconditioner for the shared-memory Jacobi variant (64 < n <= 116): diagonally pivoted Cholesky of the
// symmetric matrix in s_W (slot j = column j), in place, rows in their original order.  Per step warp 0 picks the
// largest remaining diagonal entry (s_d = running diagonal), then thread (column j, row chunk h) subtracts
// c[r] c[j] / pivot from its part of column j -- finished columns are skipped altogether; scaling by
// 1 / sqrt(pivot) and the exact zeros of rows eliminated before a column are applied in one pass at the end
// (s_step = elimination step of every index).  Returns false when a pivot is not positive.
__device__ __forceinline__ bool pivoted_cholesky_smem(const PlanDev& P, double* s_W, double* s_d, int* s_step, int tid,
                                                      int nt) {
    const int n = P.n, LD = P.LD;
    __shared__ int sh_p;
    __shared__ double sh_dp;
    const double NEG = -INFINITY;
    for (int j = tid; j < P.npad; j += nt) {
        s_d[j] = (j < n) ? s_W[j * LD + j] : NEG;
        s_step[j] = -1;
    }
    const int tpc = max(1, nt / n);                  // threads per column
    const int rc = (n + tpc - 1) / tpc;              // rows per thread
    const int j = tid / tpc, h = tid - j * tpc;
    const bool active = j < n;
    const int r0 = h * rc, r1 = min(n, r0 + rc);
    double* mine = s_W + (active ? j : 0) * LD;
    __syncthreads();
    for (int k = 0; k < n; ++k) {
        if (tid < 32) {
            double best = NEG;
            int arg = 0;
            for (int q = tid; q < n; q += 32) {
                const double v = s_d[q];
                if (v > best) { best = v; arg = q; }
            }
            const unsigned key = best > 0.0 ? (unsigned)__double2hiint(best) : 0u;
            const unsigned top = __reduce_max_sync(0xffffffffu, key);
            const int src = __ffs(__ballot_sync(0xffffffffu, key == top)) - 1;
            const int p = __shfl_sync(0xffffffffu, arg, src);
            const double dp = __shfl_sync(0xffffffffu, best, src);
            if (tid == 0) { sh_p = p; sh_dp = dp; }
        }
        __syncthreads();
        const int p = sh_p;
        const double dp = sh_dp;
        if (!(dp > 0.0) || !isfinite(dp)) return false;          // uniform over the block
        const double* cp = s_W + p * LD;
        if (tid == 0) { s_d[p] = NEG; s_step[p] = k; }           // visible to the next pick after the barrier below
        if (active && j != p && s_d[j] != NEG) {
            const double mj = cp[j] * rcp_fast(dp);
            for (int r = r0; r < r1; ++r) mine[r] = fma(-cp[r], mj, mine[r]);
            if (h == 0) s_d[j] = fma(-cp[j], mj, s_d[j]);
        }
        __syncthreads();
    }
    __syncthreads();
    const int kj = active ? s_step[j] : 0;
    const double rs = active ? rsqrt_fast(mine[j]) : 0.0;
    __syncthreads();
    if (active)
        for (int r = r0; r < r1; ++r) mine[r] = (s_step[r] < kj) ? 0.0 : mine[r] * rs;
    __syncthreads();
    return true;
}

// ------------------------------------------------------------------------------------------------
// Register-resident Jacobi for 64 < n <= 128 (round 2): the same odd-even ordering with TWO column groups.
//   Warp (cg, h): column group cg (pairs 32 cg .. 32 cg + 31, one pair per lane), row chunk h (rows h RPL ..).
//   Inside a group the moving column travels by warp shuffle exactly as in the one-group variant; the two pairs per
//   step whose neighbour lives in the OTHER group (pair 31 <-> 32 and the cyclic wrap NP-1 <-> 0) hand their column
//   over through a slot in shared memory: the exporting lane stores it while it rotates, after ONE extra block barrier
//   the importing lane loads it and redoes its part of the next dot product.  Two barriers per step instead of one,
//   but no column ever makes the round trip through shared memory that the 4-lanes-per-pair variant
//   (stage_jacobi_oe) pays on every pair and step.
// ------------------------------------------------------------------------------------------------
template <int RPL, bool ODD>
__device__ __forceinline__ void jacobi_regs2_step(double (&A)[RPL], double (&B)[RPL], double& nA, double& nB,
                                                  int& unsettled, double g, unsigned pg_mine, bool act, bool last, int src,
                                                  bool exp, unsigned exp_addr) {
    const double gg = g * g, pab = nA * nB;
    const bool live = act && !(ODD && last);
    const bool rot = live && (gg > (BFE2_JACOBI_TOL * BFE2_JACOBI_TOL) * pab);
    if (live && gg > (BFE2_JACOBI_TAU_WIDE * BFE2_JACOBI_TAU_WIDE) * pab) unsettled = 1;
    double c = 1.0, s = 0.0;
    if (ODD && last && act) {                         // exchange the two idle end columns
        c = 0.0; s = -1.0;
        const double tmp = nA; nA = nB; nB = tmp;
    }
    if (rot) {
        const double d = nB - nA, hh = 2.0 * g;
        const double ir = rsqrt_fast(fma(d, d, hh * hh));
        const double hs = (d < 0.0) ? -hh : hh;
        const double c2 = fma(0.5 * fabs(d), ir, 0.5);
        const double ic = rsqrt_fast(c2);
        c = c2 * ic;
        s = (0.5 * hs * ir) * ic;
        const double tg = (s * ic) * g;
        nA -= tg;
        nB += tg;
        if (fabs(s) > BFE2_JACOBI_TAU_WIDE) unsettled = 1;
    }
    if (exp) sts_f64(exp_addr + 8u * RPL, ODD ? nA : nB);   // norm of the column that leaves
    double acc0 = 0.0, acc1 = 0.0;
    if (__any_sync(0xffffffffu, rot || (ODD && last && act))) {
#pragma unroll
        for (int r = 0; r < RPL; ++r) {
            const double sb = s * B[r], cb = c * B[r];
            const double nb = fma(s, A[r], cb);
            const double na = fma(c, A[r], -sb);
            if (exp) sts_f64(exp_addr + 8u * r, ODD ? na : nb);
            if (!ODD) { A[r] = na; B[r] = __shfl_sync(0xffffffffu, nb, src); }
            else { B[r] = nb; A[r] = __shfl_sync(0xffffffffu, na, src); }
            if (r & 1) acc1 = fma(A[r], B[r], acc1); else acc0 = fma(A[r], B[r], acc0);
        }
    } else {                                          // nobody in this warp rotates: move and multiply only
#pragma unroll
        for (int r = 0; r < RPL; ++r) {
            if (exp) sts_f64(exp_addr + 8u * r, ODD ? A[r] : B[r]);
            if (!ODD) B[r] = __shfl_sync(0xffffffffu, B[r], src);
            else A[r] = __shfl_sync(0xffffffffu, A[r], src);
            if (r & 1) acc1 = fma(A[r], B[r], acc1); else acc0 = fma(A[r], B[r], acc0);
        }
    }
    if (!ODD) nB = __shfl_sync(0xffffffffu, nB, src);
    else nA = __shfl_sync(0xffffffffu, nA, src);
    sts_f64(pg_mine, acc0 + acc1);
}

// named barrier of the two warps (group 0 and group 1) that own row chunk h: 64 threads, barrier resource 1 + h
__device__ __forceinline__ void pair_barrier(int h) {
    asm volatile("bar.sync %0, 64;" ::"r"(1 + h) : "memory");
}

// the column (and its norm) that arrives from the other column group: reload it and redo this lane's partial dot product
template <int RPL, bool ODD>
__device__ __forceinline__ void jacobi_regs2_import(double (&A)[RPL], double (&B)[RPL], double& nA, double& nB,
                                                    unsigned in_addr, unsigned pg_mine) {
    double acc0 = 0.0, acc1 = 0.0;
#pragma unroll
    for (int r = 0; r < RPL; ++r) {
        const double v = lds_f64(in_addr + 8u * r);
        if (!ODD) B[r] = v; else A[r] = v;
        if (r & 1) acc1 = fma(A[r], B[r], acc1); else acc0 = fma(A[r], B[r], acc0);
    }
    if (!ODD) nB = lds_f64(in_addr + 8u * RPL); else nA = lds_f64(in_addr + 8u * RPL);
    sts_f64(pg_mine, acc0 + acc1);
}

// pivoted Cholesky in the two-group register layout (see pivoted_cholesky_regs): each group proposes its largest
// remaining diagonal entry, one extra barrier per step makes the choice global
template <int RPL, int NW>
__device__ __forceinline__ bool pivoted_cholesky_regs2(double (&A)[RPL], double (&B)[RPL], double dA, double dB,
                                                       int n, double* lbuf, int lane, int h, int cg, bool act) {
    constexpr int RPLP = (RPL + 2) & ~1;
    const double NEG = -INFINITY;
    const int cA = 2 * (32 * cg + lane), cB = cA + 1;
    const int iA = (cA / RPL) * RPLP + cA % RPL, iB = (cB / RPL) * RPLP + cB % RPL;
    int* step_of = reinterpret_cast<int*>(lbuf + 2 * NW * RPLP);          // [NW * RPL] ints
    double* cand = lbuf + 2 * NW * RPLP + (NW * RPL + 2) / 2;             // [2 parity][2 groups][dp, column]
    double rsA = 0.0, rsB = 0.0;
    int kA = 0, kB = 0;
    for (int k = 0; k < n; ++k) {
        const double dm = fmax(dA, dB);
        const unsigned key = dm > 0.0 ? (unsigned)__double2hiint(dm) : 0u;
        const unsigned best = __reduce_max_sync(0xffffffffu, key);
        const int src = __ffs(__ballot_sync(0xffffffffu, key == best)) - 1;
        const int p_loc = __shfl_sync(0xffffffffu, dA >= dB ? cA : cB, src);
        const double dp_loc = __shfl_sync(0xffffffffu, dm, src);
        double* cd = cand + (k & 1) * 4;
        if (h == 0 && lane == 0) { cd[2 * cg] = dp_loc; cd[2 * cg + 1] = (double)p_loc; }
        __syncthreads();
        const double d0 = cd[0], d1 = cd[2];
        const bool second = d1 > d0;                   // ties and NaNs go to group 0: the same choice in every warp
        const int p = (int)(second ? cd[3] : cd[1]);
        const double dp = second ? d1 : d0;
        if (!(dp > 0.0) || !isfinite(dp)) return false;          // uniform over the block
        const double ip = rcp_fast(dp);
        double* lb = lbuf + (k & 1) * (NW * RPLP);
        if (cg == (p >> 6) && lane == ((p >> 1) & 31)) {
            double* dst = lb + h * RPLP;
            if (p & 1) {
#pragma unroll
                for (int r = 0; r < RPL; ++r) dst[r] = B[r];
            } else {
#pragma unroll
                for (int r = 0; r < RPL; ++r) dst[r] = A[r];
            }
            if (h == 0) step_of[p] = k;
        }
        __syncthreads();
        double mA = (act && dA != NEG) ? lb[iA] : 0.0, mB = (act && dB != NEG) ? lb[iB] : 0.0;
        if (cA == p) { mA = 0.0; dA = NEG; kA = k; rsA = rsqrt_fast(dp); }
        if (cB == p) { mB = 0.0; dB = NEG; kB = k; rsB = rsqrt_fast(dp); }
        dA = fma(-mA * ip, mA, dA);
        dB = fma(-mB * ip, mB, dB);
        mA *= ip;
        mB *= ip;
        const double2* l2 = reinterpret_cast<const double2*>(lb + h * RPLP);
#pragma unroll
        for (int r = 0; r < RPL; r += 2) {
            const double2 l = l2[r >> 1];
            A[r] = fma(-l.x, mA, A[r]);
            B[r] = fma(-l.x, mB, B[r]);
            if (r + 1 < RPL) {
                A[r + 1] = fma(-l.y, mA, A[r + 1]);
                B[r + 1] = fma(-l.y, mB, B[r + 1]);
            }
        }
    }
    __syncthreads();
    const int row0 = h * RPL;
#pragma unroll
    for (int r = 0; r < RPL; ++r) {
        const int tr = (row0 + r < n) ? step_of[row0 + r] : -1;
        A[r] = (tr < kA) ? 0.0 : A[r] * rsA;
        B[r] = (tr < kB) ? 0.0 : B[r] * rsB;
    }
    __syncthreads();
    return true;
}

template <int RPL, int NW, bool PRECOND = false>
__device__ __forceinline__ int stage_jacobi_regs2(const PlanDev& P, double* s_W, int max_sweeps, int tid) {
    const int lane = tid & 31, wid = tid >> 5;
    const int cg = wid / NW, h = wid - cg * NW;        // column group, row chunk
    const int npad = P.npad, LD = P.LD, NP = npad >> 1;
    const int pr = 32 * cg + lane;                    // pair (line positions 2 pr, 2 pr + 1)
    const bool act = pr < NP;
    const bool last = pr == NP - 1;
    const int lastlane = cg == 0 ? 31 : NP - 33;      // last active lane of this group (NP > 32 here)
    const int src_dn = (act && lane < lastlane) ? lane + 1 : lane;     // the group's last lane imports instead
    const int src_up = (act && lane > 0) ? lane - 1 : lane;            // lane 0 imports instead
    double A[RPL], B[RPL];
    {
        const double* pa = s_W + (2 * pr) * LD + h * RPL;
        const double* pb = pa + LD;
#pragma unroll
        for (int r = 0; r < RPL; ++r) {
            A[r] = act ? pa[r] : 0.0;
            B[r] = act ? pb[r] : 0.0;
        }
    }
    double dA = -INFINITY, dB = -INFINITY;
    if (PRECOND && act) {
        if (2 * pr < P.n) dA = s_W[(2 * pr) * LD + 2 * pr];
        if (2 * pr + 1 < P.n) dB = s_W[(2 * pr + 1) * LD + 2 * pr + 1];
    }
    __syncthreads();                                  // W is in registers: its memory becomes scratch
    constexpr int W32 = 2 * NW * 32;                  // lanes of the block
    constexpr int SLOT = RPL + 2;                     // exported column + its norm
    constexpr int OFF_EXP = 5 * W32, OFF_LBUF = (OFF_EXP + 2 * NW * SLOT + 1) & ~1;
    if (PRECOND) {
        if (!pivoted_cholesky_regs2<RPL, NW>(A, B, dA, dB, P.n, s_W + OFF_LBUF, lane, h, cg, act)) return INT_MIN;
    }
    const unsigned base = (unsigned)__cvta_generic_to_shared(s_W);
    const unsigned pg0 = base, pg1 = base + W32 * 8u;                  // [2 parity][group][row chunk][lane]
    const unsigned mine = 8u * ((cg * NW + h) * 32 + lane), w0 = 8u * (cg * NW * 32 + lane);
    double* part_n = s_W + 2 * W32;                                    // [W32][3]
    const unsigned slot_mine = base + 8u * (OFF_EXP + (cg * NW + h) * SLOT);
    const unsigned slot_other = base + 8u * (OFF_EXP + ((1 - cg) * NW + h) * SLOT);
    // even steps: B comes from pair + 1 -- lane 0 of each group exports, the group's last active lane imports;
    // odd steps:  A comes from pair - 1 -- the group's last active lane exports, lane 0 imports
    const bool exp_even = act && lane == 0, imp_even = act && lane == lastlane;
    const bool exp_odd = act && lane == lastlane, imp_odd = act && lane == 0;
    int sweeps = 0;
    bool converged = false;
    for (; sweeps < max_sweeps; ++sweeps) {
        double nA = 0.0, nB = 0.0, g = 0.0;
        {
            double a0 = 0.0, b0 = 0.0, c0 = 0.0;
#pragma unroll
            for (int r = 0; r < RPL; ++r) {
                a0 = fma(A[r], A[r], a0);
                b0 = fma(B[r], B[r], b0);
                c0 = fma(A[r], B[r], c0);
            }
            const int me = ((cg * NW + h) * 32 + lane) * 3;
            part_n[me] = a0;
            part_n[me + 1] = b0;
            part_n[me + 2] = c0;
            __syncthreads();
#pragma unroll
            for (int w = 0; w < NW; ++w) {            // same order in every warp of the group: identical bits
                const int o = ((cg * NW + w) * 32 + lane) * 3;
                nA += part_n[o];
                nB += part_n[o + 1];
                g += part_n[o + 2];
            }
        }
        int unsettled = 0;
#pragma unroll 1
        for (int step = 0; step < npad; step += 2) {
            jacobi_regs2_step<RPL, false>(A, B, nA, nB, unsettled, g, pg1 + mine, act, last, src_dn, exp_even, slot_mine);
            pair_barrier(h);                          // only the two warps of row chunk h exchange a column
            if (imp_even) jacobi_regs2_import<RPL, false>(A, B, nA, nB, slot_other, pg1 + mine);
            __syncthreads();
            g = lds_f64(pg1 + w0);
#pragma unroll
            for (int w = 1; w < NW; ++w) g += lds_f64(pg1 + w0 + 256u * w);
            jacobi_regs2_step<RPL, true>(A, B, nA, nB, unsettled, g, pg0 + mine, act, last, src_up, exp_odd, slot_mine);
            pair_barrier(h);
            if (imp_odd) jacobi_regs2_import<RPL, true>(A, B, nA, nB, slot_other, pg0 + mine);
            __syncthreads();
            g = lds_f64(pg0 + w0);
#pragma unroll
            for (int w = 1; w < NW; ++w) g += lds_f64(pg0 + w0 + 256u * w);
        }
        if (!__syncthreads_or(unsettled)) { ++sweeps; converged = true; break; }
    }
    __syncthreads();                                  // scratch no longer read: W comes back
    if (act) {
        double* pa = s_W + (2 * pr) * LD + h * RPL;
        double* pb = pa + LD;
#pragma unroll
        for (int r = 0; r < RPL; ++r) { pa[r] = A[r]; pb[r] = B[r]; }
    }
    __syncthreads();
    return converged ? sweeps : -sweeps;
}

// PRECOND: s_W holds the symmetric matrix C itself (slot j = column j); it is first replaced, in registers, by its
// pivoted Cholesky factor Z (C = Z Z^T), and the sweeps orthogonalise the columns of Z:  Z V = U Sigma, so that
// at convergence column k is  sqrt(lam_k) u_k  with  C u_k = lam_k u_k.  Returns INT_MIN if C is not positive
// definite.
template <int RPL, int NW, bool PRECOND = false>
__device__ __forceinline__ int stage_jacobi_regs(const PlanDev& P, double* s_W, int max_sweeps, int tid) {
    const int lane = tid & 31, h = tid >> 5;          // h = row chunk owned by this warp, 0 .. NW-1
    const int npad = P.npad, LD = P.LD, NP = npad >> 1;
    const bool act = lane < NP;
    const bool last = lane == NP - 1;
    const int src_dn = act ? (lane + 1 == NP ? 0 : lane + 1) : lane;
    const int src_up = act ? (lane == 0 ? NP - 1 : lane - 1) : lane;
    double A[RPL], B[RPL];
    {
        const double* pa = s_W + (2 * lane) * LD + h * RPL;
        const double* pb = pa + LD;
#pragma unroll
        for (int r = 0; r < RPL; ++r) {
            A[r] = act ? pa[r] : 0.0;
            B[r] = act ? pb[r] : 0.0;
        }
    }
    double dA = -INFINITY, dB = -INFINITY;            // running diagonal of the two columns (PRECOND)
    if (PRECOND && act) {
        if (2 * lane < P.n) dA = s_W[(2 * lane) * LD + 2 * lane];
        if (2 * lane + 1 < P.n) dB = s_W[(2 * lane + 1) * LD + 2 * lane + 1];
    }
    __syncthreads();                                  // W is in registers: its memory becomes scratch
    if (PRECOND) {
        if (!pivoted_cholesky_regs<RPL, NW>(A, B, dA, dB, P.n, s_W + 5 * NW * 32, lane, h, act)) return INT_MIN;
    }
    const unsigned pg0 = (unsigned)__cvta_generic_to_shared(s_W);      // [2 parity][NW warps][32 lanes]
    const unsigned pg1 = pg0 + NW * 32u * 8u;
    const unsigned mine = 8u * (h * 32 + lane), w0 = 8u * lane;
    double* part_n = s_W + 2 * NW * 32;               // [NW warps][32 lanes][3]
    int sweeps = 0;
    bool converged = false;
    for (; sweeps < max_sweeps; ++sweeps) {
        // refresh the cached squared norms from the data and form the first dot product of the sweep
        double nA = 0.0, nB = 0.0, g = 0.0;
        {
            double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0, c0 = 0.0, c1 = 0.0;
#pragma unroll
            for (int r = 0; r + 1 < RPL; r += 2) {
                a0 = fma(A[r], A[r], a0); a1 = fma(A[r + 1], A[r + 1], a1);
                b0 = fma(B[r], B[r], b0); b1 = fma(B[r + 1], B[r + 1], b1);
                c0 = fma(A[r], B[r], c0); c1 = fma(A[r + 1], B[r + 1], c1);
            }
            if (RPL & 1) {
                a0 = fma(A[RPL - 1], A[RPL - 1], a0); b0 = fma(B[RPL - 1], B[RPL - 1], b0);
                c0 = fma(A[RPL - 1], B[RPL - 1], c0);
            }
            part_n[(h * 32 + lane) * 3] = a0 + a1;
            part_n[(h * 32 + lane) * 3 + 1] = b0 + b1;
            part_n[(h * 32 + lane) * 3 + 2] = c0 + c1;
            __syncthreads();
#pragma unroll
            for (int w = 0; w < NW; ++w) {            // same order in every warp: identical bits
                nA += part_n[(w * 32 + lane) * 3];
                nB += part_n[(w * 32 + lane) * 3 + 1];
                g += part_n[(w * 32 + lane) * 3 + 2];
            }
        }
        int unsettled = 0;
#pragma unroll 1
        for (int step = 0; step < npad; step += 2) {
            jacobi_regs_step<RPL, false>(A, B, nA, nB, unsettled, g, pg1 + mine, act, last, src_dn);
            __syncthreads();
            g = lds_f64(pg1 + w0);
#pragma unroll
            for (int w = 1; w < NW; ++w) g += lds_f64(pg1 + w0 + 256u * w);
            jacobi_regs_step<RPL, true>(A, B, nA, nB, unsettled, g, pg0 + mine, act, last, src_up);
            __syncthreads();
            g = lds_f64(pg0 + w0);
#pragma unroll
            for (int w = 1; w < NW; ++w) g += lds_f64(pg0 + w0 + 256u * w);
        }
        if (!__syncthreads_or(unsettled)) { ++sweeps; converged = true; break; }
    }
    __syncthreads();                                  // scratch no longer read: W comes back
    if (act) {
        double* pa = s_W + (2 * lane) * LD + h * RPL;
        double* pb = pa + LD;
#pragma unroll
        for (int r = 0; r < RPL; ++r) { pa[r] = A[r]; pb[r] = B[r]; }
    }
    __syncthreads();
    return converged ? sweeps : -sweeps;
}

// Form C = Lm^-1 K Lm^-T (the standard form of K phi = lam M phi, scipy's sygst) in the slot layout: slot j holds
// column j of C.  Two passes of band forward substitutions with the factor Lm of M: thread j turns column j of K
// (read from its unfactored band) into column j of X = Lm^-1 K; then thread i turns row i of X into row i of
// C = X Lm^-T (a substitution along the slots, stride LD).
// Column `tid` of the band matrix K (rows tid - b .. tid + b) into registers, before K is factored in place.
constexpr int KCOL_HB = 8;
__device__ __forceinline__ void load_k_column(const PlanDev& P, const double* s_Kb, double (&kcol)[2 * KCOL_HB + 1],
                                              int tid) {
    const int n = P.n, b = P.hbw;
#pragma unroll
    for (int d = -KCOL_HB; d <= KCOL_HB; ++d) {
        const int i = tid + d;
        const bool in = tid < n && d >= -b && d <= b && i >= 0 && i < n;
        kcol[d + KCOL_HB] = in ? (d < 0 ? s_Kb[(-d) * n + i] : s_Kb[d * n + tid]) : 0.0;
    }
}

// kcol != nullptr: the columns of K come from the registers filled by load_k_column (needs n <= blockDim and
// hbw <= KCOL_HB), s_Kb may already hold the factor of K.
__device__ __forceinline__ void stage_form_a(const PlanDev& P, const double* s_Kb, const double* s_Mb,
                                             const double* s_invM, double* s_W, int tid, int nt,
                                             const double (*kcol)[2 * KCOL_HB + 1] = nullptr, bool clean = false) {
    const int n = P.n, b = P.hbw, LD = P.LD;
    for (int t = tid; t < P.npad * LD; t += nt) s_W[t] = 0.0;
    __syncthreads();
    for (int j = tid; j < n; j += nt) {
        double* col = s_W + j * LD;
        const int lo = max(0, j - b), hi = min(n - 1, j + b);
        if (kcol != nullptr) {
#pragma unroll
            for (int d = -KCOL_HB; d <= KCOL_HB; ++d)
                if (j + d >= lo && j + d <= hi) col[j + d] = (*kcol)[d + KCOL_HB];
        } else {
            for (int i = lo; i < j; ++i) col[i] = s_Kb[(j - i) * n + i];
            for (int i = j; i <= hi; ++i) col[i] = s_Kb[(i - j) * n + j];
        }
        if (b <= 8) band_subst<true, false>(s_Mb, s_invM, col, 1, n, b, lo, clean);
        else {
            for (int r = lo; r < n; ++r) {
                double acc = col[r];
                for (int k = max(lo, r - b); k < r; ++k) acc -= s_Mb[(r - k) * n + k] * col[k];
                col[r] = acc * s_invM[r];
            }
        }
    }
    __syncthreads();
    for (int i = tid; i < n; i += nt) {
        double* row = s_W + i;                        // element j of the row lives at row[j * LD]
        if (b <= 8) band_subst<true, false>(s_Mb, s_invM, row, LD, n, b, 0, clean);
        else {
            for (int r = 0; r < n; ++r) {
                double acc = row[r * LD];
                for (int k = max(0, r - b); k < r; ++k) acc -= s_Mb[(r - k) * n + k] * row[k * LD];
                row[r * LD] = acc * s_invM[r];
            }
        }
    }
    __syncthreads();
}

// One pass over a column for the refinement below: phi = Lm^-T (scale * w) by back substitution with the active window
// in registers and, in the same sweep over the rows, y_j = (Lk^T phi)_j = sum_d Lk[j+d][j] phi[j+d] from the same
// window.  Returns |Lk^T phi|^2; `flip` tells whether the largest-magnitude entry of phi is negative.  All loads of a
// row are independent of the recurrence, and the row loop is unrolled so that they are issued in batches.
template <int HB>
__device__ __forceinline__ double backsolve_rq(const double* __restrict__ Mf, const double* __restrict__ invM,
                                               const double* __restrict__ Lk, double* x, int n, int b, double scale,
                                               bool& flip) {
    double win[HB];                                  // win[d-1] = phi[j+d]
#pragma unroll
    for (int d = 0; d < HB; ++d) win[d] = 0.0;
    double l0 = 0.0, l1 = 0.0, best = 0.0;
    flip = false;
#pragma unroll 4
    for (int j = n - 1; j >= 0; --j) {
        double far = x[j] * scale;
        double y = 0.0;
#pragma unroll
        for (int d = HB; d >= 1; --d) {
            const bool in = d <= b && j + d < n;
            const double lm = in ? Mf[d * n + j] : 0.0;
            const double lk = in ? Lk[d * n + j] : 0.0;
            far = fma(-lm, win[d - 1], far);
            y = fma(lk, win[d - 1], y);
        }
        const double v = far * invM[j];
        y = fma(Lk[j], v, y);
        if (j & 1) l1 = fma(y, y, l1); else l0 = fma(y, y, l0);
#pragma unroll
        for (int d = HB - 1; d >= 1; --d) win[d] = win[d - 1];
        win[0] = v;
        x[j] = v;
        const double av = fabs(v);
        if (av >= best) { best = av; flip = v < 0.0; }   // rows descend: on a tie the LOWER index decides, as before
    }
    return l0 + l1;
}

// Rayleigh-quotient refinement of every column (see stage_modal_output).  Kept OUT OF LINE on purpose: inlined, its
// register needs disturb the allocation of the register-resident sweep loop of k_solve.
static __device__ __noinline__ void refine_columns(int n, int npad, int LD, int b, const double* s_Mf, const double* s_invM,
                                                   const double* s_Lk, double* s_W, double* s_lam, int tid, int nt, bool clean) {
    (void)clean;
    for (int s = tid; s < npad; s += nt) {
        double* w = s_W + s * LD;
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        int r = 0;
#pragma unroll 2
        for (; r + 3 < n; r += 4) {
            a0 = fma(w[r], w[r], a0); a1 = fma(w[r + 1], w[r + 1], a1);
            a2 = fma(w[r + 2], w[r + 2], a2); a3 = fma(w[r + 3], w[r + 3], a3);
        }
        for (; r < n; ++r) a0 = fma(w[r], w[r], a0);
        const double acc = (a0 + a1) + (a2 + a3);
        if (acc == 0.0 || !(acc < INFINITY)) {               // the zero dummy column (n odd) sorts last
            s_lam[s] = (npad != n && acc == 0.0) ? INFINITY : acc;
            continue;
        }
        const double sc = rsqrt_fast(acc);                   // u = sc w, |u| = 1
        bool flip = false;
        double lam;
        if (b <= 5) lam = backsolve_rq<5>(s_Mf, s_invM, s_Lk, w, n, b, sc, flip);
        else if (b <= 8) lam = backsolve_rq<8>(s_Mf, s_invM, s_Lk, w, n, b, sc, flip);
        else {                                               // wide bands / dense pencils: plain loops
            for (int i = 0; i < n; ++i) w[i] *= sc;
            band_backsolve_thread(s_Mf, s_invM, w, n, b);
            lam = 0.0;
            double best = 0.0;
            for (int i = 0; i < n; ++i) {
                const int m = min(b, n - 1 - i);
                double y = 0.0;
                for (int d = 0; d <= m; ++d) y = fma(s_Lk[d * n + i], w[i + d], y);
                lam = fma(y, y, lam);
                const double v = fabs(w[i]);
                if (v > best) { best = v; flip = w[i] < 0.0; }
            }
        }
        s_lam[s] = lam;
        if (flip)
            for (int i = 0; i < n; ++i) w[i] = -w[i];        // sign convention: largest-magnitude entry positive
    }
}

// After convergence: eigenvalues (ascending) and the first n_modes shapes to global memory.
// The columns are sqrt(lam_k) u_k (C u = lam u): shapes phi_k = Lm^-T u_k, back-substituted with the band factor of M
// passed as s_Kb / s_invK.
// the plan fields the output stage needs, by value: the stage is kept OUT OF LINE (see below) and must not take the
// address of the kernel's parameter block
struct OutDims {
    int n, npad, LD, hbw, sumdof;
    const int32_t* free_of_pos;
};

static __device__ __noinline__ void modal_output_impl(OutDims P, const double* s_Kb, const double* s_invK,
                                                      double* s_W, double* s_lam, int* s_order, int n_modes,
                                                      double* g_lam, double* g_phi, int tid, int nt,
                                                      bool clean, bool compact, const double* s_Lk) {
    // s_Kb / s_invK: band factor of M (historic names).  s_Lk != nullptr: band factor of K -- every eigenvalue is then
    // REFINED as the Rayleigh quotient of its shape taken on the two Cholesky factors,
    //     lam_k = |Lk^T phi_k|^2,  phi_k = Lm^-T u_k,  |u_k| = 1  (phi^T M phi = |Lm^T phi|^2 = 1),
    // which is second-order accurate in the error of u_k and never sees the rounding of the formed matrix
    // C = Lm^-1 K Lm^-T or of its pivoted Cholesky factor: the smallest eigenvalues of badly graded pencils gain about
    // a digit (worst case of the 2.4 M-variant soak: 7e-11 -> see profiles/r2_soak.txt).  One thread per column; the
    // shapes of ALL columns are formed (the first n_modes of them are the output), which costs latency only.
    const int n = P.n, npad = P.npad, LD = P.LD, b = P.hbw;
    if (s_Lk != nullptr) {
        refine_columns(n, npad, LD, b, s_Kb, s_invK, s_Lk, s_W, s_lam, tid, nt, clean);
    } else {
        // squared column norms = eigenvalues; the dummy slot (n odd) sorts last
        for (int s = tid; s < npad; s += nt) {
            double acc = 0.0;
            const double* w = s_W + s * LD;
            for (int r = 0; r < n; ++r) acc = fma(w[r], w[r], acc);
            s_lam[s] = (npad != n && acc == 0.0) ? INFINITY : acc;   // the zero dummy column (n odd) sorts last
        }
    }
    __syncthreads();
    for (int s = tid; s < npad; s += nt) {
        const double v = s_lam[s];
        int rank = 0;
        for (int q = 0; q < npad; ++q) {
            const double w = s_lam[q];
            rank += (w < v || (w == v && q < s)) ? 1 : 0;
        }
        s_order[rank] = s;
        if (rank < n) g_lam[rank] = v;
    }
    __syncthreads();
    if (n_modes <= 0 || g_phi == nullptr) return;
    if (s_Lk == nullptr) {
        // phi = Lm^-T w, in place in the slot; then fix the sign (largest-magnitude entry positive)
        for (int m = tid; m < n_modes; m += nt) {
            double* w = s_W + s_order[m] * LD;
            if (b <= 8) band_subst<false, true>(s_Kb, s_invK, w, 1, n, b, 0, clean);
            else band_backsolve_thread(s_Kb, s_invK, w, n, b);
            double best = 0.0, sign = 1.0;
            for (int r = 0; r < n; ++r) {
                const double v = fabs(w[r]);
                if (v > best) { best = v; sign = (w[r] < 0.0) ? -1.0 : 1.0; }
            }
            sign *= rsqrt_fast(s_lam[s_order[m]]);
            for (int r = 0; r < n; ++r) w[r] *= sign;
        }
        __syncthreads();
    }
    if (compact) {                                       // shapes on the free DOFs only: [n_modes, n]
        for (int t = tid; t < n_modes * n; t += nt) g_phi[t] = s_W[s_order[t / n] * LD + t % n];
        return;
    }
    for (int t = tid; t < n_modes * P.sumdof; t += nt) {
        const int m = t / P.sumdof, pos = t % P.sumdof;
        const int j = P.free_of_pos[pos];
        g_phi[t] = (j >= 0) ? s_W[s_order[m] * LD + j] : 0.0;
    }
}

// Out of line on purpose: inlined into k_solve, the output stage changes what ptxas keeps live around the
// register-resident sweep loop and costs the P20 launch several per cent.
template <int LPP>
__device__ __forceinline__ void stage_modal_output(const PlanDev& P, const double* s_Kb, const double* s_invK,
                                                   double* s_W, double* s_lam, int* s_order, int n_modes,
                                                   double* g_lam, double* g_phi, int tid, int nt,
                                                   bool clean = false, bool compact = false,
                                                   const double* s_Lk = nullptr) {
    const OutDims d{P.n, P.npad, P.LD, P.hbw, P.sumdof, P.free_of_pos};
    modal_output_impl(d, s_Kb, s_invK, s_W, s_lam, s_order, n_modes, g_lam, g_phi, tid, nt, clean, compact, s_Lk);
}

}  // namespace bfe2
