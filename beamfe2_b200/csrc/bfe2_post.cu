// bfe2_post.cu -- post-processing kernels next to the hot path.
//
//   bfe2_internal_actions : N / V / M along every beam from the local end forces (HermitianBeam_2D.py:398-431:
//       axial, shear: -v1; moment: v1 + (-v2 - v1) pos) plus the add-on curves of UNIFORM beam loads
//       (Loads.py:214-225 axial: -xi q L;  :277-289 shear: -xi L q, moment: xi (1-xi) q L^2 / 2) and of the three
//       CONCENTRATED ones (Loads.py:347-365, 421-432, 489-502), and the per-beam maxima the reference lists as
//       future work (todo.md:23) -- with the one-sided limits at the load positions, where the curves jump.
//   bfe2_modal_participation : Gamma_k = phi_k^T M r for the rigid-body influence vectors r_x, r_y
//       (prototype arithmetic in Examples/modal_masses.py:40-49; phi is M-normalised, so the effective modal
//       mass is Gamma^2).
#include <cuda_runtime.h>

#include "../../include/bfe2.h"
#include "bfe2_device.cuh"

using namespace bfe2;

int bfe2_fail(int code, const char* fmt, ...);
int bfe2_plan_device_view(bfe2_plan* plan, PlanDev* out);   // bfe2.cu: view on the current device, < 0 on error

namespace {

// add-on curves of the concentrated beam loads at normalised position xi (Loads.py:347-365 perpendicular force,
// :421-432 axial force, :489-502 moment), with the reference's own comparisons at xi == position.  `side` != 0
// evaluates the one-sided limit at the position of load `kink` instead (left: -1, right: +1) -- used for the maxima,
// because shear / axial / moment jump there.
__device__ __forceinline__ void conc_addon(const double* __restrict__ cl, int n_conc, double xi, double L, int kink,
                                           int side, double& ax, double& sh, double& mo) {
    for (int j = 0; j < n_conc; ++j) {
        const int kind = (int)cl[3 * j];
        if (kind == 0) continue;
        const double P = cl[3 * j + 1], a = cl[3 * j + 2];
        bool below = xi < a, above = xi > a;             // the reference's tests
        if (j == kink && side != 0) { below = side < 0; above = side > 0; }
        if (kind == BFE2_LOAD_CONC_PERP) {
            double m = (1.0 - a) * P * xi;
            if (above) m -= fabs(xi - a) * P;
            mo += m * L;
            sh += below ? 0.0 : -P;
        } else if (kind == BFE2_LOAD_CONC_AXIAL) {
            ax += below ? 0.0 : -P;
        } else if (kind == BFE2_LOAD_CONC_MOMENT) {
            double m = -(P / L) * xi * L;
            if (above) m += P;
            mo += m;
        }
    }
}

__global__ void k_internal_actions(PlanDev P, long long B, int C, int npts, const double* __restrict__ coords,
                                   const double* __restrict__ endforces, const double* __restrict__ q_axial,
                                   const double* __restrict__ q_perp, int n_conc, const double* __restrict__ conc,
                                   double* __restrict__ actions, double* __restrict__ maxabs) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = B * C * P.nE;
    if (t >= total) return;
    const int e = (int)(t % P.nE);
    const long long b = t / ((long long)C * P.nE);
    const int ni = P.conn[2 * e], nj = P.conn[2 * e + 1];
    const double* xy = coords + b * 2 * P.nN;
    double L, c, s;
    beam_geometry(xy[2 * ni], xy[2 * ni + 1], xy[2 * nj], xy[2 * nj + 1], L, c, s);
    const double* q = endforces + t * 6;
    const double qa = q_axial ? q_axial[t] : 0.0, qp = q_perp ? q_perp[t] : 0.0;
    const double* cl = (conc && n_conc > 0) ? conc + t * 3 * n_conc : nullptr;
    const int nc = cl ? n_conc : 0;
    const double n1 = q[0], v1 = q[1], m1 = q[2], m2 = -q[5];
    double mx[3] = {0.0, 0.0, 0.0};
    auto eval = [&](double xi, int kink, int side, double& ax, double& sh, double& mo) {
        ax = -n1 + (-xi * qa * L);
        sh = -v1 + (xi * L * -qp);
        mo = (m1 + (m2 - m1) * xi) + ((xi * (1.0 - xi)) * qp * (L * L)) / 2.0;
        if (nc) conc_addon(cl, nc, xi, L, kink, side, ax, sh, mo);
    };
    for (int k = 0; k < npts; ++k) {
        const double xi = npts > 1 ? (double)k / (double)(npts - 1) : 0.0;
        double ax, sh, mo;
        eval(xi, -1, 0, ax, sh, mo);
        if (actions) {
            double* o = actions + t * 3 * npts;
            o[k] = ax; o[npts + k] = sh; o[2 * npts + k] = mo;
        }
        mx[0] = fmax(mx[0], fabs(ax)); mx[1] = fmax(mx[1], fabs(sh)); mx[2] = fmax(mx[2], fabs(mo));
    }
    if (maxabs) {
        // the extrema of piecewise-linear shear / axial curves and the moment kinks sit at the load positions: both
        // one-sided limits there join the sampled values
        for (int j = 0; j < nc; ++j) {
            if ((int)cl[3 * j] == 0) continue;
            const double a = cl[3 * j + 2];
            for (int side = -1; side <= 1; side += 2) {
                double ax, sh, mo;
                eval(a, j, side, ax, sh, mo);
                mx[0] = fmax(mx[0], fabs(ax)); mx[1] = fmax(mx[1], fabs(sh)); mx[2] = fmax(mx[2], fabs(mo));
            }
        }
        maxabs[t * 3] = mx[0]; maxabs[t * 3 + 1] = mx[1]; maxabs[t * 3 + 2] = mx[2];
    }
}

// thread per (structure, mode): Gamma = sum_i phi_i (M r)_i with the lower-band M of the condensed system
__global__ void k_participation(PlanDev P, long long B, int n_modes, const double* __restrict__ Mb,
                                const double* __restrict__ phi, double* __restrict__ gamma) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * n_modes) return;
    const long long b = t / n_modes;
    const int n = P.n, hb = P.hbw;
    const double* M = Mb + b * P.nslots;
    const double* f = phi + t * P.sumdof;
    double gx = 0.0, gy = 0.0;
    for (int j = 0; j < n; ++j) {
        const double fj = f[P.pos_of_free[j]];
        const int dj = P.pos_of_free[j] % 3;
        // column j of M (lower band) couples rows j .. j+hb; use symmetry for the upper part
        for (int d = 0; d <= hb && j + d < n; ++d) {
            const double m = M[d * n + j];
            const int i = j + d;
            const double fi = f[P.pos_of_free[i]];
            const int di = P.pos_of_free[i] % 3;
            // contribution phi_i M_ij r_j  and (d > 0) phi_j M_ji r_i
            if (dj == 0) gx += fi * m; else if (dj == 1) gy += fi * m;
            if (d > 0) { if (di == 0) gx += fj * m; else if (di == 1) gy += fj * m; }
        }
    }
    gamma[t * 2] = gx;
    gamma[t * 2 + 1] = gy;
}

// ---- geometric stiffness: banded assembly of K_g(N) for given axial forces ---------------
// Local matrix of a beam with axial force N (tension positive), DOF order [ux_i, uy_i, rz_i, ux_j, uy_j, rz_j]:
//   N/L * [ 6/5, L/10, -6/5, L/10 ; L/10, 2L^2/15, -L/10, -L^2/30 ; -6/5, -L/10, 6/5, -L/10 ; L/10, -L^2/30, -L/10, 2L^2/15 ]
// on the (uy_i, rz_i, uy_j, rz_j) rows / columns, zero on the axial ones (the consistent matrix of the cubic Hermite
// shape functions; H.-P. Gavin, CEE 421L, which HermitianBeam_2D.py:261-282 cites -- the reference's own transcription
// has sign slips in four entries and its solver is unfinished, README.md:27; this is the matrix it names).
__device__ __forceinline__ double kg_local(int a, int b, double g, double L) {
    if (a == 0 || a == 3 || b == 0 || b == 3) return 0.0;
    const bool ta = (a == 1 || a == 4), tb = (b == 1 || b == 4);       // translation (uy) or rotation
    const bool ja = a >= 3, jb = b >= 3;                               // end j
    if (ta && tb) return (ja == jb ? 1.2 : -1.2) * g;
    if (ta != tb) {                                                    // uy - rz coupling: +L/10 from uy_i, -L/10 from uy_j
        const bool tj = ta ? ja : jb;                                  // the end of the translation
        return (tj ? -0.1 : 0.1) * g * L;
    }
    return (ja == jb ? 2.0 / 15.0 : -1.0 / 30.0) * g * L * L;
}

// thread per (structure, band slot); contributions are added in beam-list order like every other assembly here
__global__ void k_assemble_geometric(PlanDev P, long long B, const double* __restrict__ coords,
                                     const double* __restrict__ axial, double* __restrict__ Kgb) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * P.nslots) return;
    const long long b = t / P.nslots;
    const int slot = (int)(t % P.nslots);
    const double* xy = coords + b * 2 * P.nN;
    double acc = 0.0;
    for (int p = P.slot_ptr[slot]; p < P.slot_ptr[slot + 1]; ++p) {
        const int cbt = P.contrib[p];
        const int e = cbt / 36, r = (cbt % 36) / 6, c = cbt % 6;
        const int ni = P.conn[2 * e], nj = P.conn[2 * e + 1];
        double L, cs, sn;
        beam_geometry(xy[2 * ni], xy[2 * ni + 1], xy[2 * nj], xy[2 * nj + 1], L, cs, sn);
        const double g = axial[b * P.nE + e] / L;
        // (T kl T^T)[r][c] = sum_{a in block(r), d in block(c)} R[r][a] kl[a][d] R[c][d],  R = [[cs,-sn,0],[sn,cs,0],[0,0,1]]
        const int r0 = 3 * (r / 3), c0 = 3 * (c / 3);
        const double Rr[3] = {r % 3 == 0 ? cs : (r % 3 == 1 ? sn : 0.0), r % 3 == 0 ? -sn : (r % 3 == 1 ? cs : 0.0),
                              r % 3 == 2 ? 1.0 : 0.0};
        const double Rc[3] = {c % 3 == 0 ? cs : (c % 3 == 1 ? sn : 0.0), c % 3 == 0 ? -sn : (c % 3 == 1 ? cs : 0.0),
                              c % 3 == 2 ? 1.0 : 0.0};
        double v = 0.0;
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int d = 0; d < 3; ++d) v += Rr[a] * kg_local(r0 + a, c0 + d, g, L) * Rc[d];
        acc += v;
    }
    Kgb[t] = acc;
}

// symmetric band matrix (lower band, [hbw+1][n]: band[d][j] = A[j+d][j]) times a multi-vector X [n, nrhs] (row-major):
// thread per (row, column); used by the first-k modal route of large frames (K X, M X of the subspace iteration)
__global__ void k_band_matvec(int n, int hbw, int nrhs, const double* __restrict__ band, const double* __restrict__ X,
                              double* __restrict__ Y) {
    // blockIdx.y = structure of a batch: its own band [hbw+1, n] and vectors [n, nrhs]
    band += (long long)blockIdx.y * (hbw + 1) * n;
    X += (long long)blockIdx.y * n * nrhs;
    Y += (long long)blockIdx.y * n * nrhs;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)n * nrhs) return;
    const int i = (int)(t / nrhs), r = (int)(t % nrhs);
    double acc = band[i] * X[t];
    for (int d = 1; d <= hbw; ++d) {
        if (i + d < n) acc = fma(band[d * n + i], X[(long long)(i + d) * nrhs + r], acc);          // A[i+d][i] x[i+d]
        if (i - d >= 0) acc = fma(band[d * n + i - d], X[(long long)(i - d) * nrhs + r], acc);      // A[i][i-d] x[i-d]
    }
    Y[t] = acc;
}

}  // namespace

extern "C" {

int bfe2_band_matvec_batched(int32_t n, int32_t hbw, int32_t nrhs, int64_t B, const double* band, const double* X,
                             double* Y, void* stream) {
    if (n < 1 || hbw < 0 || nrhs < 1 || B < 1 || B > 65535 || !band || !X || !Y || X == Y)
        return bfe2_fail(-1, "bfe2_band_matvec: bad argument");
    const long long total = (long long)n * nrhs;
    k_band_matvec<<<dim3((unsigned)((total + 127) / 128), (unsigned)B), 128, 0, (cudaStream_t)stream>>>(n, hbw, nrhs, band, X, Y);
    if (cudaGetLastError() != cudaSuccess) return bfe2_fail(-100, "k_band_matvec launch failed");
    return 0;
}

int bfe2_band_matvec(int32_t n, int32_t hbw, int32_t nrhs, const double* band, const double* X, double* Y, void* stream) {
    if (n < 1 || hbw < 0 || nrhs < 1 || !band || !X || !Y || X == Y) return bfe2_fail(-1, "bfe2_band_matvec: bad argument");
    const long long total = (long long)n * nrhs;
    k_band_matvec<<<(unsigned)((total + 127) / 128), 128, 0, (cudaStream_t)stream>>>(n, hbw, nrhs, band, X, Y);
    if (cudaGetLastError() != cudaSuccess) return bfe2_fail(-100, "k_band_matvec launch failed");
    return 0;
}


int bfe2_assemble_geometric(bfe2_plan* plan, int64_t B, const double* coords, const double* axial, double* Kgb,
                            void* stream) {
    if (plan && B == 0) return 0;
    if (!plan || B < 0 || !coords || !axial || !Kgb) return bfe2_fail(-1, "bfe2_assemble_geometric: bad argument");
    PlanDev pd;
    if (const int r = bfe2_plan_device_view(plan, &pd)) return r;
    const long long total = B * pd.nslots;
    k_assemble_geometric<<<(unsigned)((total + 127) / 128), 128, 0, (cudaStream_t)stream>>>(pd, B, coords, axial, Kgb);
    if (cudaGetLastError() != cudaSuccess) return bfe2_fail(-100, "k_assemble_geometric launch failed");
    return 0;
}


int bfe2_internal_actions_ex(bfe2_plan* plan, int64_t B, int32_t C, int32_t npts, const double* coords,
                             const double* endforces, const double* q_axial, const double* q_perp, int32_t n_conc,
                             const double* conc, double* actions, double* maxabs, void* stream) {
    if (plan && B == 0) return 0;
    if (!plan || B < 0 || C < 1 || npts < 1 || !coords || !endforces || (!actions && !maxabs) || n_conc < 0 ||
        (n_conc > 0 && !conc))
        return bfe2_fail(-1, "bfe2_internal_actions: bad argument");
    PlanDev pd;
    if (const int r = bfe2_plan_device_view(plan, &pd)) return r;
    const PlanDev* P = &pd;
    const long long total = B * C * P->nE;
    k_internal_actions<<<(unsigned)((total + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        *P, B, C, npts, coords, endforces, q_axial, q_perp, n_conc, conc, actions, maxabs);
    if (cudaGetLastError() != cudaSuccess) return bfe2_fail(-100, "k_internal_actions launch failed");
    return 0;
}

int bfe2_internal_actions(bfe2_plan* plan, int64_t B, int32_t C, int32_t npts, const double* coords,
                          const double* endforces, const double* q_axial, const double* q_perp, double* actions,
                          double* maxabs, void* stream) {
    return bfe2_internal_actions_ex(plan, B, C, npts, coords, endforces, q_axial, q_perp, 0, nullptr, actions, maxabs,
                                    stream);
}

int bfe2_modal_participation(bfe2_plan* plan, int64_t B, int32_t n_modes, const double* Mb, const double* phi,
                             double* gamma, void* stream) {
    if (plan && B == 0) return 0;
    if (!plan || B < 0 || n_modes < 1 || !Mb || !phi || !gamma) return bfe2_fail(-1, "bfe2_modal_participation: bad argument");
    PlanDev pd;
    if (const int r = bfe2_plan_device_view(plan, &pd)) return r;
    const PlanDev* P = &pd;
    const long long total = B * n_modes;
    k_participation<<<(unsigned)((total + 127) / 128), 128, 0, (cudaStream_t)stream>>>(*P, B, n_modes, Mb, phi, gamma);
    if (cudaGetLastError() != cudaSuccess) return bfe2_fail(-100, "k_participation launch failed");
    return 0;
}

}  // extern "C"
