// bfe2_tile.cu -- static-only fused path, tile kernel: ONE LANE PER STRUCTURE.  OPT-IN (BFE2_STATIC_TILE=1): measured
// at the speed of the group kernel for one load case and behind it for more (profiles/r2_static_tile.md) -- an SM can
// hold only 44-64 structures whatever the lane mapping, and two warps per SM cannot hide their own latencies.
//
//   The group kernel (k_static_group, bfe2.cu) gives a structure 16 lanes and runs the substitutions on one of them:
//   its warp instructions serve two structures and it sits on the shared-memory pipe of the SM (profiles/
//   r1_static_group.txt: 4.3 k warp instructions and 1.7 k shared-memory wavefronts per structure).  Here a warp
//   owns a TILE of 32 structures and lane s carries structure s through every sequential stage -- assembly
//   (LinearStaticSolver / Structure.compile_global_matrix, Structure.py:379-393 + helpers.py:17-34), the band
//   factorisation and the substitutions (Solver.py:62-66), the end forces and reactions (Results.py:48-127) -- so one
//   warp instruction serves 32 structures and every branch is warp-uniform (the topology plan is shared).
//
//   Shared memory: slot k of lane s lives at smem[k * 33 + s] (pitch 33: lane-per-structure accesses AND the
//   cooperative structure-by-structure copies are bank-conflict free).  Slots of a lane:
//       band [W n]  column-major, slot j W + d = K[j+d][j]; after the factorisation 1/D[j] and the unit-lower L
//       trash       sink for the stiffness entries of supported DOFs
//       x [C n]     right-hand sides, then displacements on the free DOFs
//       zero, xtrash
//   and once the substitutions are done the band region is the staging area of the reactions ([sumdof] per lane).
//   Per-structure inputs are read straight from HBM by their lane in 16-byte pieces that are used completely
//   (a node's two coordinates, a beam's four properties, a beam's six fixed-end forces); f_nodal, u and the
//   reactions -- three doubles per node, no 16-byte alignment -- go through cooperative, coalesced copies instead.
//   The next tile's input records are prefetched into L2 while this one is worked on.
//
//   Factorisation: root-free (L D L^T), the (HB+1) x (HB+1) window of live columns in registers with static indices
//   (column j in register set j mod (HB+1)); the substitutions have ONE dependent FMA per row on their critical
//   path (unit-lower L; the division by D happens off the chain).
#include "bfe2_tile.h"

#include <atomic>
#include <cstdlib>

#include "../../include/bfe2.h"

using namespace bfe2;

int bfe2_fail(int code, const char* fmt, ...);

namespace {

constexpr int PITCH = 33;
constexpr int NPAIR = 21;          // lower triangle of a beam's 6 x 6 matrix

struct TileLayout { int trash, x, zero, xtrash, slots; };

__host__ __device__ inline TileLayout tile_layout(int n, int W, int C) {
    TileLayout L;
    L.trash = W * n;
    L.x = L.trash + 1;
    L.zero = L.x + C * n;
    L.xtrash = L.zero + 1;
    L.slots = L.xtrash + 1;
    return L;
}

__device__ __forceinline__ double2 ldg2(const double* p) { return __ldg(reinterpret_cast<const double2*>(p)); }

__device__ __forceinline__ void prefetch_l2(const double* base, long long bytes, int lane) {
    const char* p = reinterpret_cast<const char*>(base);
    for (long long off = (long long)lane * 128; off < bytes; off += 32 * 128)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p + off));
}

struct BeamIn { double2 ci, cj, p0, p1, r01, r23, r45; };     // node coordinates, (E, A), (I, rho), fixed-end forces

// p_r: the beam's six fixed-end forces of the load case that is worked on (NULL: no beam loads)
__device__ __forceinline__ BeamIn load_beam(const double* p_c, const double* p_p, const double* p_r, const int* c_conn,
                                            int e) {
    BeamIn b;
    b.ci = ldg2(p_c + 2 * c_conn[2 * e]);
    b.cj = ldg2(p_c + 2 * c_conn[2 * e + 1]);
    b.p0 = ldg2(p_p + 4 * e);
    b.p1 = ldg2(p_p + 4 * e + 2);
    b.r01 = b.r23 = b.r45 = make_double2(0.0, 0.0);
    if (p_r != nullptr) { b.r01 = ldg2(p_r + 6 * e); b.r23 = ldg2(p_r + 6 * e + 2); b.r45 = ldg2(p_r + 6 * e + 4); }
    return b;
}

struct BeamGeo { double iL, cs, sn, a, eil; }; // 1/L, direction, EA/L, EI/L

__device__ __forceinline__ BeamGeo beam_geo(const BeamIn& b) {
    BeamGeo g;
    const double dx = b.cj.x - b.ci.x, dy = b.cj.y - b.ci.y;
    g.iL = rsqrt_fast(fma(dx, dx, dy * dy));
    g.cs = dx * g.iL;
    g.sn = dy * g.iL;
    g.a = (b.p0.x * b.p0.y) * g.iL;
    g.eil = (b.p0.x * b.p1.x) * g.iL;
    return g;
}

template <int HB>
__global__ void __launch_bounds__(32) k_static_tile(PlanDev P, bfe2_tile_args A) {
    extern __shared__ __align__(16) double smem[];
    constexpr int W = HB + 1;
    const int lane = threadIdx.x;
    const int n = P.n, nE = P.nE, nN = P.nN, sumdof = P.sumdof, C = A.C;
    const TileLayout TL = tile_layout(n, W, C);
    // plan cache: connectivity, free numbering, band slot of each lower-triangle entry of each beam (x PITCH), free
    // index of each beam DOF
    int* c_conn = reinterpret_cast<int*>(smem + (size_t)TL.slots * PITCH);
    int* c_fop = c_conn + 2 * nE;
    int* c_bslot = c_fop + sumdof;
    int* c_xfree = c_bslot + NPAIR * nE;
    for (int i = lane; i < 2 * nE; i += 32) c_conn[i] = P.conn[i];
    for (int i = lane; i < sumdof; i += 32) c_fop[i] = P.free_of_pos[i];
    __syncwarp();
    for (int i = lane; i < 6 * nE; i += 32) {
        const int e = i / 6, k = i - 6 * e;
        c_xfree[i] = c_fop[3 * c_conn[2 * e + k / 3] + k % 3];
    }
    __syncwarp();
    for (int i = lane; i < NPAIR * nE; i += 32) {
        const int e = i / NPAIR, t = i - NPAIR * e;
        int r = 0;
        while ((r + 1) * (r + 2) / 2 <= t) ++r;                  // t = r (r + 1) / 2 + c, 0 <= c <= r
        const int c = t - r * (r + 1) / 2;
        const int fr = c_xfree[6 * e + r], fc = c_xfree[6 * e + c];
        const int hi = fr > fc ? fr : fc, lo = fr > fc ? fc : fr;
        c_bslot[i] = ((lo < 0) ? TL.trash : lo * W + (hi - lo)) * PITCH;
    }
    double* mine = smem + lane;
    mine[TL.zero * PITCH] = 0.0;
    __syncwarp();
    const long long B = A.B, ntiles = (B + 31) / 32;
    const int nf = C * sumdof, nr = 6 * C * nE;
    const bool have_r = A.r_local != nullptr;
    const bool want_post = A.endforces != nullptr || A.reactions != nullptr;
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);

    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long b0 = tile * 32;
        const bool valid = b0 + lane < B;
        const long long b = valid ? b0 + lane : B - 1;         // idle lanes redo the last structure, no stores
        {   // the next tile of this block: input records into L2
            const long long nt = tile + gridDim.x;
            if (nt < ntiles) {
                const long long first = nt * 32, cnt = (B - first < 32) ? B - first : 32;
                prefetch_l2(A.coords + first * 2 * nN, cnt * 16 * nN, lane);
                prefetch_l2(A.props + first * 4 * nE, cnt * 32 * nE, lane);
                prefetch_l2(A.f_nodal + first * nf, cnt * 8 * nf, lane);
                if (have_r) prefetch_l2(A.r_local + first * nr, cnt * 8 * nr, lane);
            }
        }
        const double* p_c = A.coords + b * 2 * nN;
        const double* p_p = A.props + b * 4 * nE;
        const double* p_r = have_r ? A.r_local + b * (long long)nr : nullptr;
        BeamIn nxt = load_beam(p_c, p_p, p_r, c_conn, 0);
        // 1. empty band
#pragma unroll 6
        for (int k = 0; k < W * n; ++k) mine[k * PITCH] = 0.0;
        // 2. nodal loads into x, structure by structure (coalesced reads, conflict-free writes)
        for (int c = 0; c < C; ++c)
            for (int pos = lane; pos < sumdof; pos += 32) {
                const int j = c_fop[pos];
                double* dst = smem + (size_t)(TL.x + c * n + (j >= 0 ? j : 0)) * PITCH;
                const double* src = A.f_nodal + b0 * nf + c * sumdof + pos;
                const int last = (int)((B - 1 - b0 < 31) ? B - 1 - b0 : 31);
#pragma unroll 8
                for (int s = 0; s < 32; ++s) {
                    const double v = __ldg(src + (long long)(s <= last ? s : last) * nf);
                    if (j >= 0) dst[s] = v;
                }
            }
        __syncwarp();
        // 3. assembly: beam after beam, the 21 band entries of a beam are distinct slots (or the trash slot)
        for (int e = 0; e < nE; ++e) {
            const BeamIn in = nxt;
            if (e + 1 < nE) nxt = load_beam(p_c, p_p, p_r, c_conn, e + 1);      // load case 0 travels with the beam
            const BeamGeo g = beam_geo(in);
            const double b6 = 6.0 * g.eil * g.iL, b12 = 2.0 * b6 * g.iL;
            const double kp = fma(g.a * g.cs, g.cs, (b12 * g.sn) * g.sn);
            const double kq = ((g.a - b12) * g.cs) * g.sn;
            const double kr = fma(g.a * g.sn, g.sn, (b12 * g.cs) * g.cs);
            const double kt = b6 * g.sn, kw = b6 * g.cs, k4 = 4.0 * g.eil, k2 = 2.0 * g.eil;
            // T k T^T (HermitianBeam_2D.py:173-175, :284-297), rows r = 0..5, columns c <= r
            const double kv[NPAIR] = {kp,  kq,  kr,  -kt, kw, k4, -kp, -kq, kt,  kp, -kq,
                                      -kr, -kw, kq,  kr,  -kt, kw, k2, kt,  -kw, k4};
            int sl[NPAIR];
            double v[NPAIR];
#pragma unroll
            for (int t = 0; t < NPAIR; ++t) sl[t] = c_bslot[NPAIR * e + t];
#pragma unroll
            for (int t = 0; t < NPAIR; ++t) v[t] = mine[sl[t]];
#pragma unroll
            for (int t = 0; t < NPAIR; ++t) mine[sl[t]] = v[t] + kv[t];
            if (have_r) {
                // fixed-end forces of the beam loads, rotated to global axes, join the nodal loads (Structure.py:261-276)
                for (int c = 0; c < C; ++c) {
                    const double* r = p_r + (c * nE + e) * 6;
                    const double2 r01 = c == 0 ? in.r01 : ldg2(r), r23 = c == 0 ? in.r23 : ldg2(r + 2),
                                  r45 = c == 0 ? in.r45 : ldg2(r + 4);
                    const double add[6] = {g.cs * r01.x - g.sn * r01.y, g.sn * r01.x + g.cs * r01.y, r23.x,
                                           g.cs * r23.y - g.sn * r45.x, g.sn * r23.y + g.cs * r45.x, r45.y};
                    int xs[6];
                    double xv[6];
#pragma unroll
                    for (int k = 0; k < 6; ++k) {
                        const int j = c_xfree[6 * e + k];
                        xs[k] = (j >= 0 ? TL.x + c * n + j : TL.xtrash) * PITCH;
                    }
#pragma unroll
                    for (int k = 0; k < 6; ++k) xv[k] = mine[xs[k]];
#pragma unroll
                    for (int k = 0; k < 6; ++k) mine[xs[k]] = xv[k] + add[k];
                }
            }
        }
        // 4. K = L D L^T in place.  w[a][d] = current A[col + d][col] of the live column col = a (mod W).
        int info = 0;
        {
            double w[W][W], orig[W];
#pragma unroll
            for (int a = 0; a < W; ++a) {
#pragma unroll
                for (int d = 0; d < W; ++d) w[a][d] = (a < n) ? mine[(a * W + d) * PITCH] : (d == 0 ? 1.0 : 0.0);
                orig[a] = w[a][0];
            }
            for (int j0 = 0; j0 < n; j0 += W) {
#pragma unroll
                for (int k = 0; k < W; ++k) {
                    const int col = j0 + k;
                    if (col < n) {
                        const double piv = w[k][0];
                        if (info == 0 && !(piv > 1.4210854715202004e-14 * orig[k] && piv < INFINITY)) info = col + 1;
                        const double ip = rcp_fast(piv);
                        double l[W];
#pragma unroll
                        for (int d = 1; d < W; ++d) l[d] = w[k][d] * ip;
#pragma unroll
                        for (int q = 1; q < W; ++q)
#pragma unroll
                            for (int p = q; p < W; ++p)
                                w[(k + q) % W][p - q] = fma(-l[p], w[k][q], w[(k + q) % W][p - q]);
                        double* cj = mine + (size_t)col * W * PITCH;
                        cj[0] = ip;
#pragma unroll
                        for (int d = 1; d < W; ++d) cj[d * PITCH] = l[d];
                        if (col + W < n) {
#pragma unroll
                            for (int d = 0; d < W; ++d) w[k][d] = cj[(W * W + d) * PITCH];
                            orig[k] = w[k][0];
                        }
                    }
                }
            }
        }
        // 5. substitutions, one load case after the other.  L's columns are fetched two rows ahead of their use.
        for (int c = 0; c < C; ++c) {
            double* x = mine + (size_t)(TL.x + c * n) * PITCH;
            double win[W];                                       // win[(col + d) % W] = running x[col + d]
#pragma unroll
            for (int d = 0; d < W; ++d) win[d] = (d < n) ? x[d * PITCH] : 0.0;
            double pre[2][W];
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int d = 0; d < W; ++d) pre[h][d] = (h < n) ? mine[(h * W + d) * PITCH] : 0.0;
            for (int j0 = 0; j0 < n; j0 += W) {
#pragma unroll
                for (int k = 0; k < W; ++k) {
                    const int col = j0 + k;
                    if (col < n) {
                        double l[W];
#pragma unroll
                        for (int d = 0; d < W; ++d) l[d] = pre[k & 1][d];
                        if (col + 2 < n) {
#pragma unroll
                            for (int d = 0; d < W; ++d) pre[k & 1][d] = mine[((col + 2) * W + d) * PITCH];
                        }
                        const double tail = (col + W < n) ? x[(col + W) * PITCH] : 0.0;
                        const double y = win[k];
#pragma unroll
                        for (int d = 1; d < W; ++d) win[(k + d) % W] = fma(-l[d], y, win[(k + d) % W]);
                        win[k] = tail;
                        x[col * PITCH] = y * l[0];               // z = D^-1 y
                    }
                }
            }
            static_assert(W % 2 == 0, "the two-ahead prefetch alternates on the column parity");
            const int top = (n + W - 1) / W * W;                 // columns >= n do not exist: their u stays 0
            double uw[W];
#pragma unroll
            for (int d = 0; d < W; ++d) uw[d] = 0.0;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int col = n - 1 - h;                       // parity slot of column col is col & 1
#pragma unroll
                for (int d = 0; d < W; ++d) {
                    const double v = (col >= 0) ? mine[(col * W + d) * PITCH] : 0.0;
                    if (((n - 1 - h) & 1) == 0) pre[0][d] = v; else pre[1][d] = v;
                }
            }
            for (int j0 = top - W; j0 >= 0; j0 -= W) {
#pragma unroll
                for (int k = W - 1; k >= 0; --k) {
                    const int col = j0 + k;
                    if (col < n) {
                        double l[W];
#pragma unroll
                        for (int d = 0; d < W; ++d) l[d] = pre[k & 1][d];
                        const double z = x[col * PITCH];
                        if (col - 2 >= 0) {
#pragma unroll
                            for (int d = 0; d < W; ++d) pre[k & 1][d] = mine[((col - 2) * W + d) * PITCH];
                        }
                        double acc = z;
#pragma unroll
                        for (int d = W - 1; d >= 1; --d) acc = fma(-l[d], uw[(k + d) % W], acc);   // u[col+1] joins last
                        uw[k] = acc;
                        x[col * PITCH] = acc;
                    }
                }
            }
        }
        const bool ok = info == 0;
        if (valid) A.info[b] = info;
        const unsigned bad = __ballot_sync(0xffffffffu, info != 0);   // lanes whose structure has no factorisation
        __syncwarp();
        // 6. displacements on all DOFs (zeros at supports), structure by structure
        for (int c = 0; c < C; ++c)
            for (int pos = lane; pos < sumdof; pos += 32) {
                const int j = c_fop[pos];
                const double* src = smem + (size_t)(j >= 0 ? TL.x + c * n + j : TL.zero) * PITCH;
                double* dst = A.u + b0 * nf + c * sumdof + pos;
#pragma unroll 8
                for (int s = 0; s < 32; ++s) {
                    const bool ok_s = ((bad >> s) & 1u) == 0;
                    const double v = (j >= 0) ? src[s] : 0.0;
                    if (b0 + s < B) dst[(long long)s * nf] = ok_s ? v : qnan;
                }
            }
        if (!want_post) { __syncwarp(); continue; }
        // 7./8. local end forces Ke u_local - r_local (Results.py:48-79) and, summed over the incident beams in
        //       global axes, the reactions (Results.py:92-127); the reactions collect in the (dead) band region
        for (int c = 0; c < C; ++c) {
            __syncwarp();
#pragma unroll 7
            for (int k = 0; k < sumdof; ++k) mine[k * PITCH] = 0.0;
            double* g_ef = A.endforces ? A.endforces + b * (long long)nr + (long long)c * nE * 6 : nullptr;
            const double* p_rc = have_r ? p_r + (long long)c * nE * 6 : nullptr;
            nxt = load_beam(p_c, p_p, p_rc, c_conn, 0);
            for (int e = 0; e < nE; ++e) {
                const BeamIn in = nxt;
                if (e + 1 < nE) nxt = load_beam(p_c, p_p, p_rc, c_conn, e + 1);
                const double2 r01 = in.r01, r23 = in.r23, r45 = in.r45;
                const BeamGeo g = beam_geo(in);
                const double b6 = 6.0 * g.eil * g.iL, b12 = 2.0 * b6 * g.iL, b4 = 4.0 * g.eil, b2 = 2.0 * g.eil;
                double uu[6];
#pragma unroll
                for (int k = 0; k < 6; ++k) {
                    const int j = c_xfree[6 * e + k];
                    uu[k] = mine[(j >= 0 ? TL.x + c * n + j : TL.zero) * PITCH];
                }
                const double d0 = g.cs * uu[0] + g.sn * uu[1], d1 = g.cs * uu[1] - g.sn * uu[0];
                const double d3 = g.cs * uu[3] + g.sn * uu[4], d4 = g.cs * uu[4] - g.sn * uu[3];
                double q[6];
                q[0] = g.a * (d0 - d3);
                q[3] = -q[0];
                q[1] = fma(b12, d1 - d4, b6 * (uu[2] + uu[5]));
                q[4] = -q[1];
                q[2] = fma(b6, d1 - d4, fma(b4, uu[2], b2 * uu[5]));
                q[5] = fma(b6, d1 - d4, fma(b2, uu[2], b4 * uu[5]));
                q[0] -= r01.x; q[1] -= r01.y; q[2] -= r23.x; q[3] -= r23.y; q[4] -= r45.x; q[5] -= r45.y;
                if (g_ef != nullptr && valid) {
                    double2* o = reinterpret_cast<double2*>(g_ef + 6 * e);
                    o[0] = ok ? make_double2(q[0], q[1]) : make_double2(qnan, qnan);
                    o[1] = ok ? make_double2(q[2], q[3]) : make_double2(qnan, qnan);
                    o[2] = ok ? make_double2(q[4], q[5]) : make_double2(qnan, qnan);
                }
                if (A.reactions != nullptr) {
                    const int pi = 3 * c_conn[2 * e], pj = 3 * c_conn[2 * e + 1];
                    const double add[6] = {g.cs * q[0] - g.sn * q[1], g.sn * q[0] + g.cs * q[1], q[2],
                                           g.cs * q[3] - g.sn * q[4], g.sn * q[3] + g.cs * q[4], q[5]};
                    double rv[6];
#pragma unroll
                    for (int k = 0; k < 3; ++k) { rv[k] = mine[(pi + k) * PITCH]; rv[3 + k] = mine[(pj + k) * PITCH]; }
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        mine[(pi + k) * PITCH] = rv[k] + add[k];
                        mine[(pj + k) * PITCH] = rv[3 + k] + add[3 + k];
                    }
                }
            }
            __syncwarp();
            if (A.reactions != nullptr) {
                for (int pos = lane; pos < sumdof; pos += 32) {
                    const double* src = smem + (size_t)pos * PITCH;
                    const double* fsrc = A.f_nodal + b0 * nf + c * sumdof + pos;
                    double* dst = A.reactions + b0 * nf + c * sumdof + pos;
                    const int last = (int)((B - 1 - b0 < 31) ? B - 1 - b0 : 31);
#pragma unroll 8
                    for (int s = 0; s < 32; ++s) {
                        const bool ok_s = ((bad >> s) & 1u) == 0;
                        const double fn = __ldg(fsrc + (long long)(s <= last ? s : last) * nf);
                        const double v = src[s] - fn;
                        if (s <= last) dst[(long long)s * nf] = ok_s ? v : qnan;
                    }
                }
            }
        }
        __syncwarp();
    }
}

std::atomic<long long> g_launches{0};

}  // namespace

extern "C" int64_t bfe2_static_tile_launches(void) { return g_launches.load(); }

int bfe2_tile_static_launch(const PlanDev& pd, const bfe2_tile_args& a, cudaStream_t stream) {
    constexpr int HB = 5, W = HB + 1;
    static const bool on = getenv("BFE2_STATIC_TILE") != nullptr;        // opt-in (see the header of this file)
    if (!on) return 1;
    if (pd.hbw > HB || pd.nE < 1 || pd.n < 1 || a.C < 1 || pd.sumdof > W * pd.n) return 1;
    static const long long min_batch = getenv("BFE2_STATIC_TILE_MIN") ? atoll(getenv("BFE2_STATIC_TILE_MIN")) : 8192;
    if (a.B < min_batch) return 1;                             // too few tiles to fill the machine: group kernel
    auto misaligned = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) != 0; };
    if (misaligned(a.coords) || misaligned(a.props) || misaligned(a.r_local) || misaligned(a.endforces)) return 1;
    const TileLayout TL = tile_layout(pd.n, W, a.C);
    const size_t ints = (size_t)(2 * pd.nE + pd.sumdof + NPAIR * pd.nE + 6 * pd.nE);
    const size_t bytes = (size_t)TL.slots * PITCH * sizeof(double) + ints * sizeof(int);
    if (bytes + 1024 > (227 * 1024) / 2) return 1;             // two tiles per SM or not at all
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
        return bfe2_fail(-100, "bfe2_tile_static_launch: device query failed");
    auto kern = k_static_tile<HB>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess)
        return bfe2_fail(-100, "bfe2_tile_static_launch: shared memory size refused");
    const long long ntiles = (a.B + 31) / 32;
    const int grid = (int)(ntiles < 2LL * sms ? ntiles : 2LL * sms);
    kern<<<grid, 32, bytes, stream>>>(pd, a);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return bfe2_fail(-100, "k_static_tile launch failed: %s", cudaGetErrorString(e));
    ++g_launches;
    return 0;
}
