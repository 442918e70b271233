// bfe2.cu -- kernels + C ABI of libbfe2.so (see include/bfe2.h).  sm_100a only, no CPU fallback.
#include <cuda_runtime.h>

#include <algorithm>
#include <climits>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <atomic>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/bfe2.h"
#include "bfe2_device.cuh"
#include "bfe2_tile.h"

using namespace bfe2;

// =================================================================================================
// error plumbing
// =================================================================================================
// This file is compiled twice (Makefile): as itself -- everything except the packed-record entry points -- and, with
// BFE2_TU_PACKED defined, through bfe2_packed.cu, which keeps only bfe2_pattern_* / bfe2_solve_fused_packed and the
// PACKED instantiations of k_solve.  Two translation units compile side by side; the kernels are templates in an
// anonymous namespace, so each unit instantiates only what its entry points launch.
int bfe2_fail(int code, const char* fmt, ...);

#if !defined(BFE2_TU_PACKED) && !defined(BFE2_TU_TE)
namespace {
thread_local std::string g_err = "";

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

}  // namespace

// shared with bfe2_beam.cu, bfe2_post.cu and bfe2_packed.cu
int bfe2_fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
#else
#define fail bfe2_fail
#endif

namespace {
#define BFE2_CUDA(call)                                                                         \
    do {                                                                                        \
        cudaError_t e_ = (call);                                                                \
        if (e_ != cudaSuccess) return fail(-100, "%s failed: %s", #call, cudaGetErrorString(e_)); \
    } while (0)
}  // namespace

// =================================================================================================
// topology plan (host)
// =================================================================================================
struct bfe2_plan {
    int nN = 0, nE = 0, sumdof = 0, n = 0, hbw = 0, nslots = 0, ncontrib = 0;
    int npad = 0, LD = 0, rpl = 0, lpp = 0, nthreads = 0;
    std::vector<int32_t> conn, free_of_pos, pos_of_free, fixed, slot_ptr, contrib, node_ptr, node_inc;
    std::vector<int32_t> contrib_k, contrib_m;   // contrib decoded: beam*K7_PITCH + K number | beam*M12_PITCH + M number, sign in bit 31
    std::vector<uint8_t> lumped;
    // device mirrors: uploaded lazily on the first launch on a device, one copy per device (a plan may be used from
    // several devices and host threads at once, e.g. one thread per GPU); every entry point takes its own PlanDev
    // by value under the mutex, the plan keeps no notion of a "current" device
    std::mutex mu;
    std::map<int, std::pair<void*, PlanDev>> copies;
};

namespace {

// Jacobi lane configuration for n free DOFs: RPL rows per lane, LPP lanes per column pair.
bool jacobi_config(int n, int* rpl, int* lpp) {
    if (n <= 10) { *rpl = 5; *lpp = 2; return true; }
    if (n <= 30) { *rpl = 15; *lpp = 2; return true; }
    if (n <= 58) { *rpl = 29; *lpp = 2; return true; }
    if (n <= 64) { *rpl = 32; *lpp = 2; return true; }    // still two columns per lane: 64 columns, 2 x 32 rows
    // shared-memory variant, four lanes per column pair: the smallest row count per lane that covers n
    if (n <= 72) { *rpl = 18; *lpp = 4; return true; }
    if (n <= 88) { *rpl = 22; *lpp = 4; return true; }
    if (n <= 100) { *rpl = 25; *lpp = 4; return true; }
    if (n <= 116 && n <= BFE2_MAX_JACOBI_N) { *rpl = 29; *lpp = 4; return true; }
    return false;
}

// Reverse Cuthill-McKee ordering of the node graph (beams = edges): returns the nodes in their new order.
std::vector<int> rcm_node_order(int n_nodes, int n_elem, const int32_t* conn) {
    std::vector<std::vector<int>> adj(n_nodes);
    for (int e = 0; e < n_elem; ++e) {
        adj[conn[2 * e]].push_back(conn[2 * e + 1]);
        adj[conn[2 * e + 1]].push_back(conn[2 * e]);
    }
    for (auto& a : adj) {
        std::sort(a.begin(), a.end());
        a.erase(std::unique(a.begin(), a.end()), a.end());
    }
    std::vector<int> order;
    std::vector<char> seen(n_nodes, 0);
    while ((int)order.size() < n_nodes) {
        int start = -1;                                 // unvisited node of minimum degree
        for (int k = 0; k < n_nodes; ++k)
            if (!seen[k] && (start < 0 || adj[k].size() < adj[start].size())) start = k;
        size_t head = order.size();
        order.push_back(start);
        seen[start] = 1;
        while (head < order.size()) {
            const int u = order[head++];
            std::vector<int> nb;
            for (int v : adj[u])
                if (!seen[v]) { nb.push_back(v); seen[v] = 1; }
            std::sort(nb.begin(), nb.end(), [&](int a, int b) {
                return adj[a].size() != adj[b].size() ? adj[a].size() < adj[b].size() : a < b;
            });
            order.insert(order.end(), nb.begin(), nb.end());
        }
    }
    std::reverse(order.begin(), order.end());
    return order;
}

int half_bandwidth(const bfe2_plan* p, const int32_t* conn) {
    int hbw = 0;
    for (int e = 0; e < p->nE; ++e) {
        int lo = 1 << 30, hi = -1;
        for (int a = 0; a < 6; ++a) {
            const int f = p->free_of_pos[3 * conn[2 * e + a / 3] + a % 3];
            if (f >= 0) { lo = std::min(lo, f); hi = std::max(hi, f); }
        }
        if (hi >= 0) hbw = std::max(hbw, hi - lo);
    }
    return hbw;
}

// Device view of the plan on the CURRENT device (uploads on first use there).
int plan_upload(bfe2_plan* p, PlanDev* out) {
    int dev = 0;
    BFE2_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(p->mu);
    auto it = p->copies.find(dev);
    if (it != p->copies.end()) {
        *out = it->second.second;
        return 0;
    }
    // one blob: all int32 arrays then the uint8 flags
    constexpr int NA = 9;
    const std::vector<int32_t>* arrs[NA] = {&p->conn, &p->free_of_pos, &p->pos_of_free, &p->slot_ptr, &p->contrib,
                                            &p->node_ptr, &p->node_inc, &p->contrib_k, &p->contrib_m};
    size_t off[NA + 1];
    size_t tot = 0;
    for (int i = 0; i < NA; ++i) { off[i] = tot; tot += ((arrs[i]->size() + 3) & ~size_t(3)); }
    off[NA] = tot;
    const size_t bytes = tot * sizeof(int32_t) + ((p->lumped.size() + 15) & ~size_t(15));
    std::vector<unsigned char> host(bytes, 0);
    for (int i = 0; i < NA; ++i)
        if (!arrs[i]->empty()) memcpy(host.data() + off[i] * 4, arrs[i]->data(), arrs[i]->size() * 4);
    if (!p->lumped.empty()) memcpy(host.data() + off[NA] * 4, p->lumped.data(), p->lumped.size());
    void* blob = nullptr;
    BFE2_CUDA(cudaMalloc(&blob, bytes));
    if (cudaMemcpy(blob, host.data(), bytes, cudaMemcpyHostToDevice) != cudaSuccess) {
        cudaFree(blob);
        return fail(-100, "plan upload failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    const int32_t* base = static_cast<const int32_t*>(blob);
    PlanDev d{};
    d.nN = p->nN; d.nE = p->nE; d.sumdof = p->sumdof; d.n = p->n; d.hbw = p->hbw; d.nslots = p->nslots;
    d.npad = p->npad; d.LD = p->LD;
    d.conn = base + off[0];
    d.free_of_pos = base + off[1];
    d.pos_of_free = base + off[2];
    d.slot_ptr = base + off[3];
    d.contrib = base + off[4];
    d.node_ptr = base + off[5];
    d.node_inc = base + off[6];
    d.contrib_k = base + off[7];
    d.contrib_m = base + off[8];
    d.lumped = reinterpret_cast<const uint8_t*>(base + off[NA]);
    p->copies[dev] = std::make_pair(blob, d);
    *out = d;
    return 0;
}

}  // namespace

// shared with bfe2_post.cu: device view of a plan on the current device (uploads lazily); < 0 + last_error on failure
#if !defined(BFE2_TU_PACKED) && !defined(BFE2_TU_TE)
int bfe2_plan_device_view(bfe2_plan* plan, PlanDev* out) { return plan_upload(plan, out); }
#else
int bfe2_plan_device_view(bfe2_plan* plan, PlanDev* out);
#endif

// =================================================================================================
// kernels
// =================================================================================================
namespace {

struct BatchArgs {
    long long B;
    int C, do_static, do_modal, n_modes, max_sweeps;
    const double *coords, *props, *nodal_mass, *f_nodal, *r_local;   // inputs
    const double *Kb_in, *Mb_in;                                     // unfused: bands from HBM
    double *Kb_out, *Mb_out;                                         // K2 outputs
    double *u, *endforces, *reactions, *lam, *phi;
    int32_t *info, *sweeps;
    // packed records (bfe2_solve_fused_packed; k_solve only).  w_in / w_out > 0: every input (output) pointer is the
    // address of that array inside structure 0's record and consecutive structures are w_in (w_out) doubles apart.
    // compact_in: f_nodal / r_local hold only the nnz_f / nnz_r values named by f_idx / r_idx (dense indices into
    // [C, sumdof] and [C, nE, 6]); compact_out: u and phi on the n free DOFs, reactions at the n_fixed supported ones.
    int w_in, w_out, compact_in, compact_out, nnz_f, nnz_r, n_fixed;
    const int32_t *f_idx, *r_idx, *fixed_pos;
    // only_flagged: this launch repairs an earlier one -- a block whose structure does not carry info ==
    // BFE2_INFO_RETRY (the two-ended modal engine declined it) leaves at once
    int only_flagged;
};

__device__ __forceinline__ void copy_in(double* dst, const double* src, int count, int tid, int nt) {
    for (int i = tid; i < count; i += nt) dst[i] = src[i];
}

__device__ __forceinline__ void fill_nan(double* dst, long long count, int tid, int nt) {
    if (dst == nullptr) return;
    const double q = __longlong_as_double(0x7ff8000000000000LL);
    for (long long i = tid; i < count; i += nt) dst[i] = q;
}

// ---- K1: one thread per element, SoA in, [nE,6,6] out staged through shared memory so that the
//      global stores are fully coalesced 16-byte vectors. ------------------------------------------
constexpr int K1_WARPS = 4;
constexpr int K1_PITCH = 38;   // doubles per element in the staging buffer (even: keeps 16 B alignment,
                               // 38 mod 16 = 6 -> at most 2-way bank conflicts on the staging writes)

__global__ void __launch_bounds__(K1_WARPS * 32)
k_element_km(long long nE, const double* __restrict__ xi, const double* __restrict__ yi,
             const double* __restrict__ xj, const double* __restrict__ yj, const double* __restrict__ E,
             const double* __restrict__ A, const double* __restrict__ I, const double* __restrict__ rho,
             const uint8_t* __restrict__ lumped, double* __restrict__ Kg, double* __restrict__ Mg) {
    __shared__ __align__(16) double stage[K1_WARPS][32 * K1_PITCH];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* buf = stage[warp];
    const long long warps_total = (long long)gridDim.x * K1_WARPS;
    for (long long base = ((long long)blockIdx.x * K1_WARPS + warp) * 32; base < nE; base += warps_total * 32) {
        const long long e = base + lane;
        const bool ok = e < nE;
        double L = 1.0, c = 1.0, s = 0.0, Ee = 0.0, Ae = 0.0, Ie = 0.0, re = 0.0;
        bool lump = false;
        if (ok) {
            beam_geometry(xi[e], yi[e], xj[e], yj[e], L, c, s);
            Ee = E[e]; Ae = A[e]; Ie = I[e]; re = rho[e];
            lump = lumped != nullptr && lumped[e] != 0;
        }
        const int cnt = (int)min((long long)32, nE - base);     // elements in this warp tile
        double X[6][6];
        // ---- K
        ke_local(X, Ee, Ae, Ie, L);
        rotate_to_global(X, c, s);
        // lane-major staging: element `lane` occupies doubles [lane*K1_PITCH, +36)
#pragma unroll
        for (int r = 0; r < 6; ++r)
#pragma unroll
            for (int q = 0; q < 6; ++q) buf[lane * K1_PITCH + 6 * r + q] = X[r][q];
        __syncwarp();
        {
            double2* dst = reinterpret_cast<double2*>(Kg + base * 36);
            const double2* src = reinterpret_cast<const double2*>(buf);
            for (int i = lane; i < cnt * 18; i += 32) {
                const int el = i / 18;
                dst[i] = src[el * (K1_PITCH / 2) + (i - el * 18)];
            }
        }
        __syncwarp();
        // ---- M
        me_local(X, re, Ae, L, lump);
        rotate_to_global(X, c, s);
#pragma unroll
        for (int r = 0; r < 6; ++r)
#pragma unroll
            for (int q = 0; q < 6; ++q) buf[lane * K1_PITCH + 6 * r + q] = X[r][q];
        __syncwarp();
        {
            double2* dst = reinterpret_cast<double2*>(Mg + base * 36);
            const double2* src = reinterpret_cast<const double2*>(buf);
            for (int i = lane; i < cnt * 18; i += 32) {
                const int el = i / 18;
                dst[i] = src[el * (K1_PITCH / 2) + (i - el * 18)];
            }
        }
        __syncwarp();
    }
}

// Beam stiffness and consistent mass in global axes in closed form (T X T^T, HermitianBeam_2D.py:173-175, of the
// local matrices :284-297 and :233-249): seven resp. twelve distinct numbers per beam,
//   K: p = a c^2 + b12 s^2, q = (a - b12) c s, r = a s^2 + b12 c^2, t = b6 s, w = b6 c, b4, b2
//   M: f = rho A L / 420;  ii/jj block: f (140 c^2 + 156 s^2), -16 f c s, f (140 s^2 + 156 c^2), 22 L f s, 22 L f c,
//      4 L^2 f;  ij block: f (70 c^2 + 54 s^2), 16 f c s, f (70 s^2 + 54 c^2), 13 L f s, 13 L f c, -3 L^2 f
// The plan carries, for every contribution e*36 + 6 r + c, which of these numbers entry (r, c) equals and its sign
// (h_ke7 / h_me12 in the plan builder).
// ---- K2: WARP-PER-STRUCTURE assembly: bands to HBM.  The integer plan is cached once per block in shared
//      memory, its contributions already decoded to (beam, number, sign); blocks are persistent (grid-stride over
//      structures).  A lane works out the 7 + 12 numbers of its beams straight from the structure's record, then
//      each lane owns band slots lane, lane+32, ... and adds their contributions in beam order (no atomics); the
//      sums go straight to HBM as coalesced 8-byte stores.  160 B of shared memory per beam: 64 warps per SM.
constexpr int K2_WARPS = 8;

__global__ void __launch_bounds__(K2_WARPS * 32)
k_assemble(PlanDev P, BatchArgs A, int ncontrib, int cached) {
    extern __shared__ __align__(16) double smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool want_mass = A.Mb_out != nullptr;
    const int per_warp = (K7_PITCH + M12_PITCH) * P.nE + 2 * P.nN;         // doubles: [nE][7] K numbers, [nE][13] M, masses
    const int nw = blockDim.x >> 5;                                        // warps (= structures in flight) per block
    // block-shared copy of the plan; a plan too large for it (cached == 0) is read from global memory / L2
    const int* c_conn = P.conn;
    const int* c_pos = P.pos_of_free;
    const int* c_ptr = P.slot_ptr;
    const int* c_conk = P.contrib_k;
    const int* c_conm = P.contrib_m;
    if (cached) {
        int* s_conn = reinterpret_cast<int*>(smem + (size_t)nw * per_warp);
        int* s_pos = s_conn + 2 * P.nE;
        int* s_ptr = s_pos + P.n;
        int* s_conk = s_ptr + P.nslots + 1;
        int* s_conm = s_conk + ncontrib;
        for (int i = tid; i < 2 * P.nE; i += blockDim.x) s_conn[i] = P.conn[i];
        for (int i = tid; i < P.n; i += blockDim.x) s_pos[i] = P.pos_of_free[i];
        for (int i = tid; i <= P.nslots; i += blockDim.x) s_ptr[i] = P.slot_ptr[i];
        for (int i = tid; i < ncontrib; i += blockDim.x) { s_conk[i] = P.contrib_k[i]; s_conm[i] = P.contrib_m[i]; }
        c_conn = s_conn; c_pos = s_pos; c_ptr = s_ptr; c_conk = s_conk; c_conm = s_conm;
    }
    __syncthreads();
    double* w_k = smem + (size_t)warp * per_warp;
    double* w_m = w_k + K7_PITCH * P.nE;
    double* w_mass = w_m + M12_PITCH * P.nE;
    const long long nwarps = (long long)gridDim.x * nw;
    for (long long b = (long long)blockIdx.x * nw + warp; b < A.B; b += nwarps) {
        const double* g_coords = A.coords + b * 2 * P.nN;
        const double* g_props = A.props + b * 4 * P.nE;
        const bool have_mass = want_mass && A.nodal_mass != nullptr;
        if (have_mass)                                 // in flight together with the beam inputs
            for (int i = lane; i < 2 * P.nN; i += 32) w_mass[i] = A.nodal_mass[b * 2 * P.nN + i];
        for (int e = lane; e < P.nE; e += 32) {
            const int ni = c_conn[2 * e], nj = c_conn[2 * e + 1];
            double L, cs, sn;
            beam_numbers(g_coords[2 * ni], g_coords[2 * ni + 1], g_coords[2 * nj], g_coords[2 * nj + 1], g_props[4 * e],
                         g_props[4 * e + 1], g_props[4 * e + 2], g_props[4 * e + 3], want_mass, P.lumped[e] != 0,
                         w_k + K7_PITCH * e, w_m + M12_PITCH * e, L, cs, sn);
        }
        __syncwarp();
        for (int slot = lane; slot < P.nslots; slot += 32) {
            const int p0 = c_ptr[slot], p1 = c_ptr[slot + 1];
            double k = 0.0, m = 0.0;
            auto add = [&](int p) {
                const int ck = c_conk[p];
                const double vk = w_k[ck & 0x7fffffff];
                k += (ck < 0) ? -vk : vk;
                if (want_mass) {
                    const int cm = c_conm[p];
                    const double vm = w_m[cm & 0x7fffffff];
                    m += (cm < 0) ? -vm : vm;
                }
            };
            if (p1 > p0) add(p0);                      // one or two contributions per slot on a chain: no loop
            if (p1 > p0 + 1) add(p0 + 1);
            for (int p = p0 + 2; p < p1; ++p) add(p);
            A.Kb_out[b * P.nslots + slot] = k;
            if (want_mass) {
                if (slot < P.n && have_mass) {
                    const int pos = c_pos[slot];
                    const int dof = pos % 3;
                    if (dof < 2) m += w_mass[2 * (pos / 3) + dof];
                }
                A.Mb_out[b * P.nslots + slot] = m;
            }
        }
        __syncwarp();
    }
}

// ---- K3 (unfused static solve), throughput variant: WARP PER STRUCTURE, persistent blocks, integer plan cached
//      in shared memory, warp-cooperative band Cholesky (fewer warp instructions than the one-thread version the
//      fused kernel uses for latency), substitutions one lane per load case, everything else lane-strided.
constexpr int K3_WARPS = 8;

__global__ void __launch_bounds__(K3_WARPS * 32)
k_static_warp(PlanDev P, BatchArgs A) {
    extern __shared__ __align__(16) double smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    const int n = P.n, C = A.C;
    const int per_warp = ((2 * P.nN + 4 * P.nE + 3 * P.nE + P.nslots + n + C * (n + 3 * P.nN + 6 * P.nE)) + 1) & ~1;
    // block-shared plan cache
    int* c_conn = reinterpret_cast<int*>(smem + (size_t)nw * per_warp);
    int* c_fop = c_conn + 2 * P.nE;
    int* c_pof = c_fop + P.sumdof;
    int* c_nptr = c_pof + n;
    int* c_ninc = c_nptr + P.nN + 1;
    for (int i = tid; i < 2 * P.nE; i += blockDim.x) { c_conn[i] = P.conn[i]; c_ninc[i] = P.node_inc[i]; }
    for (int i = tid; i < P.sumdof; i += blockDim.x) c_fop[i] = P.free_of_pos[i];
    for (int i = tid; i < n; i += blockDim.x) c_pof[i] = P.pos_of_free[i];
    for (int i = tid; i <= P.nN; i += blockDim.x) c_nptr[i] = P.node_ptr[i];
    __syncthreads();
    PlanDev Q = P;
    Q.conn = c_conn; Q.free_of_pos = c_fop; Q.pos_of_free = c_pof; Q.node_ptr = c_nptr; Q.node_inc = c_ninc;
    double* w_coords = smem + (size_t)warp * per_warp;
    double* w_props = w_coords + 2 * P.nN;
    double* w_geo = w_props + 4 * P.nE;
    double* w_Kb = w_geo + 3 * P.nE;
    double* w_inv = w_Kb + P.nslots;
    double* w_f = w_inv + n;
    double* w_ufull = w_f + C * n;
    double* w_ef = w_ufull + C * P.sumdof;
    const long long nwarps = (long long)gridDim.x * nw;
    for (long long b = (long long)blockIdx.x * nw + warp; b < A.B; b += nwarps) {
        for (int i = lane; i < 2 * P.nN; i += 32) w_coords[i] = A.coords[b * 2 * P.nN + i];
        for (int i = lane; i < 4 * P.nE; i += 32) w_props[i] = A.props[b * 4 * P.nE + i];
        for (int i = lane; i < P.nslots; i += 32) w_Kb[i] = A.Kb_in[b * P.nslots + i];
        __syncwarp();
        stage_geometry(Q, w_coords, w_geo, lane, 32);
        const int info = band_cholesky_warp(w_Kb, w_inv, n, P.hbw, lane);
        __syncwarp();
        if (info != 0) {
            if (lane == 0) A.info[b] = info;
            fill_nan(A.u + b * C * P.sumdof, (long long)C * P.sumdof, lane, 32);
            fill_nan(A.endforces ? A.endforces + b * C * P.nE * 6 : nullptr, (long long)C * P.nE * 6, lane, 32);
            fill_nan(A.reactions ? A.reactions + b * C * P.sumdof : nullptr, (long long)C * P.sumdof, lane, 32);
            __syncwarp();
            continue;
        }
        stage_static<true>(Q, C, w_props, w_geo, w_Kb, w_inv, w_f, w_ufull, w_ef, A.f_nodal + b * C * P.sumdof,
                           A.r_local ? A.r_local + b * C * P.nE * 6 : nullptr, A.u + b * C * P.sumdof,
                           A.endforces ? A.endforces + b * C * P.nE * 6 : nullptr,
                           A.reactions ? A.reactions + b * C * P.sumdof : nullptr, lane, 32);
        if (lane == 0) A.info[b] = 0;
        __syncwarp();
    }
}

// ---- K3 (narrow bands) and the static-only fused path: G lanes per structure -----------------------
//   A warp carries 32/G structures at once.  The stages whose work items are independent (beam
//   coefficients, band slots, right-hand sides, end forces, reactions, loads / stores) are spread over
//   the G lanes of the group; the two sequential stages run on lane 0 (register-window band Cholesky)
//   and lanes 0..C-1 (one substitution per load case) of EVERY group of the warp at the same time, so a
//   warp instruction of those stages serves 32/G structures instead of one.  No block barrier after the
//   plan cache is filled: groups only need __syncwarp.
//
//   Beam stiffness in global axes in closed form (the seven distinct numbers of T k T^T,
//   HermitianBeam_2D.py:173-175 / :284-297):
//     p = a c^2 + b12 s^2, q = (a - b12) c s, r = a s^2 + b12 c^2, t = b6 s, w = b6 c, b4, b2
//   c_ke7[6 r + c] = +-(1 + index into {p,q,r,t,w,b4,b2}).

struct GroupLayout { int a, kb, per_group; };

__host__ __device__ inline GroupLayout group_layout(int nE, int n, int band_w, int C, int G) {
    GroupLayout L;
    const int nslots = band_w * n;                // column-major band: band_w doubles per column
    const int size_a = (K7_PITCH * nE > C * n) ? K7_PITCH * nE : C * n;   // beam stiffness numbers | right-hand sides
    L.a = 5 * nE;                                 // after geo[nE][5] = 1/L, cos, sin, EA/L, EI/L
    L.kb = (L.a + size_a + 1) & ~1;               // 16-byte aligned: band [W n] + original diagonal [n] | end forces [C][nE][6]
    int tot = L.kb + ((nslots + n > 6 * C * nE) ? nslots + n : 6 * C * nE);
    // group stride = G (mod 16 doubles): the groups of a half-warp tile the banks in the lane-parallel stages;
    // G >= 16: stride = 2 (mod 16) keeps lane 0 of the warp's groups apart in the lane-serial stages
    const int want = G < 16 ? G : 2;
    tot += ((want - tot) % 16 + 16) % 16;
    L.per_group = tot;
    return L;
}

// 8-byte asynchronous global -> shared copies (LDGSTS): the unfused path puts the whole band of a structure in
// flight at once and it lands while the group works out the beam geometry.
__device__ __forceinline__ void cp_async8(double* dst, const double* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int G, int HB, bool ASSEMBLE>
__global__ void __launch_bounds__(640)
k_static_group(PlanDev P, BatchArgs A, int ncontrib) {
    extern __shared__ __align__(16) double smem[];
    const int tid = threadIdx.x, g = tid % G, gi = tid / G, ngb = blockDim.x / G;
    const int n = P.n, nE = P.nE, nN = P.nN, sumdof = P.sumdof, C = A.C, hbw = P.hbw, nslots = P.nslots;
    constexpr int W = BandCm<HB>::W;
    const GroupLayout GL = group_layout(nE, n, W, C, G);
    // block-shared plan cache (contributions already decoded to beam*8 + value index, sign in bit 31)
    int* c_conn = reinterpret_cast<int*>(smem + (size_t)ngb * GL.per_group);
    int* c_fop = c_conn + 2 * nE;
    int* c_pof = c_fop + sumdof;
    int* c_nptr = c_pof + n;
    int* c_ninc = c_nptr + nN + 1;
    int* c_sptr = c_ninc + 2 * nE;
    int* c_con = c_sptr + nslots + 1;
    for (int i = tid; i < 2 * nE; i += blockDim.x) { c_conn[i] = P.conn[i]; c_ninc[i] = P.node_inc[i]; }
    for (int i = tid; i < sumdof; i += blockDim.x) c_fop[i] = P.free_of_pos[i];
    for (int i = tid; i < n; i += blockDim.x) c_pof[i] = P.pos_of_free[i];
    for (int i = tid; i <= nN; i += blockDim.x) c_nptr[i] = P.node_ptr[i];
    if (ASSEMBLE) {
        for (int i = tid; i <= nslots; i += blockDim.x) c_sptr[i] = P.slot_ptr[i];
        for (int i = tid; i < ncontrib; i += blockDim.x) c_con[i] = P.contrib_k[i];
    }
    __syncthreads();
    double* s_geo = smem + (size_t)gi * GL.per_group;            // [nE][5] = 1/L, cos, sin, EA/L, EI/L
    double* s_a = s_geo + GL.a;                                   // beam stiffness numbers, then right-hand sides
    double* s_Kb = s_geo + GL.kb;                                 // column-major: s_Kb[j * W + d] = K[j + d][j]
    double* s_inv = s_Kb + W * n;                                 // original diagonal during the factorisation
    double* s_ef = s_Kb;                                          // the band is dead once u is known
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    const long long per_pass = (long long)gridDim.x * ngb;
    const long long passes = (A.B + per_pass - 1) / per_pass;
    const int nr = 6 * C * nE, nf = C * sumdof;
    const bool have_r = A.r_local != nullptr;
    auto structure_of = [&](long long it) {
        const long long b0 = (it * gridDim.x + blockIdx.x) * ngb + gi;
        return b0 < A.B ? b0 : A.B - 1;                           // idle groups redo the last structure, no stores
    };
    for (long long it = 0; it < passes; ++it) {
        const bool valid = (it * gridDim.x + blockIdx.x) * ngb + gi < A.B;
        const long long b = structure_of(it);
        const double* g_f = A.f_nodal + b * (long long)nf;
        const double* g_r = have_r ? A.r_local + b * (long long)nr : nullptr;
        const double* p_cp = A.coords + b * 2 * nN;
        const double* p_pr = A.props + b * 4 * nE;
        if (!ASSEMBLE) {                                          // unfused K3: the band comes from HBM
            const double* g_Kb = A.Kb_in + b * nslots;
            for (int d = 0; d <= hbw; ++d)
                for (int j = g; j < n; j += G) cp_async8(s_Kb + j * W + d, g_Kb + d * n + j);
            cp_async_commit();
        }
        // 0. (fused path) this lane's right-hand-side entries go into registers and r_local into the band region,
        //    which is free until the assembly: every load of the right-hand sides is in flight before the first
        //    use, instead of one dependent HBM round trip per incident beam.  Item k of a lane is the (load case,
        //    node) pair number t = g + G k = c nN + node; it carries the three DOFs of the node.
        constexpr int MAXI = 4;
        const bool early_rhs = ASSEMBLE && have_r && (C * nN + G - 1) / G <= MAXI && nr <= W * n + n;
        double xr[3 * MAXI];
        if (early_rhs) {
            for (int i = g; i < nr; i += G) cp_async8(s_Kb + i, g_r + i);
            cp_async_commit();
            int c = 0, node = g;
#pragma unroll
            for (int k = 0; k < MAXI; ++k) {
                while (node >= nN) { node -= nN; ++c; }
                const double* f3 = g_f + c * sumdof + 3 * node;
#pragma unroll
                for (int d = 0; d < 3; ++d) xr[3 * k + d] = (c < C) ? f3[d] : 0.0;
                node += G;
            }
        }
        // 1. beams: direction, 1/L, EA/L, EI/L and (fused path) the seven stiffness numbers
        for (int e = g; e < nE; e += G) {
            const int ni = c_conn[2 * e], nj = c_conn[2 * e + 1];
            const double dx = p_cp[2 * nj] - p_cp[2 * ni], dy = p_cp[2 * nj + 1] - p_cp[2 * ni + 1];
            const double* pr = p_pr + 4 * e;
            const double iL = rsqrt_fast(fma(dx, dx, dy * dy));
            const double cs = dx * iL, sn = dy * iL;
            const double a = (pr[0] * pr[1]) * iL, eil = (pr[0] * pr[2]) * iL;
            double* ge = s_geo + 5 * e;
            ge[0] = iL; ge[1] = cs; ge[2] = sn; ge[3] = a; ge[4] = eil;
            if (ASSEMBLE) {
                const double b6 = 6.0 * eil * iL, b12 = 2.0 * b6 * iL;
                double* k7 = s_a + K7_PITCH * e;
                k7[0] = fma(a * cs, cs, (b12 * sn) * sn);
                k7[1] = ((a - b12) * cs) * sn;
                k7[2] = fma(a * sn, sn, (b12 * cs) * cs);
                k7[3] = b6 * sn;
                k7[4] = b6 * cs;
                k7[5] = 4.0 * eil;
                k7[6] = 2.0 * eil;
            }
        }
        if (early_rhs) {
            cp_async_wait<0>();
            __syncwarp();
            int c = 0, node = g;
#pragma unroll
            for (int k = 0; k < MAXI; ++k) {
                while (node >= nN) { node -= nN; ++c; }
                if (c < C) {
                    for (int q = c_nptr[node]; q < c_nptr[node + 1]; ++q) {
                        const int inc = c_ninc[q];
                        const int e = inc >> 1;
                        const double* r = s_Kb + (c * nE + e) * 6 + 3 * (inc & 1);
                        const double cs = s_geo[5 * e + 1], sn = s_geo[5 * e + 2];
                        xr[3 * k] += cs * r[0] - sn * r[1];
                        xr[3 * k + 1] += sn * r[0] + cs * r[1];
                        xr[3 * k + 2] += r[2];
                    }
                }
                node += G;
            }
        }
        __syncwarp();
        // 2. band of K on the free DOFs
        if (ASSEMBLE) {
            for (int d = 0; d <= hbw; ++d)
            for (int j = g; j < n; j += G) {
                const int slot = d * n + j;
                const int q0 = c_sptr[slot], cnt = c_sptr[slot + 1] - q0;
                // the first two contributions without a loop (a chain has one or two per slot), same beam order
                double k = 0.0;
                if (cnt > 0) {
                    const int code = c_con[q0];
                    const double v = s_a[code & 0x7fffffff];
                    k = (code < 0) ? -v : v;
                }
                if (cnt > 1) {
                    const int code = c_con[q0 + 1];
                    const double v = s_a[code & 0x7fffffff];
                    k += (code < 0) ? -v : v;
                }
                for (int q = q0 + 2; q < q0 + cnt; ++q) {
                    const int code = c_con[q];
                    const double v = s_a[code & 0x7fffffff];
                    k += (code < 0) ? -v : v;
                }
                s_Kb[j * W + d] = k;
            }
        } else {
            cp_async_wait<0>();
        }
        for (int d = hbw + 1; d < W; ++d)                         // beyond the band (narrower than HB, padding)
            for (int j = g; j < n; j += G) s_Kb[j * W + d] = 0.0;
        __syncwarp();
        // 3. factor: lane 0 of each group
        const int info = band_cholesky_group_cm<G, HB>(s_Kb, s_inv, n, g);
        // 4. right-hand sides in free numbering (over the stiffness numbers, which are dead now)
        if (early_rhs) {
            int c = 0, node = g;
#pragma unroll
            for (int k = 0; k < MAXI; ++k) {
                while (node >= nN) { node -= nN; ++c; }
                if (c < C) {
#pragma unroll
                    for (int d = 0; d < 3; ++d) {
                        const int j = c_fop[3 * node + d];
                        if (j >= 0) s_a[c * n + j] = xr[3 * k + d];
                    }
                }
                node += G;
            }
        }
        for (int c = 0; c < (early_rhs ? 0 : C); ++c) {
            for (int node = g; node < nN; node += G) {
                const double* f3 = g_f + c * sumdof + 3 * node;
                double v0 = f3[0], v1 = f3[1], v2 = f3[2];
                if (have_r) {
                    for (int q = c_nptr[node]; q < c_nptr[node + 1]; ++q) {
                        const int inc = c_ninc[q];
                        const int e = inc >> 1;
                        const double* r = g_r + (c * nE + e) * 6 + 3 * (inc & 1);
                        const double cs = s_geo[5 * e + 1], sn = s_geo[5 * e + 2];
                        v0 += cs * r[0] - sn * r[1];
                        v1 += sn * r[0] + cs * r[1];
                        v2 += r[2];
                    }
                }
                const int j0 = c_fop[3 * node], j1 = c_fop[3 * node + 1], j2 = c_fop[3 * node + 2];
                if (j0 >= 0) s_a[c * n + j0] = v0;
                if (j1 >= 0) s_a[c * n + j1] = v1;
                if (j2 >= 0) s_a[c * n + j2] = v2;
            }
        }
        __syncwarp();
        // 5. substitute: one lane per load case
        for (int c = g; c < C; c += G) band_subst_cm<HB>(s_Kb, s_a + c * n, n);
        __syncwarp();
        const bool ok = info == 0;
        if (valid && g == 0) A.info[b] = info;
        // 6. displacements on all DOFs (exact zeros at supports)
        {
            double* g_u = A.u + b * (long long)nf;
            for (int c = 0; c < C; ++c)
                for (int pos = g; pos < sumdof; pos += G) {
                    const int j = c_fop[pos];
                    const double v = (j >= 0) ? s_a[c * n + j] : 0.0;
                    if (valid) g_u[c * sumdof + pos] = ok ? v : qnan;
                }
        }
        if (A.endforces == nullptr && A.reactions == nullptr) { __syncwarp(); continue; }
        // 7. local end forces  Ke u_local - r_local  (Results.py:48-79)
        double* g_ef = A.endforces ? A.endforces + b * (long long)nr : nullptr;
        for (int e = g; e < nE; e += G) {
            const int ni = c_conn[2 * e], nj = c_conn[2 * e + 1];
            const double* ge = s_geo + 5 * e;
            const double iL = ge[0], cs = ge[1], sn = ge[2], a = ge[3], eil = ge[4];
            const double b6 = 6.0 * eil * iL, b12 = 2.0 * b6 * iL, b4 = 4.0 * eil, b2 = 2.0 * eil;
            const int fi0 = c_fop[3 * ni], fi1 = c_fop[3 * ni + 1], fi2 = c_fop[3 * ni + 2];
            const int fj0 = c_fop[3 * nj], fj1 = c_fop[3 * nj + 1], fj2 = c_fop[3 * nj + 2];
            for (int c = 0; c < C; ++c) {
                const double* x = s_a + c * n;
                double rl[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};     // in flight while Ke u_local is formed
                if (have_r) {
                    const double* r = g_r + (c * nE + e) * 6;
#pragma unroll
                    for (int k = 0; k < 6; ++k) rl[k] = r[k];
                }
                const double ui0 = fi0 >= 0 ? x[fi0] : 0.0, ui1 = fi1 >= 0 ? x[fi1] : 0.0, ui2 = fi2 >= 0 ? x[fi2] : 0.0;
                const double uj0 = fj0 >= 0 ? x[fj0] : 0.0, uj1 = fj1 >= 0 ? x[fj1] : 0.0, uj2 = fj2 >= 0 ? x[fj2] : 0.0;
                const double d0 = cs * ui0 + sn * ui1, d1 = cs * ui1 - sn * ui0;
                const double d3 = cs * uj0 + sn * uj1, d4 = cs * uj1 - sn * uj0;
                double q[6];
                q[0] = a * (d0 - d3);
                q[3] = -q[0];
                q[1] = fma(b12, d1 - d4, b6 * (ui2 + uj2));
                q[4] = -q[1];
                q[2] = fma(b6, d1 - d4, fma(b4, ui2, b2 * uj2));
                q[5] = fma(b6, d1 - d4, fma(b2, ui2, b4 * uj2));
                double* d = s_ef + (c * nE + e) * 6;
#pragma unroll
                for (int k = 0; k < 6; ++k) q[k] -= rl[k];
#pragma unroll
                for (int k = 0; k < 6; ++k) d[k] = q[k];
                if (g_ef != nullptr && valid) {
                    double* o = g_ef + (c * nE + e) * 6;
#pragma unroll
                    for (int k = 0; k < 6; ++k) o[k] = ok ? q[k] : qnan;
                }
            }
        }
        __syncwarp();
        // 8. reactions per (load case, node): the rotated end forces of the incident beams minus the nodal loads
        //    (Results.py:92-127)
        if (A.reactions != nullptr) {
            double* g_re = A.reactions + b * (long long)nf;
            for (int c = 0; c < C; ++c)
                for (int node = g; node < nN; node += G) {
                    const double* f3 = g_f + c * sumdof + 3 * node;
                    const double fn0 = f3[0], fn1 = f3[1], fn2 = f3[2];     // in flight while the end forces are summed
                    double v0 = 0.0, v1 = 0.0, v2 = 0.0;
                    for (int q = c_nptr[node]; q < c_nptr[node + 1]; ++q) {
                        const int inc = c_ninc[q];
                        const int e = inc >> 1;
                        const double* f = s_ef + (c * nE + e) * 6 + 3 * (inc & 1);
                        const double cs = s_geo[5 * e + 1], sn = s_geo[5 * e + 2];
                        v0 += cs * f[0] - sn * f[1];
                        v1 += sn * f[0] + cs * f[1];
                        v2 += f[2];
                    }
                    if (valid) {
                        double* o = g_re + c * sumdof + 3 * node;
                        o[0] = ok ? v0 - fn0 : qnan;
                        o[1] = ok ? v1 - fn1 : qnan;
                        o[2] = ok ? v2 - fn2 : qnan;
                    }
                }
        }
        __syncwarp();
    }
}

template <int G, int HB, bool ASSEMBLE>
int launch_static_group(bfe2_plan* plan, const PlanDev& pd, const BatchArgs& A, cudaStream_t stream) {
    const GroupLayout GL = group_layout(plan->nE, plan->n, BandCm<HB>::W, A.C, G);
    const size_t cache = (size_t)(4 * plan->nE + plan->sumdof + plan->n + plan->nN + 1 + plan->nslots + 1 +
                                  plan->ncontrib) * sizeof(int) + 16;
    const size_t per_warp = (size_t)(32 / G) * GL.per_group * sizeof(double);
    // as many warps per SM as the shared memory holds, in blocks of at most 20 warps
    int nw = 0;
    for (int nblk = 1; nblk <= 8 && nw == 0; ++nblk) {
        const size_t avail = (227 * 1024) / nblk - 1024;
        if (avail <= cache) break;
        const int fit = (int)((avail - cache) / per_warp);
        if (fit <= 20) nw = fit;
    }
    if (nw < 1) return 1;                         // does not fit: the caller falls back to block-per-structure
    // small batches: do not leave SMs idle behind a few fat blocks
    int dev = 0, sms = 0;
    BFE2_CUDA(cudaGetDevice(&dev));
    BFE2_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    while (nw > 1 && (long long)(nw / 2) * (32 / G) * sms >= A.B) nw /= 2;
    const size_t bytes = nw * per_warp + cache;
    auto kern = k_static_group<G, HB, ASSEMBLE>;
    BFE2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    const long long per_block = (long long)nw * (32 / G);
    const int resident = std::max<int>(1, (int)((227 * 1024) / (bytes + 1024)));
    const int grid = (int)std::min<long long>((A.B + per_block - 1) / per_block, (long long)sms * resident);
    kern<<<grid, nw * 32, bytes, stream>>>(pd, A, plan->ncontrib);
    BFE2_CUDA(cudaGetLastError());
    return 0;
}

// 0: launched; 1: this plan does not fit the group kernel (wide band / too much shared memory); < 0: error
template <bool ASSEMBLE>
int dispatch_static_group(bfe2_plan* plan, const PlanDev& pd, const BatchArgs& A, cudaStream_t stream) {
    // 16 lanes per structure: measured best of 4 / 8 / 16 / 32 (profiles/r1_ablation.md)
    if (plan->hbw > 8 || plan->nE == 0) return 1;
    return plan->hbw <= 5 ? launch_static_group<16, 5, ASSEMBLE>(plan, pd, A, stream)
                          : launch_static_group<16, 8, ASSEMBLE>(plan, pd, A, stream);
}

// ---- K3 / K4 / fused: one block per structure ----------------------------------------------------
//   ASSEMBLE = 1: inputs are coords/props(/mass) and the bands are assembled in shared memory;
//   ASSEMBLE = 0: bands are read from HBM (unfused K3 / K4).
//   REGS = 1: register-resident Jacobi (n <= 64); 2: the two-column-group register variant (64 < n <= 116, fused
//   path); 0: columns exchanged through shared memory (stage_jacobi_oe).
//   PACKED = 1: packed records / compact layouts (bfe2_solve_fused_packed); a separate instantiation, so that the
//   dense kernel carries none of that code (the register-resident sweep loop is sensitive to what ptxas has to
//   keep around it: with run-time flags the dense P20 launch went from 46.0 to 49.9 ms).
template <int RPL, int LPP, int NT, int MINB, bool ASSEMBLE, int REGS, bool PACKED = false>
__global__ void __launch_bounds__(NT, MINB)
k_solve(PlanDev P, BatchArgs A) {
    extern __shared__ __align__(16) double smem[];
    const SmemLayout L = make_layout(P.nN, P.nE, P.n, P.hbw, P.npad, P.LD, A.C);
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
    const long long b = blockIdx.x;
    const int n = P.n;
    int* s_order = reinterpret_cast<int*>(smem + L.order);
    int* s_flag = s_order + P.npad;                  // [0] = K cholesky info, [1] = M cholesky info
    if (A.only_flagged && A.info[b * ((PACKED && A.w_out) ? 2LL * A.w_out : 1LL)] != BFE2_INFO_RETRY) return;

    // per-structure strides (doubles): natural sizes of the dense tensors, or the record widths when packed.  The
    // output addresses are formed where they are used (after the sweeps), so nothing extra stays live across them.
    const long long in_w = PACKED ? A.w_in : 0;
    const bool compact_out = PACKED && A.compact_out;
    const bool need_geo = ASSEMBLE || A.do_static;
    if (need_geo) {
        copy_in(smem + L.coords, A.coords + b * (in_w ? in_w : 2LL * P.nN), 2 * P.nN, tid, nt);
        copy_in(smem + L.props, A.props + b * (in_w ? in_w : 4LL * P.nE), 4 * P.nE, tid, nt);
    }
    if (ASSEMBLE && A.do_modal && A.nodal_mass != nullptr)
        copy_in(smem + L.mass, A.nodal_mass + b * (in_w ? in_w : 2LL * P.nN), 2 * P.nN, tid, nt);
    if (!ASSEMBLE) {
        copy_in(smem + L.Kb, A.Kb_in + b * P.nslots, P.nslots, tid, nt);
        if (A.do_modal) copy_in(smem + L.Mb, A.Mb_in + b * P.nslots, P.nslots, tid, nt);
    }
    if (tid < 2) s_flag[tid] = 0;
    __syncthreads();
    if (ASSEMBLE) {
        stage_elements_cf(P, smem + L.coords, smem + L.props, smem + L.geo, smem + L.U, smem + L.U + K7_PITCH * P.nE,
                          A.do_modal != 0, tid, nt);
        __syncthreads();
        stage_assemble_cf(P, smem + L.U, smem + L.U + K7_PITCH * P.nE,
                          (A.do_modal && A.nodal_mass) ? smem + L.mass : nullptr, smem + L.Kb, smem + L.Mb,
                          A.do_modal != 0, tid, nt);
        __syncthreads();
    } else if (A.do_static) {
        stage_geometry(P, smem + L.coords, smem + L.geo, tid, nt);
    }
    // band Cholesky of K (warp 0) and of M (warp 1)
    auto factor = [&](bool do_k, bool do_m) {
        if (P.hbw <= 8) {                            // narrow band: root-free elimination by the lanes of one warp
            if (warp == 0 && do_k) {
                const int r = P.hbw <= 5 ? band_cholesky_group<32, 5>(smem + L.Kb, smem + L.invK, n, P.hbw, lane)
                                         : band_cholesky_group<32, 8>(smem + L.Kb, smem + L.invK, n, P.hbw, lane);
                if (lane == 0) s_flag[0] = r;
            } else if (warp == 1 && do_m) {
                const int r = P.hbw <= 5 ? band_cholesky_group<32, 5>(smem + L.Mb, smem + L.invM, n, P.hbw, lane)
                                         : band_cholesky_group<32, 8>(smem + L.Mb, smem + L.invM, n, P.hbw, lane);
                if (lane == 0) s_flag[1] = r;
            }
        } else if (warp == 0 && do_k) {
            const int r = band_cholesky_warp(smem + L.Kb, smem + L.invK, n, P.hbw, lane);
            if (lane == 0) s_flag[0] = r;
        } else if (warp == 1 && do_m) {
            const int r = band_cholesky_warp(smem + L.Mb, smem + L.invM, n, P.hbw, lane);
            if (lane == 0) s_flag[1] = r;
        }
        __syncthreads();
    };
    struct Outs {
        double *u, *ef, *re, *lam, *phi;
        int32_t *info, *sweeps;
        int nu, nre;
    };
    auto outs = [&]() {
        Outs o;
        const long long out_w = PACKED ? A.w_out : 0;
        o.nu = compact_out ? n : P.sumdof;           // u / phi entries per load case / mode
        o.nre = compact_out ? A.n_fixed : P.sumdof;
        o.u = A.u ? A.u + b * (out_w ? out_w : (long long)A.C * o.nu) : nullptr;
        o.ef = A.endforces ? A.endforces + b * (out_w ? out_w : (long long)A.C * P.nE * 6) : nullptr;
        o.re = A.reactions ? A.reactions + b * (out_w ? out_w : (long long)A.C * o.nre) : nullptr;
        o.lam = A.lam ? A.lam + b * (out_w ? out_w : (long long)n) : nullptr;
        o.phi = A.phi ? A.phi + b * (out_w ? out_w : (long long)A.n_modes * o.nu) : nullptr;
        o.info = A.info + b * (out_w ? 2 * out_w : 1);
        o.sweeps = A.sweeps ? A.sweeps + b * (out_w ? 2 * out_w : 1) : nullptr;
        return o;
    };
    auto fail_modal = [&](const Outs& o) {
        fill_nan(o.lam, n, tid, nt);
        fill_nan(o.phi, (long long)A.n_modes * o.nu, tid, nt);
    };
    auto fail_all = [&](int code) {                  // numerical failure: NaN outputs, per-structure info
        const Outs o = outs();
        if (tid == 0) {
            *o.info = code;
            if (o.sweeps) *o.sweeps = 0;
        }
        if (A.do_static) {
            fill_nan(o.u, (long long)A.C * o.nu, tid, nt);
            fill_nan(o.ef, (long long)A.C * P.nE * 6, tid, nt);
            fill_nan(o.re, (long long)A.C * o.nre, tid, nt);
        }
        if (A.do_modal) fail_modal(o);
    };
    auto static_stage = [&]() {
        double* s_f = smem + L.U;
        double* s_ufull = s_f + A.C * n;
        double* s_ef = s_ufull + A.C * P.sumdof;
        const Outs o = outs();
        const double* g_fnod = A.f_nodal ? A.f_nodal + b * (in_w ? in_w : (long long)A.C * P.sumdof) : nullptr;
        const double* g_rloc = A.r_local ? A.r_local + b * (in_w ? in_w : (long long)A.C * P.nE * 6) : nullptr;
        const double* fn = g_fnod;
        const double* rl = g_rloc;
        if (PACKED && A.compact_in) {
            // compact loads: expand the nnz values of the load pattern into dense [C, sumdof] / [C, nE, 6] arrays in
            // shared memory (the modal stage is over: region U is free), then run the dense stage on them
            double* s_fd = s_ef + A.C * P.nE * 6;
            double* s_rd = s_fd + A.C * P.sumdof;
            const int nd = A.C * (P.sumdof + 6 * P.nE);
            for (int i = tid; i < nd; i += nt) s_fd[i] = 0.0;
            __syncthreads();
            for (int i = tid; i < A.nnz_f; i += nt) s_fd[A.f_idx[i]] = g_fnod[i];
            for (int i = tid; i < A.nnz_r; i += nt) s_rd[A.r_idx[i]] = g_rloc[i];
            __syncthreads();
            fn = s_fd;
            rl = A.nnz_r > 0 ? s_rd : nullptr;
        }
        stage_static(P, A.C, smem + L.props, smem + L.geo, smem + L.Kb, smem + L.invK, s_f, s_ufull, s_ef,
                     fn, rl, o.u, o.ef, o.re, tid, nt, ASSEMBLE, compact_out ? A.fixed_pos : nullptr, A.n_fixed);
    };
    int info = 0;
    {
        // Modal part first, while the band of K is still unfactored:  C = Lm^-1 K Lm^-T, pivoted Cholesky
        // C = Z Z^T and one-sided Jacobi on Z (in registers for n <= 64, in shared memory above); shapes
        // phi = Lm^-T u.  Then K = Lk Lk^T for the static part (and for the "K is not positive definite"
        // diagnosis, which wins over the modal codes).
        int sw = 0;
        // narrow band: every thread keeps its column of K in registers, so that K and M are factored at the same
        // time (warps 0 and 1) instead of one after the other
        const bool early_k = A.do_modal && P.hbw <= KCOL_HB && n <= nt;
        double kcol[2 * KCOL_HB + 1];
        if (early_k) {
            load_k_column(P, smem + L.Kb, kcol, tid);
            __syncthreads();
            factor(true, true);
            if (s_flag[0] != 0) {
                fail_all(s_flag[0]);
                return;
            }
        } else if (A.do_modal) {
            factor(false, true);
        }
        if (A.do_modal) {
            if (s_flag[1] == 0) {
                double* s_W = smem + L.U;
                stage_form_a(P, smem + L.Kb, smem + L.Mb, smem + L.invM, s_W, tid, nt, early_k ? &kcol : nullptr, ASSEMBLE);
                if constexpr (REGS == 1) {
                    sw = stage_jacobi_regs<RPL, LPP, true>(P, s_W, A.max_sweeps, tid);
                } else if constexpr (REGS == 2) {   // 64 < n <= 116: two column groups x LPP row chunks, NT = 64 LPP
                    sw = stage_jacobi_regs2<RPL, LPP, true>(P, s_W, A.max_sweeps, tid);
                } else {
                    sw = pivoted_cholesky_smem(P, s_W, smem + L.lam, s_order, tid, nt)
                             ? stage_jacobi_oe<RPL, LPP>(P, s_W, smem + L.lam, A.max_sweeps, tid) : INT_MIN;
                }
            }
        }
        // K = Lk Lk^T now (unless it was factored up front): the static part needs it, and the modal output stage
        // refines every eigenvalue as a Rayleigh quotient on the two factors
        if (!early_k) factor(true, false);
        if (A.do_modal && s_flag[1] == 0 && sw != INT_MIN) {
            const Outs o = outs();
            stage_modal_output<LPP>(P, smem + L.Mb, smem + L.invM, smem + L.U, smem + L.lam, s_order, A.n_modes,
                                    o.lam, o.phi, tid, nt, ASSEMBLE, compact_out, s_flag[0] == 0 ? smem + L.Kb : nullptr);
            __syncthreads();
        }
        const int infoK = s_flag[0], infoM = s_flag[1];
        if (infoK != 0) {
            fail_all(infoK);
            return;
        }
        if (infoM != 0 || sw == INT_MIN) {
            // the mass matrix is not positive definite (-1) or the pivoted Cholesky broke down (-2): only the MODAL
            // outputs are invalid -- K factored fine, so the static results of this structure are still delivered
            const Outs o = outs();
            if (tid == 0) {
                *o.info = infoM != 0 ? -1 : -2;
                if (o.sweeps) *o.sweeps = 0;
            }
            fail_modal(o);
            if (A.do_static) static_stage();
            return;
        }
        if (sw < 0) info = -2;
        if (tid == 0 && A.sweeps) *outs().sweeps = sw < 0 ? -sw : sw;
        if (A.do_static) static_stage();
    }
    if (tid == 0) *outs().info = info;
}

template <int RPL, int LPP, int NT, int MINB, bool ASSEMBLE, int REGS, bool PACKED = false>
int launch_solve(bfe2_plan* plan, const PlanDev& pd, const BatchArgs& A, cudaStream_t stream) {
    const SmemLayout L = make_layout(plan->nN, plan->nE, plan->n, plan->hbw, plan->npad, plan->LD, A.C);
    PlanDev P = pd;
    const size_t bytes = (size_t)L.total * sizeof(double);
    if (bytes > 227 * 1024) return fail(-3, "structure needs %zu bytes of shared memory (> 227 KB)", bytes);
    auto kern = k_solve<RPL, LPP, NT, MINB, ASSEMBLE, REGS, PACKED>;
    BFE2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    if (A.B > 0x7fffffffLL) return fail(-2, "batch too large for one launch");
    kern<<<(unsigned)A.B, NT, bytes, stream>>>(P, A);
    BFE2_CUDA(cudaGetLastError());
    return 0;
}

// ---- static path for structures whose band does not fit the shared memory of an SM: the same stages with the
//      per-structure arrays in a stream-ordered HBM workspace (L2-resident for a few structures at a time).
//      Slow per structure, but it removes the size limit of the static solve (and of the first-k modal route,
//      which applies K^-1 through bfe2_static_solve).
template <bool ASSEMBLE>
__global__ void __launch_bounds__(128)
k_static_global(PlanDev P, BatchArgs A, double* ws, long long ws_stride) {
    double* w = ws + (long long)blockIdx.x * ws_stride;
    const SmemLayout L = make_layout(P.nN, P.nE, P.n, P.hbw, 0, 0, A.C);
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
    const long long b = blockIdx.x;
    const int n = P.n;
    __shared__ int s_info;
    copy_in(w + L.coords, A.coords + b * 2 * P.nN, 2 * P.nN, tid, nt);
    copy_in(w + L.props, A.props + b * 4 * P.nE, 4 * P.nE, tid, nt);
    if (!ASSEMBLE) copy_in(w + L.Kb, A.Kb_in + b * P.nslots, P.nslots, tid, nt);
    __syncthreads();
    if (ASSEMBLE) {
        stage_elements_cf(P, w + L.coords, w + L.props, w + L.geo, w + L.U, w + L.U + K7_PITCH * P.nE, false, tid, nt);
        __syncthreads();
        stage_assemble_cf(P, w + L.U, w + L.U + K7_PITCH * P.nE, nullptr, w + L.Kb, w + L.Mb, false, tid, nt);
    } else {
        stage_geometry(P, w + L.coords, w + L.geo, tid, nt);
    }
    __syncthreads();
    if (warp == 0) {
        const int r = P.hbw <= 5 ? band_cholesky_group<32, 5>(w + L.Kb, w + L.invK, n, P.hbw, lane)
                    : P.hbw <= 8 ? band_cholesky_group<32, 8>(w + L.Kb, w + L.invK, n, P.hbw, lane)
                                 : band_cholesky_warp(w + L.Kb, w + L.invK, n, P.hbw, lane);
        if (lane == 0) s_info = r;
    }
    __syncthreads();
    if (s_info != 0) {
        if (tid == 0) A.info[b] = s_info;
        fill_nan(A.u + b * A.C * P.sumdof, (long long)A.C * P.sumdof, tid, nt);
        fill_nan(A.endforces ? A.endforces + b * A.C * P.nE * 6 : nullptr, (long long)A.C * P.nE * 6, tid, nt);
        fill_nan(A.reactions ? A.reactions + b * A.C * P.sumdof : nullptr, (long long)A.C * P.sumdof, tid, nt);
        return;
    }
    double* s_f = w + L.U;
    double* s_ufull = s_f + A.C * n;
    double* s_ef = s_ufull + A.C * P.sumdof;
    stage_static(P, A.C, w + L.props, w + L.geo, w + L.Kb, w + L.invK, s_f, s_ufull, s_ef,
                 A.f_nodal + b * A.C * P.sumdof, A.r_local ? A.r_local + b * A.C * P.nE * 6 : nullptr,
                 A.u + b * A.C * P.sumdof, A.endforces ? A.endforces + b * A.C * P.nE * 6 : nullptr,
                 A.reactions ? A.reactions + b * A.C * P.sumdof : nullptr, tid, nt, false);
    if (tid == 0) A.info[b] = 0;
}

template <bool ASSEMBLE>
int launch_static_global(bfe2_plan* plan, const PlanDev& pd, const BatchArgs& A, cudaStream_t stream) {
    PlanDev P = pd;
    P.npad = 0; P.LD = 0;
    const SmemLayout L = make_layout(plan->nN, plan->nE, plan->n, plan->hbw, 0, 0, A.C);
    const long long stride = ((long long)L.total + 1) & ~1LL;
    const long long chunk = std::max<long long>(1, std::min<long long>(A.B, (256LL << 20) / (stride * 8)));
    double* ws = nullptr;
    BFE2_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&ws), (size_t)(chunk * stride * 8), stream));
    const long long C = A.C, sd = plan->sumdof, nE = plan->nE, nN = plan->nN;
    for (long long b0 = 0; b0 < A.B; b0 += chunk) {
        BatchArgs Ac = A;
        Ac.B = std::min(chunk, A.B - b0);
        Ac.coords = A.coords + b0 * 2 * nN;
        Ac.props = A.props + b0 * 4 * nE;
        Ac.f_nodal = A.f_nodal + b0 * C * sd;
        if (A.r_local) Ac.r_local = A.r_local + b0 * C * nE * 6;
        if (A.Kb_in) Ac.Kb_in = A.Kb_in + b0 * plan->nslots;
        Ac.u = A.u + b0 * C * sd;
        if (A.endforces) Ac.endforces = A.endforces + b0 * C * nE * 6;
        if (A.reactions) Ac.reactions = A.reactions + b0 * C * sd;
        Ac.info = A.info + b0;
        k_static_global<ASSEMBLE><<<(unsigned)Ac.B, 128, 0, stream>>>(P, Ac, ws, stride);
    }
    const cudaError_t e = cudaGetLastError();
    cudaFreeAsync(ws, stream);
    if (e != cudaSuccess) return fail(-100, "k_static_global launch failed: %s", cudaGetErrorString(e));
    return 0;
}

template <bool ASSEMBLE, bool PACKED = false>
int dispatch_solve(bfe2_plan* plan, const PlanDev& pd, const BatchArgs& A, cudaStream_t stream) {
    // n <= 64: register-resident, preconditioned odd-even Jacobi (two warps per structure);
    // 64 < n <= 116: shared-memory round-robin variant (four lanes per column pair).
    if (plan->lpp == 2) {
        if (plan->rpl == 5) return launch_solve<5, 2, 64, 8, ASSEMBLE, 1, PACKED>(plan, pd, A, stream);
        if (plan->rpl == 15) return launch_solve<15, 2, 64, 8, ASSEMBLE, 1, PACKED>(plan, pd, A, stream);
        // __launch_bounds__(64, 5): same 168 registers and 6 resident blocks as (64, 6), but ptxas schedules the
        // software-pipelined step better (+4.7 % measured); (64, 4) takes 216 registers, 4 blocks and is slower.
        if (plan->rpl == 29) return launch_solve<29, 2, 64, 5, ASSEMBLE, 1, PACKED>(plan, pd, A, stream);
        if (plan->rpl == 32) return launch_solve<32, 2, 64, 5, ASSEMBLE, 1, PACKED>(plan, pd, A, stream);
    }
    if constexpr (ASSEMBLE) {
        // fused path, 100 < n <= 116: registers, two column groups x four row chunks (256 threads): +20 % over the
        // shared-memory exchange variant there; below, with fewer rows per lane, the exchange variant's three resident
        // blocks win (profiles/r2_sizes.txt).  BFE2_MIDSIZE_SMEM=1 forces the exchange variant for comparison.
        static const bool smem_variant = getenv("BFE2_MIDSIZE_SMEM") != nullptr;
        if (!smem_variant && plan->lpp == 4 && plan->rpl == 29)
            return launch_solve<29, 4, 256, 1, ASSEMBLE, 2, PACKED>(plan, pd, A, stream);
    }
    if (plan->rpl == 18 && plan->lpp == 4) return launch_solve<18, 4, 160, 3, ASSEMBLE, 0, PACKED>(plan, pd, A, stream);
    if (plan->rpl == 22 && plan->lpp == 4) return launch_solve<22, 4, 192, 2, ASSEMBLE, 0, PACKED>(plan, pd, A, stream);
    if (plan->rpl == 25 && plan->lpp == 4) return launch_solve<25, 4, 224, 2, ASSEMBLE, 0, PACKED>(plan, pd, A, stream);
    if (plan->rpl == 29 && plan->lpp == 4) return launch_solve<29, 4, 256, 1, ASSEMBLE, 0, PACKED>(plan, pd, A, stream);
    return fail(-3, "n_free = %d is beyond the shared-memory Jacobi path (max %d)", plan->n, BFE2_MAX_JACOBI_N);
}

}  // namespace

// ---- two-ended modal engine (bfe2_te.cuh; compiled in its own translation unit, bfe2_te_tu.cu) ---------------
// bfe2_te_launch: plan / PlanDev / BatchArgs pass as untyped pointers because BatchArgs lives in this file's anonymous
// namespace (every unit compiles the same text).  Returns 0 (launched), 1 (not eligible: nothing launched), < 0.
int bfe2_te_launch(bfe2_plan* plan, const void* plan_dev, const void* batch_args, void* stream, int packed);
extern "C" int bfe2_get_modal_engine(void);

namespace {
[[maybe_unused]] inline bool te_wanted(const bfe2_plan* plan, const BatchArgs& A) {
    return A.do_modal && plan->nE > 0 && plan->n >= 3 && plan->n <= 64 && A.n_modes <= 16 && plan->hbw < plan->n &&
           bfe2_get_modal_engine() == BFE2_ENGINE_TWO_ENDED;
}
}  // namespace

#ifdef BFE2_TU_TE
#include "bfe2_te.cuh"
namespace {

struct TeLayout {
    int coords, props, mass, geo, Kb, Mb, invK, invM, te, total, LD;
};
__host__ __device__ inline TeLayout make_te_layout(int nN, int nE, int n, int hbw, int C, int n_modes) {
    TeLayout L;
    int o = 0;
    L.coords = o; o += 2 * nN;
    L.props = o;  o += 4 * nE;
    L.mass = o;   o += 2 * nN;
    L.geo = o;    o += 3 * nE;
    L.Kb = o;     o += (hbw + 1) * n;
    L.Mb = o;     o += (hbw + 1) * n;
    L.invK = o;   o += n;
    L.invM = o;   o += n;
    o = (o + 1) & ~1;
    L.te = o;
    L.LD = n | 1;
    // region `te` is a union: beam numbers (elements / assembly), the modal block, the static stage
    o += imax3((K7_PITCH + M12_PITCH) * nE, te::smem_layout(n, L.LD, n_modes).total, C * (n + 6 * nN + 12 * nE));
    L.total = o;
    return L;
}

// fused K1 -> K4 with the two-ended modal engine: one block of 64 threads per structure
template <int N, bool PACKED>
__global__ void __launch_bounds__(64, 5)
k_solve_te(PlanDev P, BatchArgs A) {
    extern __shared__ __align__(16) double smem[];
    const TeLayout L = make_te_layout(P.nN, P.nE, P.n, P.hbw, A.C, A.n_modes);
    const int tid = threadIdx.x, nt = blockDim.x;
    const long long b = blockIdx.x;
    const int n = P.n;
    const long long in_w = PACKED ? A.w_in : 0, out_w = PACKED ? A.w_out : 0;
    const bool compact_out = PACKED && A.compact_out;
    const int nu = compact_out ? n : P.sumdof, nre = compact_out ? A.n_fixed : P.sumdof;
    double* g_u = A.u ? A.u + b * (out_w ? out_w : (long long)A.C * nu) : nullptr;
    double* g_ef = A.endforces ? A.endforces + b * (out_w ? out_w : (long long)A.C * P.nE * 6) : nullptr;
    double* g_re = A.reactions ? A.reactions + b * (out_w ? out_w : (long long)A.C * nre) : nullptr;
    double* g_lam = A.lam + b * (out_w ? out_w : (long long)n);
    double* g_phi = A.phi ? A.phi + b * (out_w ? out_w : (long long)A.n_modes * nu) : nullptr;
    int32_t* g_info = A.info + b * (out_w ? 2 * out_w : 1);
    int32_t* g_sweeps = A.sweeps ? A.sweeps + b * (out_w ? 2 * out_w : 1) : nullptr;

    copy_in(smem + L.coords, A.coords + b * (in_w ? in_w : 2LL * P.nN), 2 * P.nN, tid, nt);
    copy_in(smem + L.props, A.props + b * (in_w ? in_w : 4LL * P.nE), 4 * P.nE, tid, nt);
    if (A.nodal_mass != nullptr) copy_in(smem + L.mass, A.nodal_mass + b * (in_w ? in_w : 2LL * P.nN), 2 * P.nN, tid, nt);
    __syncthreads();
    double* base = smem + L.te;
    stage_elements_cf(P, smem + L.coords, smem + L.props, smem + L.geo, base, base + K7_PITCH * P.nE, true, tid, nt);
    __syncthreads();
    stage_assemble_cf(P, base, base + K7_PITCH * P.nE, A.nodal_mass ? smem + L.mass : nullptr, smem + L.Kb, smem + L.Mb,
                      true, tid, nt);
    __syncthreads();
    const te::Smem S = te::smem_layout(n, L.LD, A.n_modes);
    TeCtx ctx{tid, base + S.red, 0};
    int info_k = 0;
    const int rc = te::modal_block<N>(ctx, n, P.hbw, smem + L.Kb, smem + L.Mb, smem + L.invK, smem + L.invM, base, L.LD,
                                      A.n_modes, &info_k);
    if (rc > 0 || rc == BFE2_INFO_RETRY) {
        // K not positive definite: every output is NaN.  Engine declined: the Jacobi launch that follows redoes the
        // whole structure; leave NaNs so that a missing repair cannot go unnoticed.
        if (tid == 0) {
            *g_info = rc;
            if (g_sweeps) *g_sweeps = 0;
        }
        if (A.do_static) {
            fill_nan(g_u, (long long)A.C * nu, tid, nt);
            fill_nan(g_ef, (long long)A.C * P.nE * 6, tid, nt);
            fill_nan(g_re, (long long)A.C * nre, tid, nt);
        }
        fill_nan(g_lam, n, tid, nt);
        fill_nan(g_phi, (long long)A.n_modes * nu, tid, nt);
        return;
    }
    if (rc == -1) {                                   // mass matrix not positive definite: modal outputs NaN, static valid
        fill_nan(g_lam, n, tid, nt);
        fill_nan(g_phi, (long long)A.n_modes * nu, tid, nt);
    } else {
        for (int j = tid; j < n; j += nt) g_lam[j] = base[S.lam + j];
        const double* Z = base + S.Z;
        if (g_phi != nullptr) {
            if (compact_out) {
                for (int t = tid; t < A.n_modes * n; t += nt) g_phi[t] = Z[(t / n) * te::NT + t % n];
            } else {
                for (int t = tid; t < A.n_modes * P.sumdof; t += nt) {
                    const int m = t / P.sumdof, j = P.free_of_pos[t % P.sumdof];
                    g_phi[t] = (j >= 0) ? Z[m * te::NT + j] : 0.0;
                }
            }
        }
    }
    if (tid == 0) {
        *g_info = rc;                                 // 0 or -1
        if (g_sweeps) *g_sweeps = 0;                  // no Jacobi sweeps in this engine
    }
    if (!A.do_static) return;
    __syncthreads();                                  // region `te` becomes the static stage's scratch
    double* s_f = base;
    double* s_ufull = s_f + A.C * n;
    double* s_ef = s_ufull + A.C * P.sumdof;
    const double* fn = A.f_nodal + b * (in_w ? in_w : (long long)A.C * P.sumdof);
    const double* rl = A.r_local ? A.r_local + b * (in_w ? in_w : (long long)A.C * P.nE * 6) : nullptr;
    if (PACKED && A.compact_in) {
        double* s_fd = s_ef + A.C * P.nE * 6;
        double* s_rd = s_fd + A.C * P.sumdof;
        const int nd = A.C * (P.sumdof + 6 * P.nE);
        for (int i = tid; i < nd; i += nt) s_fd[i] = 0.0;
        __syncthreads();
        for (int i = tid; i < A.nnz_f; i += nt) s_fd[A.f_idx[i]] = fn[i];
        for (int i = tid; i < A.nnz_r; i += nt) s_rd[A.r_idx[i]] = rl[i];
        __syncthreads();
        fn = s_fd;
        rl = A.nnz_r > 0 ? s_rd : nullptr;
    }
    stage_static(P, A.C, smem + L.props, smem + L.geo, smem + L.Kb, smem + L.invK, s_f, s_ufull, s_ef, fn, rl, g_u, g_ef,
                 g_re, tid, nt, true, compact_out ? A.fixed_pos : nullptr, A.n_fixed);
}

template <int N, bool PACKED>
int launch_te(bfe2_plan* plan, const PlanDev& pd, const BatchArgs& A, cudaStream_t stream) {
    const TeLayout L = make_te_layout(plan->nN, plan->nE, plan->n, plan->hbw, A.C, A.n_modes);
    const size_t bytes = (size_t)L.total * sizeof(double);
    if (bytes > 227 * 1024) return 1;
    auto kern = k_solve_te<N, PACKED>;
    BFE2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    if (A.B > 0x7fffffffLL) return fail(-2, "batch too large for one launch");
    kern<<<(unsigned)A.B, 64, bytes, stream>>>(pd, A);
    BFE2_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace

int bfe2_te_launch(bfe2_plan* plan, const void* plan_dev, const void* batch_args, void* stream, int packed) {
    const PlanDev& pd = *static_cast<const PlanDev*>(plan_dev);
    const BatchArgs& A = *static_cast<const BatchArgs*>(batch_args);
    if (!te_wanted(plan, A)) return 1;
    cudaStream_t st = (cudaStream_t)stream;
    if (plan->n <= 32) return packed ? launch_te<32, true>(plan, pd, A, st) : launch_te<32, false>(plan, pd, A, st);
    if (plan->n <= 58) return packed ? launch_te<58, true>(plan, pd, A, st) : launch_te<58, false>(plan, pd, A, st);
    return packed ? launch_te<64, true>(plan, pd, A, st) : launch_te<64, false>(plan, pd, A, st);
}
#endif  // BFE2_TU_TE

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" {

#if !defined(BFE2_TU_PACKED) && !defined(BFE2_TU_TE)
int bfe2_version(void) { return BFE2_VERSION; }

const char* bfe2_last_error(void) { return g_err.c_str(); }

int bfe2_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int bfe2_plan_create(bfe2_plan** out, int32_t n_nodes, int32_t n_elem, const int32_t* conn,
                     int32_t n_fixed, const int32_t* fixed_pos, const uint8_t* lumped) {
    return bfe2_plan_create_ex(out, n_nodes, n_elem, conn, n_fixed, fixed_pos, lumped, 0);
}

int bfe2_plan_create_ex(bfe2_plan** out, int32_t n_nodes, int32_t n_elem, const int32_t* conn,
                        int32_t n_fixed, const int32_t* fixed_pos, const uint8_t* lumped, int32_t flags) {
    if (!out) return fail(-1, "out is NULL");
    *out = nullptr;
    if (n_nodes < 2 || n_elem < 1 || !conn) return fail(-1, "need at least 2 nodes and 1 beam");
    if (n_fixed < 0 || (n_fixed > 0 && !fixed_pos)) return fail(-1, "bad fixed_pos");
    bfe2_plan* p = new bfe2_plan();
    p->nN = n_nodes; p->nE = n_elem; p->sumdof = 3 * n_nodes;
    p->conn.assign(conn, conn + 2 * (size_t)n_elem);
    std::vector<char> seen(n_nodes, 0);
    for (int e = 0; e < n_elem; ++e) {
        const int i = conn[2 * e], j = conn[2 * e + 1];
        if (i < 0 || j < 0 || i >= n_nodes || j >= n_nodes || i == j) {
            delete p;
            return fail(-1, "beam %d has invalid node indices (%d, %d)", e, i, j);
        }
        seen[i] = seen[j] = 1;
    }
    for (int k = 0; k < n_nodes; ++k)
        if (!seen[k]) {   // Structure.node_numbering_is_ok (Structure.py:334-341): IDs must be 1..N, all used
            delete p;
            return fail(-1, "node %d is not attached to any beam (node numbering must be contiguous)", k + 1);
        }
    p->fixed.assign(fixed_pos, fixed_pos + n_fixed);
    std::sort(p->fixed.begin(), p->fixed.end());                       // Structure.py:180
    for (int k = 0; k < n_fixed; ++k) {
        if (p->fixed[k] < 0 || p->fixed[k] >= p->sumdof || (k > 0 && p->fixed[k] == p->fixed[k - 1])) {
            const int bad = p->fixed[k];
            delete p;
            return fail(-1, "fixed position %d out of range or duplicated", bad);
        }
    }
    p->free_of_pos.assign(p->sumdof, 0);
    for (int f : p->fixed) p->free_of_pos[f] = -1;
    p->n = 0;
    for (int pos = 0; pos < p->sumdof; ++pos)
        if (p->free_of_pos[pos] == 0) { p->free_of_pos[pos] = p->n++; p->pos_of_free.push_back(pos); }
    // n == 0 (every DOF supported, e.g. the clamped-clamped element of Tests/test_internal_beamloads.py)
    // is legal for the static path: u = 0, end forces = -r_local.  The modal path is refused.
    // half bandwidth in free numbering
    p->hbw = half_bandwidth(p, conn);
    if ((flags & BFE2_PLAN_REORDER) && p->n > 0) {
        // internal bandwidth-reducing renumbering of the FREE DOFs (reverse Cuthill-McKee on the node graph, the
        // DOFs of a node stay together).  Purely internal: inputs and outputs keep the reference's positions
        // (Structure.py:379-393) because every kernel goes through free_of_pos / pos_of_free.
        std::vector<int32_t> fop(p->sumdof, -1), pof;
        for (int f : p->fixed) fop[f] = -2;
        for (int node : rcm_node_order(n_nodes, n_elem, conn))
            for (int k = 0; k < 3; ++k)
                if (fop[3 * node + k] != -2) { fop[3 * node + k] = (int32_t)pof.size(); pof.push_back(3 * node + k); }
        for (int f : p->fixed) fop[f] = -1;
        bfe2_plan trial;
        trial.nE = p->nE;
        trial.free_of_pos = fop;
        const int hb2 = half_bandwidth(&trial, conn);
        if (hb2 < p->hbw) { p->free_of_pos = fop; p->pos_of_free = pof; p->hbw = hb2; }
    }
    p->nslots = (p->hbw + 1) * p->n;
    // scatter plan: slot (d, j) <- e*36 + r*6 + c for every lower-triangle contribution, beam order
    std::vector<std::vector<int32_t>> lists(p->nslots);
    for (int e = 0; e < n_elem; ++e)
        for (int r = 0; r < 6; ++r)
            for (int c = 0; c < 6; ++c) {
                const int fr = p->free_of_pos[3 * conn[2 * e + r / 3] + r % 3];   // helpers.py:28-31
                const int fc = p->free_of_pos[3 * conn[2 * e + c / 3] + c % 3];
                if (fr < 0 || fc < 0 || fr < fc) continue;
                lists[(fr - fc) * p->n + fc].push_back(e * 36 + r * 6 + c);
            }
    p->slot_ptr.assign(p->nslots + 1, 0);
    for (int s = 0; s < p->nslots; ++s) {
        p->slot_ptr[s + 1] = p->slot_ptr[s] + (int)lists[s].size();
        p->contrib.insert(p->contrib.end(), lists[s].begin(), lists[s].end());
    }
    p->ncontrib = (int)p->contrib.size();
    // which of the 7 (K) / 12 (M) closed-form numbers of a beam entry (r, c) of its global 6 x 6 matrix equals
    static const signed char h_ke7[36] = {+1, +2, -4, -1, -2, -4, +2, +3, +5, -2, -3, +5, -4, +5, +6, +4, -5, +7,
                                          -1, -2, +4, +1, +2, +4, -2, -3, -5, +2, +3, -5, -4, +5, +7, +4, -5, +6};
    static const signed char h_me12[36] = {+1, +2, -4, +7, +8, +10, +2, +3, +5, +8, +9, -11, -4, +5, +6, -10, +11, +12,
                                           +7, +8, -10, +1, +2, +4, +8, +9, +11, +2, +3, -5, +10, -11, +12, +4, -5, +6};
    p->contrib_k.resize(p->ncontrib);
    p->contrib_m.resize(p->ncontrib);
    for (int i = 0; i < p->ncontrib; ++i) {
        const int idx = p->contrib[i], e = idx / 36;
        const int ck = h_ke7[idx % 36], cm = h_me12[idx % 36];
        p->contrib_k[i] = (e * bfe2::K7_PITCH + std::abs(ck) - 1) | (ck < 0 ? (int32_t)0x80000000 : 0);
        p->contrib_m[i] = (e * bfe2::M12_PITCH + std::abs(cm) - 1) | (cm < 0 ? (int32_t)0x80000000 : 0);
    }
    // node -> incident beam ends, beam order
    std::vector<std::vector<int32_t>> inc(n_nodes);
    for (int e = 0; e < n_elem; ++e) {
        inc[conn[2 * e]].push_back(2 * e);
        inc[conn[2 * e + 1]].push_back(2 * e + 1);
    }
    p->node_ptr.assign(n_nodes + 1, 0);
    for (int k = 0; k < n_nodes; ++k) {
        p->node_ptr[k + 1] = p->node_ptr[k] + (int)inc[k].size();
        p->node_inc.insert(p->node_inc.end(), inc[k].begin(), inc[k].end());
    }
    p->lumped.assign(n_elem, 0);
    if (lumped) p->lumped.assign(lumped, lumped + n_elem);
    p->npad = p->n + (p->n & 1);
    if (p->n > 0 && jacobi_config(p->n, &p->rpl, &p->lpp)) {
        p->LD = p->rpl * p->lpp;
        if (p->n <= 64) p->LD |= 1;               // register path: odd slot pitch, thread-per-slot access is conflict-free
        p->nthreads = std::max(64, ((p->npad / 2) * p->lpp + 31) / 32 * 32);
    } else {
        p->rpl = p->lpp = 0;          // static / assembly still work; modal is refused
        p->LD = 0;
        p->nthreads = 64;
    }
    *out = p;
    return 0;
}

void bfe2_plan_destroy(bfe2_plan* plan) {
    if (!plan) return;
    int cur = 0;
    const bool have_cur = cudaGetDevice(&cur) == cudaSuccess;
    for (auto& kv : plan->copies) {
        if (cudaSetDevice(kv.first) == cudaSuccess) cudaFree(kv.second.first);
    }
    if (have_cur) cudaSetDevice(cur);
    delete plan;
}

int bfe2_plan_dims(const bfe2_plan* p, int32_t* n_nodes, int32_t* n_elem, int32_t* sumdof, int32_t* n_free,
                   int32_t* hbw, int32_t* n_contrib) {
    if (!p) return fail(-1, "plan is NULL");
    if (n_nodes) *n_nodes = p->nN;
    if (n_elem) *n_elem = p->nE;
    if (sumdof) *sumdof = p->sumdof;
    if (n_free) *n_free = p->n;
    if (hbw) *hbw = p->hbw;
    if (n_contrib) *n_contrib = p->ncontrib;
    return 0;
}

int bfe2_plan_get_array(const bfe2_plan* p, int which, int32_t* out, int64_t capacity) {
    if (!p || !out) return fail(-1, "NULL argument");
    const std::vector<int32_t>* v = nullptr;
    switch (which) {
        case BFE2_ARR_CONN: v = &p->conn; break;
        case BFE2_ARR_FREE_OF_POS: v = &p->free_of_pos; break;
        case BFE2_ARR_POS_OF_FREE: v = &p->pos_of_free; break;
        case BFE2_ARR_FIXED: v = &p->fixed; break;
        case BFE2_ARR_SLOT_PTR: v = &p->slot_ptr; break;
        case BFE2_ARR_CONTRIB: v = &p->contrib; break;
        case BFE2_ARR_NODE_PTR: v = &p->node_ptr; break;
        case BFE2_ARR_NODE_INC: v = &p->node_inc; break;
        default: return fail(-1, "unknown plan array %d", which);
    }
    if ((int64_t)v->size() > capacity) return fail(-1, "capacity %lld < %zu", (long long)capacity, v->size());
    if (!v->empty()) memcpy(out, v->data(), v->size() * sizeof(int32_t));
    return (int)v->size();
}

size_t bfe2_plan_smem_bytes(const bfe2_plan* p, int32_t C) {
    if (!p) return 0;
    const SmemLayout L = make_layout(p->nN, p->nE, p->n, p->hbw, p->npad, p->LD, C);
    return (size_t)L.total * sizeof(double);
}

size_t bfe2_workspace_bytes(const bfe2_plan*, int64_t) { return 0; }

int bfe2_element_km(int64_t nE, const double* xi, const double* yi, const double* xj, const double* yj,
                    const double* E, const double* A, const double* I, const double* rho,
                    const uint8_t* lumped, double* Kg, double* Mg, void* stream) {
    if (nE == 0) return 0;                        // empty arrays may come with NULL pointers
    if (nE < 0 || !xi || !yi || !xj || !yj || !E || !A || !I || !rho || !Kg || !Mg)
        return fail(-1, "bfe2_element_km: NULL argument");
    int dev = 0, sms = 0;
    BFE2_CUDA(cudaGetDevice(&dev));
    BFE2_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const long long tiles = (nE + 31) / 32;
    const long long want = (tiles + K1_WARPS - 1) / K1_WARPS;
    const int grid = (int)std::min<long long>(want, (long long)sms * 3 * 8);   // 3 resident blocks/SM, 8 waves
    k_element_km<<<grid, K1_WARPS * 32, 0, (cudaStream_t)stream>>>(nE, xi, yi, xj, yj, E, A, I, rho, lumped, Kg, Mg);
    BFE2_CUDA(cudaGetLastError());
    return 0;
}

int bfe2_assemble_banded(bfe2_plan* plan, int64_t B, const double* coords, const double* props,
                         const double* nodal_mass, double* Kb, double* Mb, void* stream) {
    if (plan && B == 0) return 0;
    if (!plan || B < 0 || !coords || !props || !Kb) return fail(-1, "bfe2_assemble_banded: bad argument");
    PlanDev pd;
    if (int r = plan_upload(plan, &pd)) return r;
    BatchArgs A{};
    A.B = B; A.coords = coords; A.props = props; A.nodal_mass = nodal_mass; A.Kb_out = Kb; A.Mb_out = Mb;
    const size_t per_warp = (size_t)((K7_PITCH + M12_PITCH) * plan->nE + 2 * plan->nN) * sizeof(double);
    size_t cache = (size_t)(2 * plan->nE + plan->n + plan->nslots + 1 + 2 * plan->ncontrib) * sizeof(int) + 16;
    const int cached = cache + per_warp <= 100 * 1024;     // a larger plan stays in global memory / L2
    if (!cached) cache = 16;
    int nw = K2_WARPS;                                 // fewer warps per block when a structure's slice is large
    while (nw > 1 && nw * per_warp + cache > 110 * 1024) --nw;
    const size_t bytes = nw * per_warp + cache;
    if (bytes > 227 * 1024) return fail(-3, "structure needs %zu bytes of shared memory", bytes);
    BFE2_CUDA(cudaFuncSetAttribute(k_assemble, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    int dev = 0, sms = 0;
    BFE2_CUDA(cudaGetDevice(&dev));
    BFE2_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int resident = std::max<int>(1, (int)((227 * 1024) / (bytes + 1024)));
    const long long want = (B + nw - 1) / nw;
    const int grid = (int)std::min<long long>(want, (long long)sms * resident);
    k_assemble<<<grid, nw * 32, bytes, (cudaStream_t)stream>>>(pd, A, plan->ncontrib, cached);
    BFE2_CUDA(cudaGetLastError());
    return 0;
}

int bfe2_static_solve(bfe2_plan* plan, int64_t B, int32_t C, const double* Kb, const double* coords,
                      const double* props, const double* f_nodal, const double* r_local, double* u,
                      double* endforces, double* reactions, int32_t* info, void* stream) {
    if (plan && B == 0) return 0;
    if (!plan || B < 0 || C < 1 || !Kb || !coords || !props || !f_nodal || !u || !info)
        return fail(-1, "bfe2_static_solve: bad argument");
    PlanDev pd;
    if (int r = plan_upload(plan, &pd)) return r;
    BatchArgs A{};
    A.B = B; A.C = C; A.do_static = 1; A.do_modal = 0; A.max_sweeps = BFE2_MAX_SWEEPS;
    A.coords = coords; A.props = props; A.f_nodal = f_nodal; A.r_local = r_local; A.Kb_in = Kb;
    A.u = u; A.endforces = endforces; A.reactions = reactions; A.info = info;
    if (const int r = dispatch_static_group<false>(plan, pd, A, (cudaStream_t)stream); r <= 0) return r;
    const size_t per_warp = (size_t)(((2 * plan->nN + 4 * plan->nE + 3 * plan->nE + plan->nslots + plan->n +
                                       C * (plan->n + 3 * plan->nN + 6 * plan->nE)) + 1) & ~1) * sizeof(double);
    const size_t cache = (size_t)(4 * plan->nE + plan->sumdof + plan->n + plan->nN + 1) * sizeof(int) + 16;
    int nw = K3_WARPS;
    while (nw > 1 && nw * per_warp + cache > 110 * 1024) --nw;
    const size_t bytes = nw * per_warp + cache;
    if (bytes > 227 * 1024) return launch_static_global<false>(plan, pd, A, (cudaStream_t)stream);
    BFE2_CUDA(cudaFuncSetAttribute(k_static_warp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    int dev = 0, sms = 0;
    BFE2_CUDA(cudaGetDevice(&dev));
    BFE2_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int resident = std::max<int>(1, (int)((227 * 1024) / (bytes + 1024)));
    const long long want = (B + nw - 1) / nw;
    const int grid = (int)std::min<long long>(want, (long long)sms * resident);
    k_static_warp<<<grid, nw * 32, bytes, (cudaStream_t)stream>>>(pd, A);
    BFE2_CUDA(cudaGetLastError());
    return 0;
}

int bfe2_modal_solve(bfe2_plan* plan, int64_t B, int32_t n_modes, const double* Kb, const double* Mb,
                     double* lam, double* phi, int32_t* info, int32_t* sweeps, void* stream) {
    if (plan && B == 0) return 0;
    if (!plan || B < 0 || !Kb || !Mb || !lam || !info) return fail(-1, "bfe2_modal_solve: bad argument");
    if (n_modes < 0 || n_modes > plan->n || (n_modes > 0 && !phi)) return fail(-1, "bad n_modes");
    if (plan->rpl == 0) return fail(-3, "n_free = %d is outside the Jacobi path (1..%d)", plan->n, BFE2_MAX_JACOBI_N);
    if (B == 0) return 0;
    PlanDev pd;
    if (int r = plan_upload(plan, &pd)) return r;
    BatchArgs A{};
    A.B = B; A.C = 0; A.do_static = 0; A.do_modal = 1; A.n_modes = n_modes; A.max_sweeps = BFE2_MAX_SWEEPS;
    A.Kb_in = Kb; A.Mb_in = Mb; A.lam = lam; A.phi = phi; A.info = info; A.sweeps = sweeps;
    return dispatch_solve<false>(plan, pd, A, (cudaStream_t)stream);
}

int bfe2_solve_fused(bfe2_plan* plan, int64_t B, int32_t C, int32_t do_modal, int32_t n_modes,
                     const double* coords, const double* props, const double* nodal_mass,
                     const double* f_nodal, const double* r_local, double* u, double* endforces,
                     double* reactions, double* lam, double* phi, int32_t* info, int32_t* sweeps,
                     void* stream) {
    if (plan && B == 0) return 0;
    if (!plan || B < 0 || C < 0 || !coords || !props || !info) return fail(-1, "bfe2_solve_fused: bad argument");
    if (C > 0 && (!f_nodal || !u)) return fail(-1, "static outputs requested without f_nodal / u");
    if (C == 0 && !do_modal) return fail(-1, "nothing to do");
    if (do_modal) {
        if (!lam) return fail(-1, "lam is NULL");
        if (n_modes < 0 || n_modes > plan->n || (n_modes > 0 && !phi)) return fail(-1, "bad n_modes");
        if (plan->rpl == 0)
            return fail(-3, "n_free = %d is outside the Jacobi path (1..%d)", plan->n, BFE2_MAX_JACOBI_N);
    }
    if (B == 0) return 0;
    PlanDev pd;
    if (int r = plan_upload(plan, &pd)) return r;
    BatchArgs A{};
    A.B = B; A.C = C; A.do_static = C > 0; A.do_modal = do_modal != 0; A.n_modes = do_modal ? n_modes : 0;
    A.max_sweeps = BFE2_MAX_SWEEPS;
    A.coords = coords; A.props = props; A.nodal_mass = nodal_mass; A.f_nodal = f_nodal; A.r_local = r_local;
    A.u = u; A.endforces = endforces; A.reactions = reactions; A.lam = lam; A.phi = phi;
    A.info = info; A.sweeps = sweeps;
    if (!do_modal) {
        {   // large batches of narrow-band structures: one lane per structure (bfe2_tile.cu)
            bfe2_tile_args T{B, C, coords, props, f_nodal, r_local, u, endforces, reactions, info};
            if (const int r = bfe2_tile_static_launch(pd, T, (cudaStream_t)stream); r <= 0) return r;
        }
        if (const int r = dispatch_static_group<true>(plan, pd, A, (cudaStream_t)stream); r <= 0) return r;
        PlanDev P = pd;
        P.npad = 0; P.LD = 0;
        const SmemLayout L = make_layout(plan->nN, plan->nE, plan->n, plan->hbw, 0, 0, C);
        const size_t bytes = (size_t)L.total * sizeof(double);
        if (bytes > 227 * 1024) return launch_static_global<true>(plan, pd, A, (cudaStream_t)stream);
        auto kern = k_solve<5, 2, 64, 8, true, 0>;
        BFE2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        if (B > 0x7fffffffLL) return fail(-2, "batch too large for one launch");
        kern<<<(unsigned)B, 64, bytes, (cudaStream_t)stream>>>(P, A);
        BFE2_CUDA(cudaGetLastError());
        return 0;
    }
    // two-ended engine selected (and n_free <= 64, few shapes wanted): it runs first; the Jacobi engine then only
    // repairs the structures that engine declined (blocks of the others leave at once)
    if (const int r = bfe2_te_launch(plan, &pd, &A, stream, 0); r < 0) return r;
    else if (r == 0) A.only_flagged = 1;
    return dispatch_solve<true>(plan, pd, A, (cudaStream_t)stream);
}

static std::atomic<int> g_engine{BFE2_ENGINE_JACOBI};
int bfe2_set_modal_engine(int32_t engine) {
    if (engine != BFE2_ENGINE_JACOBI && engine != BFE2_ENGINE_TWO_ENDED) return fail(-1, "unknown modal engine %d", engine);
    g_engine.store(engine);
    return 0;
}
int bfe2_get_modal_engine(void) { return g_engine.load(); }

#endif  // !BFE2_TU_PACKED

#ifdef BFE2_TU_PACKED
// ---- packed records + compact load / result layouts (end-to-end path from HOST memory) -----------------
// One record per structure and direction, so that a chunk of the batch moves with ONE copy each way, and only the
// numbers that carry information travel: the non-zero entries of the load tensors (a load pattern shared by the
// batch -- the reference's add_nodal_load / add_internal_loads touch a handful of positions, Structure.py:258-287),
// displacements and shapes on the free DOFs (the reference re-inserts the zero rows on the host, Solver.py:57-65),
// reactions at the supported DOFs (the only ones Results.py:129-143 reads).
struct bfe2_pattern {
    int C = 0, do_modal = 0, n_modes = 0, nnz_f = 0, nnz_r = 0, n_fixed = 0;
    int in_off[5] = {0, 0, 0, 0, 0}, out_off[7] = {0, 0, 0, 0, 0, 0, 0};
    int w_in = 0, w_out = 0;
    std::vector<int32_t> host;                        // f_idx | r_idx | fixed_pos
    std::mutex mu;
    std::map<int, int32_t*> copies;                   // device mirrors, one per device
    ~bfe2_pattern() {
        for (auto& kv : copies) cudaFree(kv.second);
    }
};

int bfe2_pattern_create(bfe2_pattern** out, bfe2_plan* plan, int32_t C, int32_t do_modal, int32_t n_modes,
                        int32_t nnz_f, const int32_t* f_index, int32_t nnz_r, const int32_t* r_index) {
    if (!out || !plan || C < 0 || nnz_f < 0 || nnz_r < 0 || (nnz_f > 0 && !f_index) || (nnz_r > 0 && !r_index))
        return fail(-1, "bfe2_pattern_create: bad argument");
    if (C == 0 && !do_modal) return fail(-1, "bfe2_pattern_create: nothing to do");
    if (do_modal && (n_modes < 0 || n_modes > plan->n)) return fail(-1, "bfe2_pattern_create: bad n_modes");
    if (do_modal && plan->rpl == 0)
        return fail(-3, "n_free = %d is outside the Jacobi path (1..%d)", plan->n, BFE2_MAX_JACOBI_N);
    for (int i = 0; i < nnz_f; ++i)
        if (f_index[i] < 0 || f_index[i] >= C * plan->sumdof) return fail(-1, "bfe2_pattern_create: f_index[%d] out of range", i);
    for (int i = 0; i < nnz_r; ++i)
        if (r_index[i] < 0 || r_index[i] >= C * plan->nE * 6) return fail(-1, "bfe2_pattern_create: r_index[%d] out of range", i);
    bfe2_pattern* q = new bfe2_pattern();
    q->C = C; q->do_modal = do_modal != 0; q->n_modes = do_modal ? n_modes : 0;
    q->nnz_f = nnz_f; q->nnz_r = nnz_r; q->n_fixed = (int)plan->fixed.size();
    q->host.insert(q->host.end(), f_index, f_index + nnz_f);
    q->host.insert(q->host.end(), r_index, r_index + nnz_r);
    q->host.insert(q->host.end(), plan->fixed.begin(), plan->fixed.end());
    // input record: coords | props | nodal_mass | f values | r values
    int o = 0;
    q->in_off[0] = o; o += 2 * plan->nN;
    q->in_off[1] = o; o += 4 * plan->nE;
    q->in_off[2] = o; o += 2 * plan->nN;
    q->in_off[3] = o; o += nnz_f;
    q->in_off[4] = o; o += nnz_r;
    q->w_in = o;
    // output record: u [C, n] | endforces [C, nE, 6] | reactions [C, n_fixed] | lam [n] | phi [n_modes, n] | info, sweeps
    o = 0;
    q->out_off[0] = o; o += C * plan->n;
    q->out_off[1] = o; o += C * plan->nE * 6;
    q->out_off[2] = o; o += C * q->n_fixed;
    q->out_off[3] = o; o += q->do_modal ? plan->n : 0;
    q->out_off[4] = o; o += q->n_modes * plan->n;
    q->out_off[5] = o; o += 1;                        // two int32: info, sweeps
    q->out_off[6] = o;
    q->w_out = o;
    *out = q;
    return 0;
}

void bfe2_pattern_destroy(bfe2_pattern* q) { delete q; }

int bfe2_pattern_layout(const bfe2_pattern* q, int32_t* in_width, int32_t* out_width, int32_t* in_offsets,
                        int32_t* out_offsets) {
    if (!q) return fail(-1, "pattern is NULL");
    if (in_width) *in_width = q->w_in;
    if (out_width) *out_width = q->w_out;
    if (in_offsets) memcpy(in_offsets, q->in_off, sizeof(q->in_off));
    if (out_offsets) memcpy(out_offsets, q->out_off, 6 * sizeof(int));
    return 0;
}

int bfe2_solve_fused_packed(bfe2_plan* plan, bfe2_pattern* q, int64_t B, const double* in_rec, double* out_rec,
                            void* stream) {
    if (!plan || !q || B < 0 || (B > 0 && (!in_rec || !out_rec))) return fail(-1, "bfe2_solve_fused_packed: bad argument");
    if (B == 0) return 0;
    if (q->n_fixed != (int)plan->fixed.size()) return fail(-1, "pattern was created for another plan");
    PlanDev pd;
    if (int r = bfe2_plan_device_view(plan, &pd)) return r;
    int dev = 0;
    BFE2_CUDA(cudaGetDevice(&dev));
    const int32_t* dpat = nullptr;
    {
        std::lock_guard<std::mutex> lock(q->mu);
        auto it = q->copies.find(dev);
        if (it == q->copies.end()) {
            int32_t* d = nullptr;
            const size_t bytes = std::max<size_t>(4, q->host.size() * sizeof(int32_t));
            BFE2_CUDA(cudaMalloc(&d, bytes));
            if (!q->host.empty() && cudaMemcpy(d, q->host.data(), q->host.size() * sizeof(int32_t), cudaMemcpyHostToDevice) != cudaSuccess) {
                cudaFree(d);
                return fail(-100, "pattern upload failed: %s", cudaGetErrorString(cudaGetLastError()));
            }
            it = q->copies.emplace(dev, d).first;
        }
        dpat = it->second;
    }
    BatchArgs A{};
    A.B = B; A.C = q->C; A.do_static = q->C > 0; A.do_modal = q->do_modal; A.n_modes = q->n_modes;
    A.max_sweeps = BFE2_MAX_SWEEPS;
    A.w_in = q->w_in; A.w_out = q->w_out; A.compact_in = 1; A.compact_out = 1;
    A.nnz_f = q->nnz_f; A.nnz_r = q->nnz_r; A.n_fixed = q->n_fixed;
    A.f_idx = dpat; A.r_idx = dpat + q->nnz_f; A.fixed_pos = dpat + q->nnz_f + q->nnz_r;
    A.coords = in_rec + q->in_off[0]; A.props = in_rec + q->in_off[1]; A.nodal_mass = in_rec + q->in_off[2];
    if (q->C > 0) {
        A.f_nodal = in_rec + q->in_off[3]; A.r_local = in_rec + q->in_off[4];
        A.u = out_rec + q->out_off[0]; A.endforces = out_rec + q->out_off[1]; A.reactions = out_rec + q->out_off[2];
    }
    if (q->do_modal) {
        A.lam = out_rec + q->out_off[3];
        A.phi = q->n_modes > 0 ? out_rec + q->out_off[4] : nullptr;
    }
    A.info = reinterpret_cast<int32_t*>(out_rec + q->out_off[5]);
    A.sweeps = A.info + 1;
    if (!q->do_modal) {                               // static only: the block-per-structure kernel without the modal stage
        PlanDev P = pd;
        P.npad = 0; P.LD = 0;
        const SmemLayout L = make_layout(plan->nN, plan->nE, plan->n, plan->hbw, 0, 0, q->C);
        const size_t bytes = (size_t)L.total * sizeof(double);
        if (bytes > 227 * 1024) return fail(-3, "structure needs %zu bytes of shared memory (> 227 KB)", bytes);
        auto kern = k_solve<5, 2, 64, 8, true, 0, true>;
        BFE2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        if (B > 0x7fffffffLL) return fail(-2, "batch too large for one launch");
        kern<<<(unsigned)B, 64, bytes, (cudaStream_t)stream>>>(P, A);
        BFE2_CUDA(cudaGetLastError());
        return 0;
    }
    if (const int r = bfe2_te_launch(plan, &pd, &A, stream, 1); r < 0) return r;
    else if (r == 0) A.only_flagged = 1;
    return dispatch_solve<true, true>(plan, pd, A, (cudaStream_t)stream);
}

#endif  // BFE2_TU_PACKED

#if !defined(BFE2_TU_PACKED) && !defined(BFE2_TU_TE)
// ---- small dense generalized symmetric eigenproblem A q = theta B q (Rayleigh-Ritz step of the subspace
//      iteration): the same Cholesky-reduction + Jacobi kernel on a synthetic "plan" with identity maps ---
int bfe2_plan_create_dense(bfe2_plan** out, int32_t n) {
    if (!out || n < 1 || n > BFE2_MAX_JACOBI_N) return fail(-1, "bfe2_plan_create_dense: n must be in 1..%d", BFE2_MAX_JACOBI_N);
    bfe2_plan* p = new bfe2_plan();
    p->nN = 0; p->nE = 0; p->sumdof = n; p->n = n; p->hbw = n - 1; p->nslots = n * n; p->ncontrib = 0;
    p->free_of_pos.resize(n);
    p->pos_of_free.resize(n);
    for (int i = 0; i < n; ++i) p->free_of_pos[i] = p->pos_of_free[i] = i;
    p->slot_ptr.assign(1, 0);
    p->node_ptr.assign(1, 0);
    p->npad = n + (n & 1);
    jacobi_config(n, &p->rpl, &p->lpp);
    p->LD = p->rpl * p->lpp;
    if (p->n <= 64) p->LD |= 1;
    p->nthreads = 64;
    *out = p;
    return 0;
}

namespace {
__global__ void k_dense_to_band(int n, long long B, const double* __restrict__ A, double* __restrict__ Ab) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * n * n) return;
    const long long b = t / (n * n);
    const int r = (int)(t % (n * n));
    const int d = r / n, j = r % n;                     // Ab[d][j] = A[j+d][j]
    Ab[t] = (j + d < n) ? A[b * n * n + (long long)(j + d) * n + j] : 0.0;
}
}  // namespace

int bfe2_sym_eig_dense(bfe2_plan* plan, int64_t B, const double* A, const double* Bm, double* lam, double* V,
                       int32_t* info, int32_t* sweeps, double* workspace, void* stream) {
    if (plan && B == 0) return 0;
    if (!plan || B < 0 || !A || !Bm || !lam || !V || !info || !workspace) return fail(-1, "bfe2_sym_eig_dense: bad argument");
    if (plan->nE != 0 || plan->hbw != plan->n - 1) return fail(-1, "bfe2_sym_eig_dense needs a plan from bfe2_plan_create_dense");
    const int n = plan->n;
    const long long tot = B * n * n;
    k_dense_to_band<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, B, A, workspace);
    k_dense_to_band<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, B, Bm, workspace + tot);
    BFE2_CUDA(cudaGetLastError());
    return bfe2_modal_solve(plan, B, n, workspace, workspace + tot, lam, V, info, sweeps, stream);
}

#endif  // !BFE2_TU_PACKED

}  // extern "C"
