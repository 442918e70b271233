// tools/beam_host_check.cu -- DEVELOPMENT TOOL, not part of the product and not shipped in libbfe2.so.
//
// Runs the per-thread routines of beamfe2_b200/csrc/bfe2_beam_core.cuh (the arithmetic of the config-4 kernels)
// in host loops, so that the double-double formulation can be checked at n = 1e5 in a container without a GPU.
// Build:  nvcc -O2 -std=c++17 -Xcompiler -fPIC,-fopenmp,-ffp-contract=off -shared -o /tmp/libbeam_host.so tools/beam_host_check.cu
// Driver: tools/beam_host_check.py
#include <cstdint>
#include <cmath>
#include <cstdlib>
#include <vector>

#include "../beamfe2_b200/csrc/bfe2_beam_core.cuh"

using namespace bfe2;
using namespace bfe2::chain;

template <class S> struct HostBeam {
    long long N;
    int P, P2 = 0;                                     // P2 > 0: the separator system is a second-level chain
    std::vector<int> sep, part, sep2, part2;
    std::vector<S> Kd, Kl, Md, Ml, Sinv, G, Vl, Vr, RSinv, RG, Rl;
    std::vector<S> Kd2, Kl2, Sinv2, G2, Vl2, Vr2;      // level 2: N2 = P - 1 block rows
    std::vector<unsigned char> fixed;
};

// partition factor + spikes + sequential separator system of one level
template <class S>
static int factor_level(long long N, int P, const std::vector<int>& sep, const S* Kd, const S* Kl, S* Sinv, S* G, S* Vl,
                        S* Vr, S* RSinv, S* RG, S* Rl, bool reduced) {
    int bad = 0;
#pragma omp parallel for
    for (int p = 0; p < P; ++p) {
        const int b = factor_partition<S>(sep[p] + 1, sep[p + 1] - 1, Kd, Kl, Sinv, G);
        if (b) bad = b;
    }
#pragma omp parallel for
    for (int t = 0; t < 6 * P; ++t) spike_column<S>(t / 6, t % 6, N, sep.data(), Kl, Sinv, G, Vl, Vr);
    if (!bad && reduced) bad = reduced_factor<S>(P, sep.data(), Kd, Kl, Vl, Vr, RSinv, RG, Rl);
    return bad;
}

template <class S> static HostBeam<S>* create(long long n_elem, const double* coords, const double* props,
                                              const unsigned char* fixed, int P, int* info) {
    auto* h = new HostBeam<S>();
    h->N = n_elem + 1;
    const long long N = h->N;
    choose_levels(N, P, h->P, h->P2);
    P = h->P;
    cut_chain(N, P, h->sep, h->part);
    h->fixed.assign(fixed, fixed + 3 * N);
    for (auto* v : {&h->Kd, &h->Kl, &h->Md, &h->Ml, &h->Sinv, &h->G, &h->Vl, &h->Vr}) v->resize(9 * N);
    for (auto* v : {&h->RSinv, &h->RG, &h->Rl}) v->resize(9 * (P + 1));
#pragma omp parallel for
    for (long long k = 0; k < N; ++k) {
        assemble_node<S>(k, N, coords, props, h->fixed.data(), false, h->Kd.data(), h->Kl.data());
        assemble_node<S>(k, N, coords, props, h->fixed.data(), true, h->Md.data(), h->Ml.data());
    }
    const bool two = h->P2 > 0;
    int bad = factor_level<S>(N, P, h->sep, h->Kd.data(), h->Kl.data(), h->Sinv.data(), h->G.data(), h->Vl.data(),
                              h->Vr.data(), h->RSinv.data(), h->RG.data(), h->Rl.data(), !two);
    if (!bad && two) {
        const long long N2 = P - 1;
        cut_chain(N2, h->P2, h->sep2, h->part2);
        for (auto* v : {&h->Kd2, &h->Kl2, &h->Sinv2, &h->G2, &h->Vl2, &h->Vr2}) v->assign(9 * N2, S{});
        for (auto* v : {&h->RSinv, &h->RG, &h->Rl}) v->assign(9 * (h->P2 + 1), S{});
#pragma omp parallel for
        for (int j = 0; j < (int)N2; ++j)
            reduced_build_node<S>(j, h->sep.data(), h->Kd.data(), h->Kl.data(), h->Vl.data(), h->Vr.data(), h->Kd2.data(),
                                  h->Kl2.data());
        bad = factor_level<S>(N2, h->P2, h->sep2, h->Kd2.data(), h->Kl2.data(), h->Sinv2.data(), h->G2.data(),
                              h->Vl2.data(), h->Vr2.data(), h->RSinv.data(), h->RG.data(), h->Rl.data(), true);
    }
    *info = bad;
    return h;
}

template <class S> static void solve(HostBeam<S>* h, int nrhs, const double* F, double* U, double* Ulo) {
    const long long n = 3 * h->N;
    for (long long i = 0; i < n * nrhs; ++i) { U[i] = F[i]; if (Ulo) Ulo[i] = 0.0; }
    MVec<S> mv{U, Ulo, nrhs};
    const int P = h->P;
#pragma omp parallel for collapse(2)
    for (int p = 0; p < P; ++p)
        for (int r = 0; r < nrhs; ++r)
            part_solve_inplace<S>(mv, r, 3LL * nrhs, nrhs, h->sep[p] + 1, h->sep[p + 1] - 1, h->Sinv.data(), h->G.data());
    if (P > 1 && h->P2 == 0) {
#pragma omp parallel for
        for (int r = 0; r < nrhs; ++r)
            solve_reduced_rhs<S>(P, r, h->sep.data(), h->Kl.data(), h->RSinv.data(), h->RG.data(), mv);
    }
    if (h->P2 > 0) {                                   // the separator system through the second level
        const long long N2 = P - 1;
        std::vector<double> u2(3 * N2 * nrhs), u2lo(3 * N2 * nrhs, 0.0);
        MVec<S> mv2{u2.data(), Ulo ? u2lo.data() : nullptr, nrhs};
#pragma omp parallel for collapse(2)
        for (int j = 0; j < (int)N2; ++j)
            for (int r = 0; r < nrhs; ++r) reduced_rhs_node<S>(j, r, h->sep.data(), h->Kl.data(), mv, mv2);
#pragma omp parallel for collapse(2)
        for (int p = 0; p < h->P2; ++p)
            for (int r = 0; r < nrhs; ++r)
                part_solve_inplace<S>(mv2, r, 3LL * nrhs, nrhs, h->sep2[p] + 1, h->sep2[p + 1] - 1, h->Sinv2.data(),
                                      h->G2.data());
        if (h->P2 > 1) {
#pragma omp parallel for
            for (int r = 0; r < nrhs; ++r)
                solve_reduced_rhs<S>(h->P2, r, h->sep2.data(), h->Kl2.data(), h->RSinv.data(), h->RG.data(), mv2);
#pragma omp parallel for
            for (long long k = 0; k < N2; ++k) {
                if (h->part2[k] < 0) continue;
                for (int r = 0; r < nrhs; ++r)
                    solve_fix_node<S>(k, r, h->part2[k], N2, h->sep2.data(), h->Vl2.data(), h->Vr2.data(), mv2);
            }
        }
#pragma omp parallel for collapse(2)
        for (int j = 0; j < (int)N2; ++j)
            for (int r = 0; r < nrhs; ++r) reduced_scatter_node<S>(j, r, h->sep.data(), mv2, mv);
    }
    if (P > 1) {
#pragma omp parallel for
        for (long long k = 0; k < h->N; ++k) {
            if (h->part[k] < 0) continue;
            for (int r = 0; r < nrhs; ++r)
                solve_fix_node<S>(k, r, h->part[k], h->N, h->sep.data(), h->Vl.data(), h->Vr.data(), mv);
        }
    }
}

template <class S> static void matvec(HostBeam<S>* h, int which, int nrhs, const double* X, const double* Xlo,
                                      const double* F, double* Y) {
    MVec<S> mv{const_cast<double*>(X), const_cast<double*>(Xlo), nrhs};
#pragma omp parallel for
    for (long long k = 0; k < h->N; ++k)
        for (int r = 0; r < nrhs; ++r)
            matvec_node<S>(k, r, h->N, (which ? h->Md : h->Kd).data(), (which ? h->Ml : h->Kl).data(), mv, F, Y);
}

extern "C" {
void* hb_create(int prec, long long n_elem, const double* coords, const double* props, const unsigned char* fixed,
                int P, int* info) {
    return prec ? (void*)create<dd>(n_elem, coords, props, fixed, P, info)
                : (void*)create<double>(n_elem, coords, props, fixed, P, info);
}
void hb_solve(int prec, void* h, int nrhs, const double* F, double* U, double* Ulo) {
    if (prec) solve<dd>((HostBeam<dd>*)h, nrhs, F, U, Ulo);
    else solve<double>((HostBeam<double>*)h, nrhs, F, U, nullptr);
}
void hb_matvec(int prec, void* h, int which, int nrhs, const double* X, const double* Xlo, const double* F, double* Y) {
    if (prec) matvec<dd>((HostBeam<dd>*)h, which, nrhs, X, Xlo, F, Y);
    else matvec<double>((HostBeam<double>*)h, which, nrhs, X, nullptr, F, Y);
}
void hb_destroy(int prec, void* h) {
    if (prec) delete (HostBeam<dd>*)h;
    else delete (HostBeam<double>*)h;
}
int hb_partitions(int prec, void* h) { return prec ? ((HostBeam<dd>*)h)->P : ((HostBeam<double>*)h)->P; }
}
