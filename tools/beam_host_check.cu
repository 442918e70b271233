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
    int P;
    std::vector<int> sep, part;
    std::vector<S> Kd, Kl, Md, Ml, Sinv, G, Vl, Vr, RSinv, RG, Rl;
    std::vector<unsigned char> fixed;
};

template <class S> static HostBeam<S>* create(long long n_elem, const double* coords, const double* props,
                                              const unsigned char* fixed, int P, int* info) {
    auto* h = new HostBeam<S>();
    h->N = n_elem + 1;
    const long long N = h->N;
    if (P <= 0) P = (int)std::min<long long>(4096, std::max<long long>(1, (long long)sqrt((double)N)));
    while (P > 1 && N < 3LL * P) --P;
    h->P = P;
    h->sep.resize(P + 1);
    h->sep[0] = -1;
    h->sep[P] = (int)N;
    for (int j = 1; j < P; ++j) h->sep[j] = (int)((long long)j * N / P);
    h->part.assign(N, 0);
    for (int p = 0; p < P; ++p) {
        for (int k = h->sep[p] + 1; k < h->sep[p + 1]; ++k) h->part[k] = p;
        if (p + 1 < P) h->part[h->sep[p + 1]] = -1;
    }
    h->fixed.assign(fixed, fixed + 3 * N);
    for (auto* v : {&h->Kd, &h->Kl, &h->Md, &h->Ml, &h->Sinv, &h->G, &h->Vl, &h->Vr}) v->resize(9 * N);
    for (auto* v : {&h->RSinv, &h->RG, &h->Rl}) v->resize(9 * (P + 1));
#pragma omp parallel for
    for (long long k = 0; k < N; ++k) {
        assemble_node<S>(k, N, coords, props, h->fixed.data(), false, h->Kd.data(), h->Kl.data());
        assemble_node<S>(k, N, coords, props, h->fixed.data(), true, h->Md.data(), h->Ml.data());
    }
    int bad = 0;
#pragma omp parallel for
    for (int p = 0; p < P; ++p) {
        const int b = factor_partition<S>(h->sep[p] + 1, h->sep[p + 1] - 1, h->Kd.data(), h->Kl.data(), h->Sinv.data(),
                                          h->G.data());
        if (b) bad = b;
    }
#pragma omp parallel for
    for (int t = 0; t < 6 * P; ++t)
        spike_column<S>(t / 6, t % 6, N, h->sep.data(), h->Kl.data(), h->Sinv.data(), h->G.data(), h->Vl.data(),
                        h->Vr.data());
    if (!bad)
        bad = reduced_factor<S>(P, h->sep.data(), h->Kd.data(), h->Kl.data(), h->Vl.data(), h->Vr.data(),
                                h->RSinv.data(), h->RG.data(), h->Rl.data());
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
    if (P > 1) {
#pragma omp parallel for
        for (int r = 0; r < nrhs; ++r)
            solve_reduced_rhs<S>(P, r, h->sep.data(), h->Kl.data(), h->RSinv.data(), h->RG.data(), mv);
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
