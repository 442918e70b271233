// tools/te_host_check.cu -- DEVELOPMENT TOOL, not part of the product: runs the modal block of
// beamfe2_b200/csrc/bfe2_te.cuh (the two-ended dense engine) on the CPU, 64 std::threads + one std::barrier per block,
// so that its numerics can be checked against the 40-digit truth fixture without a GPU.
// Build:  g++ -std=c++20 -O2 -pthread -fPIC -shared -DBFE2_TE_HOST -o /tmp/libte_host.so -x c++ tools/te_host_check.cu
#include <thread>
#include <vector>

#include "../beamfe2_b200/csrc/bfe2_te.cuh"

using namespace bfe2::te;

extern "C" int te_modal(int n, int b, const double* Kb, const double* Mb, int n_modes, double* lam, double* phi) {
    const int LD = n | 1;
    const Smem S = smem_layout(n, LD, n_modes);
    std::vector<double> base(S.total, 0.0), sKb(Kb, Kb + (b + 1) * n), sMb(Mb, Mb + (b + 1) * n), invK(n), invM(n);
    std::barrier<> bar(NT);
    std::vector<int> rc(NT, 0), infok(NT, 0);
    std::vector<std::thread> th;
    for (int t = 0; t < NT; ++t)
        th.emplace_back([&, t] {
            TeCtx c{t, &bar, base.data() + S.red, 0};
            rc[t] = modal_block<64>(c, n, b, sKb.data(), sMb.data(), invK.data(), invM.data(), base.data(), LD, n_modes,
                                    &infok[t]);
        });
    for (auto& x : th) x.join();
    for (int i = 0; i < n; ++i) lam[i] = base[S.lam + i];
    for (int m = 0; m < n_modes; ++m)
        for (int i = 0; i < n; ++i) phi[m * n + i] = base[S.Z + m * NT + i];
    return rc[0];
}
