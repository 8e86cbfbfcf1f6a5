// Host build of jues.jl_b200/csrc/pt_fused.h for the CPU test-suite: sums pt_triple_energy over every
// (kk, a >= b >= c) exactly as the CUDA kernel's threads do (test infrastructure only).
#include "../../jues.jl_b200/csrc/pt_fused.h"

extern "C" double pt_fused_host(const double* X, const double* Vv, const double* t1, const double* eo,
                                const double* ev, int o, int v, int i, int j, int k0, int kb) {
    jues::PtFusedArgs g{X, Vv, t1, eo, ev, o, v, i, j, k0, kb};
    // the kernel's work items: (kk, pair b >= c) decoded from a linear index, threads walk a = b..v-1
    const long long nbc = (long long)v * (v + 1) / 2, items = nbc * kb;
    double acc = 0.0;
    for (long long it = 0; it < items; ++it) {
        const int kk = (int)(it / nbc);
        int b, c;
        jues::pt_pair_decode(it % nbc, &b, &c);
        for (int a = b; a < v; ++a) acc += jues::pt_triple_energy(g, kk, a, b, c);
    }
    return acc;
}
