// (T): energy contribution of one virtual triple a >= b >= c of the occupied triple (i, j, k = k0 + kk),
// computed straight from the six X-block families (no W / V arrays in memory).  Shared by the CUDA
// kernel (pt.cu: pt_fused_kernel) and by a host build of the same function that the CPU test-suite
// checks against the numpy model (tests/test_pt_fused_host.py), so the index arithmetic of the kernel is
// verified without a GPU.  PerturbativeTriples.jl:96-103 (W, V) and :117-131 (energy expression).
#pragma once

#if defined(__CUDACC__)
#define JUES_HD __host__ __device__ __forceinline__
#else
#define JUES_HD inline
#endif

#include <cmath>

namespace jues {

// P = b(b+1)/2 + c with 0 <= c <= b
JUES_HD void pt_pair_decode(long long P, int* b, int* c) {
    long long bb = (long long)((sqrt(8.0 * (double)P + 1.0) - 1.0) * 0.5);
    while (bb * (bb + 1) / 2 > P) --bb;
    while ((bb + 1) * (bb + 2) / 2 <= P) ++bb;
    *b = (int)bb;
    *c = (int)(P - bb * (bb + 1) / 2);
}

struct PtFusedArgs {
    const double* X;    // six families, each kb*v^3: 0,1,4,5 as [kk][p0,p1,p2]; 2,3 as [p0,p1,kk,p2]
    const double* Vv;   // [a,b,i,j] = <ij|ab>
    const double* t1;   // [i,a]
    const double* eo;
    const double* ev;
    int o, v, i, j, k0, kb;
};

// W[x,y,z] of the triple kk (PerturbativeTriples.jl:96-101)
JUES_HD double pt_w(const PtFusedArgs& g, int kk, int x, int y, int z) {
    const long long v = g.v, v2 = v * v, v3 = v2 * v, fam = v3 * g.kb;
    const double* q = g.X + (long long)kk * v3;      // families batched over q or r
    const double* p = g.X + (long long)kk * v2;      // families batched over p
    return q[x + v * y + v2 * z]                      // X(i,j,k)[x,y,z]
         + q[fam + x + v * z + v2 * y]                // X(i,k,j)[x,z,y]
         + p[2 * fam + z + v * x + v2 * g.kb * y]     // X(k,i,j)[z,x,y]
         + p[3 * fam + z + v * y + v2 * g.kb * x]     // X(k,j,i)[z,y,x]
         + q[4 * fam + y + v * z + v2 * x]            // X(j,k,i)[y,z,x]
         + q[5 * fam + y + v * x + v2 * z];           // X(j,i,k)[y,x,z]
}

// V[x,y,z] - W[x,y,z] (PerturbativeTriples.jl:103)
JUES_HD double pt_v_extra(const PtFusedArgs& g, int k, int x, int y, int z) {
    const long long v = g.v, v2 = v * v, o = g.o;
    return g.Vv[y + v * z + v2 * (g.j + o * k)] * g.t1[g.i + o * x]
         + g.Vv[x + v * z + v2 * (g.i + o * k)] * g.t1[g.j + o * y]
         + g.Vv[x + v * y + v2 * (g.i + o * g.j)] * g.t1[k + o * z];
}

// E(a,b,c) (2 - d_ij - d_jk) / (Dd (1 + d_ab + d_bc)) for a >= b >= c (PerturbativeTriples.jl:117-131)
JUES_HD double pt_triple_energy(const PtFusedArgs& g, int kk, int a, int b, int c) {
    const int k = g.k0 + kk;
    const double Wabc = pt_w(g, kk, a, b, c), Wacb = pt_w(g, kk, a, c, b), Wbac = pt_w(g, kk, b, a, c);
    const double Wbca = pt_w(g, kk, b, c, a), Wcab = pt_w(g, kk, c, a, b), Wcba = pt_w(g, kk, c, b, a);
    const double Vabc = Wabc + pt_v_extra(g, k, a, b, c), Vacb = Wacb + pt_v_extra(g, k, a, c, b);
    const double Vbac = Wbac + pt_v_extra(g, k, b, a, c), Vbca = Wbca + pt_v_extra(g, k, b, c, a);
    const double Vcab = Wcab + pt_v_extra(g, k, c, a, b), Vcba = Wcba + pt_v_extra(g, k, c, b, a);
    const double Xs = Wabc * Vabc + Wacb * Vacb + Wbac * Vbac + Wbca * Vbca + Wcab * Vcab + Wcba * Vcba;
    const double Y = Vabc + Vbca + Vcab;
    const double Z = Vacb + Vbac + Vcba;
    const double E = (Y - 2.0 * Z) * (Wabc + Wbca + Wcab) + (Z - 2.0 * Y) * (Wacb + Wbac + Wcba) + 3.0 * Xs;
    const double occ = 2.0 - (g.i == g.j ? 1.0 : 0.0) - (g.j == k ? 1.0 : 0.0);
    const double Dd = g.eo[g.i] + g.eo[g.j] + g.eo[k] - g.ev[a] - g.ev[b] - g.ev[c];
    const double sym = 1.0 + (a == b ? 1.0 : 0.0) + (b == c ? 1.0 : 0.0);
    return E * occ / (Dd * sym);
}

}  // namespace jues
