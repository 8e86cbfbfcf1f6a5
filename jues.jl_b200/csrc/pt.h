// Perturbative triples (T) on the device (PerturbativeTriples.jl:35-138).  Internal.
#pragma once
#include "tensor_ops.h"

namespace jues {

// Device tensors the (T) driver reads (all dense, column-major, padded even extents o, v; K = v + o):
//   Acat[a,b,p,kap] (v,v,o,K):  kap <  v: <p kap|ab>  (= ovvv[p,kap,a,b]; the CC driver's
//                                          OA[e,f,m,b] = <ef|mb> IS this block)
//                               kap >= v: -T2[p,l,a,b], l = kap - v
//   Bq[c,q,kap,r]   (v,o,K,o):  kap <  v: T2[r,q,c,kap];   kap >= v: <qr|lc> (= ooov[q,r,l,c])
//   Br[c,r,kap,q]   (v,o,K,o):  the same numbers with q and r exchanged in the layout
//   Vv[a,b,i,j]  = <ij|ab>,  t1[i,a],  eo[i], ev[a] (diagonal Fock / orbital energies; padded entries
//   far away).
// With them every X block  X(p,q,r)[(a,b),c] = sum_kap Acat[(a,b),p,kap] B[c,kap]  is ONE GEMM with the two
// contractions (over d and over l) concatenated along K, and the blocks of consecutive q (Bq) or r (Br)
// or p (rows of Acat) are one launch with N = v*nb resp. M = v^2*nb.
// nocc = number of physical occupied orbitals (the loops skip the padded ones, whose amplitudes
// and integrals vanish).  tests/pt_model.py is the numpy statement of exactly this algorithm.
struct PtInputs {
    int64_t o = 0, v = 0, nocc = 0;
    const double* Acat = nullptr;
    const double* Bq = nullptr;
    const double* Br = nullptr;
    const double* Vv = nullptr;
    const double* t1 = nullptr;
    const double* eo = nullptr;
    const double* ev = nullptr;
};

// Build Acat / Bq / Br from OAp[a,b,p,d] (nullable when `acat` already holds the block kap < v, e.g.
// all-gathered in place), T2[i,j,a,b] and ooov[m,n,i,e].  acat: (v,v,o,v+o), bq, br: (v,o,v+o,o).
void pt_build_operands(jues_ctx* ctx, int64_t o, int64_t v, const double* OAp, const double* T2,
                       const double* ooov, double* acat, double* bq, double* br);

// E(T).  Occupied pairs (i >= j) are dealt round-robin to the ranks of the context; the scalar is
// summed over ranks.  Blocking (returns the host value).  `scratch` (nullable): a device block of
// `scratch_elems` doubles the caller no longer needs (the CC driver hands over its <vv|vv> slab) --
// used for the X / W / V work arrays instead of a fresh multi-GB cudaMalloc when it holds at least
// four triples.
double pt_dev(jues_ctx* ctx, const PtInputs& in, double* scratch = nullptr, size_t scratch_elems = 0);

}  // namespace jues
