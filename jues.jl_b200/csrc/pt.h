// Perturbative triples (T) on the device (PerturbativeTriples.jl:35-138).  Internal.
#pragma once
#include "tensor_ops.h"

namespace jues {

// Device tensors the (T) driver reads (all dense, column-major, padded even extents o, v):
//   OAp[a,b,p,d] = <pd|ab>  (= ovvv[p,d,a,b]; the CC driver's OA[e,f,m,b] = <ef|mb> IS this array)
//   Ov[l,c,q,r]  = <qr|lc>  (= ooov[q,r,l,c])
//   Vv[a,b,i,j]  = <ij|ab>
//   Tq[a,b,j,i]  = T2[i,j,a,b]
//   t1[i,a], eo[i], ev[a]   (diagonal Fock / orbital energies; padded entries far away)
// nocc = number of physical occupied orbitals (the loops skip the padded ones, whose amplitudes
// and integrals vanish).  tests/pt_model.py is the numpy statement of exactly this algorithm.
struct PtInputs {
    int64_t o = 0, v = 0, nocc = 0;
    const double* OAp = nullptr;
    const double* Ov = nullptr;
    const double* Vv = nullptr;
    const double* Tq = nullptr;
    const double* t1 = nullptr;
    const double* eo = nullptr;
    const double* ev = nullptr;
};

// E(T).  Occupied pairs (i >= j) are dealt round-robin to the ranks of the context; the scalar is
// summed over ranks.  Blocking (returns the host value).
double pt_dev(jues_ctx* ctx, const PtInputs& in);

}  // namespace jues
