// Pairwise tensor contraction helper (one DGEMM + optional permutations).  Internal.
#pragma once
#include "tensor_ops.h"

#include <string>
#include <vector>

namespace jues {

// Memo of permuted operand copies.  contract() permutes an operand whose axis order does not form
// a matrix; when the operand lies inside a registered allocation the permuted copy is kept:
// static integrals are permuted once per calculation, amplitude-derived tensors once per sweep.
struct PermCache {
    struct Range { const double* lo; const double* hi; bool sweep; };
    struct Entry { std::string key; DBuf buf; bool sweep; long long hits = 0; };
    std::vector<Range> ranges;
    std::vector<Entry> entries;
    void add(const Ten& t, bool sweep) { ranges.push_back({t.p, t.p + t.size(), sweep}); }
    // a permuted copy somebody else already made (amp_combos writes the operand layouts of the sweep's ring
    // and Fae products while it has the amplitudes in shared memory): contract() finds it under `key`
    void provide(const std::string& key, DBuf&& buf) {
        Entry e;
        e.key = key; e.buf = std::move(buf); e.sweep = true;
        entries.push_back(std::move(e));
    }
    // drop everything derived from per-sweep tensors
    void end_sweep() {
        for (size_t k = 0; k < entries.size();)
            if (entries[k].sweep) { entries[k] = std::move(entries.back()); entries.pop_back(); } else ++k;
        for (size_t k = 0; k < ranges.size();)
            if (ranges[k].sweep) { ranges[k] = ranges.back(); ranges.pop_back(); } else ++k;
    }
    void clear() { entries.clear(); ranges.clear(); }
};

// key of the permuted copy of X (axes ix) in axis order tgt
std::string perm_key(const Ten& X, const char* ix, const std::string& tgt);

// C[ic] = alpha * sum_{shared letters} A[ia] * B[ib] + beta * C[ic]
// batch_last: the last letter of ic (also the last letter of ia and/or ib) is a batch index -- one batched
// GEMM launch instead of a GEMM with an interleaved output and a permutation pass (contract_plan.h)
// Cin: beta multiplies Cin (a tensor with the layout of C) instead of C -- a product that starts from a static
// tensor needs no copy pass (only for contractions whose output needs no permutation)
void contract(jues_ctx* ctx, double alpha, const Ten& A, const char* ia, const Ten& B, const char* ib,
              double beta, const Ten& C, const char* ic, bool batch_last = false, const double* Cin = nullptr);
}  // namespace jues
