// Pairwise tensor contraction helper (one DGEMM + optional permutations).  Internal.
#pragma once
#include "tensor_ops.h"

namespace jues {
// C[ic] = alpha * sum_{shared letters} A[ia] * B[ib] + beta * C[ic]
void contract(jues_ctx* ctx, double alpha, const Ten& A, const char* ia, const Ten& B, const char* ib,
              double beta, const Ten& C, const char* ic);
}  // namespace jues
