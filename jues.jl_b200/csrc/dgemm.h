// Host-side interface of the sm_100a DGEMM (device pointers).  Internal.
#pragma once
#include "jues_common.h"

namespace jues {

struct GemmCall {
    // C[b] = alpha * op(A[b]) * op(B[b]) + beta * C[b],  b = 0..batch-1, column-major.
    // op(A) is M x K: transA=false -> A stored M x K (lda >= M); true -> stored K x M (lda >= K).
    // op(B) is K x N: transB=false -> B stored K x N (ldb >= K); true -> stored N x K (ldb >= N).
    bool transA = false, transB = false;
    int64_t M = 0, N = 0, K = 0;
    const double* A = nullptr; int64_t lda = 0; int64_t strideA = 0;
    const double* B = nullptr; int64_t ldb = 0; int64_t strideB = 0;
    double* C = nullptr;       int64_t ldc = 0; int64_t strideC = 0;
    int64_t batch = 1;
    double alpha = 1.0, beta = 0.0;
    // beta multiplies Cin instead of C when given (same ldc / strideC as C): C = alpha*op(A)op(B) + beta*Cin
    const double* Cin = nullptr;
    int force_cfg = -1;  // testing: pick a tile configuration explicitly
};

// Launch on ctx->stream (asynchronous).  Requirements (all internal tensors satisfy them because
// every orbital dimension is padded to an even count): A, B 16-byte aligned; lda, ldb, strideA,
// strideB even.  C has no alignment requirement beyond 8 bytes.
void dgemm(jues_ctx* ctx, const GemmCall& g);

// Streaming FMA kernels for bandwidth-bound skinny products (skinny.cu); returns false when the shape is not
// one of theirs (the caller then uses the tile kernel).  JUES_B200_NO_SKINNY=1 switches them off.
bool skinny_gemm(jues_ctx* ctx, const GemmCall& g);

int dgemm_num_configs();
const char* dgemm_config_name(int cfg);

}  // namespace jues
