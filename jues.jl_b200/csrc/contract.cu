// Pairwise tensor contraction on device tensors:  C[ic] = alpha * sum_K A[ia] * B[ib] + beta * C[ic].
// Index strings name the axes (first index fastest); letters shared by A and B are summed.
// The contraction is mapped onto ONE call of the sm_100a DGEMM; operands whose axis order does not
// already form a (M,K)/(K,M) resp. (K,N)/(N,K) matrix are permuted into a temporary first (the
// hot contractions of the path are laid out so that no permutation is needed).  This plays the
// role TensorOperations' @tensor / @tensoropt plays in the reference (TTGT).
#include "contract.h"
#include "contract_plan.h"
#include "dgemm.h"

#include <algorithm>

namespace jues {

namespace {

int64_t extent_of(char c, const Ten& A, const char* ia, const Ten& B, const char* ib, const Ten& C,
                  const char* ic) {
    int64_t e = -1;
    auto chk = [&](const Ten& T, const char* s) {
        const char* f = strchr(s, c);
        if (f) {
            const int64_t d = T.d[f - s];
            JUES_REQUIRE(e < 0 || e == d, "contract: inconsistent extents for an index letter");
            e = d;
        }
    };
    chk(A, ia); chk(B, ib); chk(C, ic);
    return e;
}

// permuted copy of X (axes `ix`) in axis order `tgt`: from the context's PermCache when X is a
// registered tensor (and memory allows keeping the copy), otherwise into the temporary `tmp`
const double* permuted_operand(jues_ctx* ctx, const Ten& X, const char* ix, const std::string& tgt,
                               const int64_t* tgt_dims, DTen& tmp) {
    Ten t = X;
    for (size_t q = 0; q < tgt.size(); ++q) t.d[q] = tgt_dims[q];
    PermCache* pc = static_cast<PermCache*>(ctx->perm_cache);
    if (pc) {
        for (const auto& r : pc->ranges) {
            if (X.p < r.lo || X.p >= r.hi) continue;
            const std::string key = perm_key(X, ix, tgt);
            for (auto& e : pc->entries)
                if (e.key == key) { ++e.hits; return e.buf.p; }
            size_t free_b = 0, total_b = 0;
            cudaMemGetInfo(&free_b, &total_b);
            const size_t bytes = (size_t)X.size() * 8;
            if (free_b + ctx->big_cached_bytes < 4 * bytes + (size_t(8) << 30)) break;  // keep headroom
            PermCache::Entry e;
            e.key = key;
            e.sweep = r.sweep;
            if (r.sweep) {
                e.buf.alloc(ctx, (size_t)X.size());
            } else {
                // copies of static integrals live for the whole calculation: neither a temporary of the sweep
                // that happens to build them nor a tenant of the sweep arena
                MeasurePause mp(ctx);
                ArenaPause ap(ctx);
                e.buf.alloc(ctx, (size_t)X.size());
            }
            t.p = e.buf.p;
            permute_axpby(ctx, 1.0, X, ix, 0.0, t, tgt.c_str());
            pc->entries.push_back(std::move(e));
            return pc->entries.back().buf.p;
        }
    }
    tmp.buf.alloc(ctx, (size_t)X.size());
    t.p = tmp.buf.p;
    permute_axpby(ctx, 1.0, X, ix, 0.0, t, tgt.c_str());
    return tmp.buf.p;
}

}  // namespace

std::string perm_key(const Ten& X, const char* ix, const std::string& tgt) {
    char key[160];
    snprintf(key, sizeof key, "%p:%lld,%lld,%lld,%lld:%s>%s", (const void*)X.p, (long long)X.d[0],
             (long long)X.d[1], (long long)X.d[2], (long long)X.d[3], ix, tgt.c_str());
    return key;
}

void contract(jues_ctx* ctx, double alpha, const Ten& A, const char* ia, const Ten& B, const char* ib,
              double beta, const Ten& C, const char* ic, bool batch_last, const double* Cin) {
    JUES_REQUIRE((int)strlen(ia) == A.rank && (int)strlen(ib) == B.rank && (int)strlen(ic) == C.rank,
                 "contract: index string length != tensor rank");
    auto ext = [&](char c) { return extent_of(c, A, ia, B, ib, C, ic); };
    ContractPlan p;
    try {
        p = plan_contraction(ia, ib, ic, ext, batch_last);
    } catch (const std::invalid_argument& e) {
        throw Error(JUES_B200_EINVAL, std::string("invalid argument: ") + e.what());
    }
    // ---- operands as matrices ----------------------------------------------------------------------
    DTen tmpA, tmpB, tmpC;
    const double* pa = A.p;
    const double* pb = B.p;
    if (!p.permA.empty()) {
        int64_t td[4] = {1, 1, 1, 1};
        for (size_t q = 0; q < p.permA.size(); ++q) td[q] = ext(p.permA[q]);
        pa = permuted_operand(ctx, A, ia, p.permA, td, tmpA);
    }
    if (!p.permB.empty()) {
        int64_t td[4] = {1, 1, 1, 1};
        for (size_t q = 0; q < p.permB.size(); ++q) td[q] = ext(p.permB[q]);
        pb = permuted_operand(ctx, B, ib, p.permB, td, tmpB);
    }
    GemmCall g;
    g.transA = p.transX; g.transB = p.transY;
    g.M = p.M; g.N = p.N; g.K = p.K; g.batch = p.batch;
    g.A = p.swapped ? pb : pa; g.lda = p.ldx; g.strideA = p.strideX;
    g.B = p.swapped ? pa : pb; g.ldb = p.ldy; g.strideB = p.strideY;
    g.ldc = p.ldc; g.strideC = p.strideC;
    JUES_REQUIRE(Cin == nullptr || p.tempC.empty(), "contract: Cin with an interleaved output");
    if (p.tempC.empty()) {
        g.C = C.p;
        g.alpha = alpha; g.beta = beta;
        g.Cin = Cin;
        dgemm(ctx, g);
    } else {
        // interleaved output: GEMM into a temporary [M..., N...(, batch)] and permute-accumulate into C
        tmpC.buf.alloc(ctx, (size_t)C.size());
        Ten t = C; t.p = tmpC.buf.p;
        for (size_t q = 0; q < p.tempC.size(); ++q) t.d[q] = ext(p.tempC[q]);
        g.C = tmpC.buf.p;
        g.alpha = 1.0; g.beta = 0.0;
        dgemm(ctx, g);
        permute_axpby(ctx, alpha, t, p.tempC.c_str(), beta, C, ic);
    }
}

}  // namespace jues
