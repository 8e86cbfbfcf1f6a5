// Pairwise tensor contraction on device tensors:  C[ic] = alpha * sum_K A[ia] * B[ib] + beta * C[ic].
// Index strings name the axes (first index fastest); letters shared by A and B are summed.
// The contraction is mapped onto ONE call of the sm_100a DGEMM; operands whose axis order does not
// already form a (M,K)/(K,M) resp. (K,N)/(N,K) matrix are permuted into a temporary first (the
// hot contractions of the path are laid out so that no permutation is needed).  This plays the
// role TensorOperations' @tensor / @tensoropt plays in the reference (TTGT).
#include "contract.h"
#include "dgemm.h"

#include <algorithm>

namespace jues {

namespace {

struct Grp {
    std::string s;
    int64_t n = 1;
};

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

// does `idx` equal first+second ?
bool is_concat(const std::string& idx, const std::string& first, const std::string& second) {
    return idx == first + second;
}

// letters of `from` that are in `set`, in the order they appear in `from`
std::string pick(const char* from, const std::string& set) {
    std::string r;
    for (const char* p = from; *p; ++p)
        if (set.find(*p) != std::string::npos) r.push_back(*p);
    return r;
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
            char key[160];
            snprintf(key, sizeof key, "%p:%lld,%lld,%lld,%lld:%s>%s", (const void*)X.p, (long long)X.d[0],
                     (long long)X.d[1], (long long)X.d[2], (long long)X.d[3], ix, tgt.c_str());
            for (auto& e : pc->entries)
                if (e.key == key) return e.buf.p;
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

void contract(jues_ctx* ctx, double alpha, const Ten& A, const char* ia, const Ten& B, const char* ib,
              double beta, const Ten& C, const char* ic) {
    JUES_REQUIRE((int)strlen(ia) == A.rank && (int)strlen(ib) == B.rank && (int)strlen(ic) == C.rank,
                 "contract: index string length != tensor rank");
    std::string sa(ia), sb(ib), sc(ic);
    std::string Mset, Nset, Kset;
    for (char c : sa) {
        const bool inB = sb.find(c) != std::string::npos, inC = sc.find(c) != std::string::npos;
        JUES_REQUIRE(inB != inC, "contract: every index of A must appear in exactly one of B, C");
        (inC ? Mset : Kset).push_back(c);
    }
    for (char c : sb) {
        const bool inA = sa.find(c) != std::string::npos, inC = sc.find(c) != std::string::npos;
        JUES_REQUIRE(inA != inC, "contract: every index of B must appear in exactly one of A, C");
        if (inC) Nset.push_back(c);
    }
    JUES_REQUIRE(Mset.size() + Nset.size() == sc.size(), "contract: C has indices found in neither A nor B");
    JUES_REQUIRE(!Kset.empty(), "contract: no summed index");

    auto ext = [&](const std::string& s) {
        int64_t n = 1;
        for (char c : s) n *= extent_of(c, A, ia, B, ib, C, ic);
        return n;
    };

    // ---- choose the orders of the M, N and K groups ---------------------------------------------
    std::string mC = pick(ic, Mset), nC = pick(ic, Nset);
    bool c_mn = is_concat(sc, mC, nC);  // C = [M..., N...]
    bool c_nm = is_concat(sc, nC, mC);  // C = [N..., M...]
    std::string mord, nord;
    if (c_mn || c_nm) { mord = mC; nord = nC; }
    else { mord = pick(ia, Mset); nord = pick(ib, Nset); }
    // K order: prefer the order that leaves the LARGER operand unpermuted
    const std::string kA = pick(ia, Kset), kB = pick(ib, Kset);
    auto conforms = [&](const std::string& idx, const std::string& r, const std::string& k) {
        return is_concat(idx, r, k) || is_concat(idx, k, r);
    };
    std::string kord = kA;
    if (kA != kB) {
        const bool a_ok_kA = conforms(sa, mord, kA), b_ok_kB = conforms(sb, nord, kB);
        const bool a_ok_kB = conforms(sa, mord, kB), b_ok_kA = conforms(sb, nord, kA);
        const int64_t szA = A.size(), szB = B.size();
        // cost = elements that must be permuted
        const int64_t cost_kA = (a_ok_kA ? 0 : szA) + (b_ok_kA ? 0 : szB);
        const int64_t cost_kB = (a_ok_kB ? 0 : szA) + (b_ok_kB ? 0 : szB);
        kord = cost_kB < cost_kA ? kB : kA;
    }
    const int64_t M = ext(mord), N = ext(nord), K = ext(kord);

    // ---- operands as matrices ----------------------------------------------------------------------
    DTen tmpA, tmpB, tmpC;
    const double* pa = A.p;
    const double* pb = B.p;
    bool a_t, b_t;  // BLAS transposition flags
    // A product with a huge K and small M, N is bandwidth bound: the streaming kernel (skinny.cu) wants both
    // operands contiguous along their SMALL index, so an operand that must be re-ordered anyway is put there
    const bool khuge = M <= 128 && N <= 128 && M * N <= 4096 && K >= 8192;
    if (is_concat(sa, mord, kord)) a_t = false;       // stored M x K
    else if (is_concat(sa, kord, mord)) a_t = true;   // stored K x M
    else {
        // permute into K-contiguous form [K..., M...] (K-huge products: [M..., K...])
        const std::string tgt = khuge ? mord + kord : kord + mord;
        int64_t td[4] = {1, 1, 1, 1};
        for (size_t q = 0; q < tgt.size(); ++q) td[q] = extent_of(tgt[q], A, ia, B, ib, C, ic);
        pa = permuted_operand(ctx, A, ia, tgt, td, tmpA);
        a_t = !khuge;
    }
    if (is_concat(sb, kord, nord)) b_t = false;       // stored K x N
    else if (is_concat(sb, nord, kord)) b_t = true;   // stored N x K
    else {
        const std::string tgt = khuge ? nord + kord : kord + nord;
        int64_t td[4] = {1, 1, 1, 1};
        for (size_t q = 0; q < tgt.size(); ++q) td[q] = extent_of(tgt[q], A, ia, B, ib, C, ic);
        pb = permuted_operand(ctx, B, ib, tgt, td, tmpB);
        b_t = khuge;
    }

    GemmCall g;
    g.K = K;
    g.batch = 1;
    if (c_mn || (!c_mn && !c_nm)) {
        // C(M x N) = op(A) op(B)
        g.transA = a_t; g.transB = b_t;
        g.M = M; g.N = N;
        g.A = pa; g.lda = a_t ? K : M;
        g.B = pb; g.ldb = b_t ? N : K;
    } else {
        // C stored [N..., M...]:  C^T(N x M) = op(B)^T op(A)^T
        g.transA = !b_t; g.transB = !a_t;
        g.M = N; g.N = M;
        g.A = pb; g.lda = b_t ? N : K;
        g.B = pa; g.ldb = a_t ? K : M;
    }
    if (c_mn || c_nm) {
        g.C = C.p; g.ldc = g.M;
        g.alpha = alpha; g.beta = beta;
        dgemm(ctx, g);
    } else {
        // interleaved output: GEMM into a temporary [M..., N...] and permute-accumulate into C
        const std::string tgt = mord + nord;
        tmpC.buf.alloc(ctx, (size_t)C.size());
        Ten t = C; t.p = tmpC.buf.p;
        for (size_t q = 0; q < tgt.size(); ++q) t.d[q] = extent_of(tgt[q], A, ia, B, ib, C, ic);
        g.C = tmpC.buf.p; g.ldc = M;
        g.alpha = 1.0; g.beta = 0.0;
        dgemm(ctx, g);
        permute_axpby(ctx, alpha, t, tgt.c_str(), beta, C, ic);
    }
}

}  // namespace jues
