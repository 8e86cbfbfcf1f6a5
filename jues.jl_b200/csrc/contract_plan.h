// Planning of a pairwise tensor contraction as ONE (batched) GEMM: pure host logic, no CUDA types, so
// that it can be exercised on the CPU (tests/host/contract_plan_test.cpp runs every plan through a naive
// executor against a brute-force contraction).  contract.cu turns a plan into permute + DGEMM launches.
//
//   C[ic] = alpha * sum_{letters shared by A and B} A[ia] * B[ib] + beta * C[ic]      (first index fastest)
//
// batch_last: the LAST letter of ic is a batch index -- it must also be the last letter of ia and/or ib and
// appear nowhere else.  The contraction then runs as `extent(batch letter)` independent products in one
// launch (an operand that does not carry the letter is shared, batch stride 0).  Used where the output of
// the plain mapping would be interleaved (M and N letters mixed in C), which costs a permutation pass over
// C: e.g. WJ[m,e,j,b] += sum_f <ef|mb> t[j,f] is, for every b, the matrix product [(m,e) x f][f x j].
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>

namespace jues {

struct ContractPlan {
    // permuted copies the operands need first ("" = the operand is used where it lies); the strings are the
    // axis orders of the copies, batch letter (if the operand carries it) last
    std::string permA, permB;
    // the GEMM as issued: C_g = op(X) op(Y) with (X, Y) = (A, B), or (B, A) when `swapped` (C stored [N..., M...])
    bool swapped = false;
    bool transX = false, transY = false;
    int64_t M = 0, N = 0, K = 0, batch = 1;         // of the issued GEMM
    int64_t ldx = 0, ldy = 0, ldc = 0;
    int64_t strideX = 0, strideY = 0, strideC = 0;  // batch strides (0: shared operand)
    // "" = the GEMM writes C directly (alpha, beta applied by it); otherwise it writes a dense temporary with
    // this axis order (alpha = 1, beta = 0) which is then permute-accumulated into C
    std::string tempC;
    bool khuge = false;                             // K-huge product: operands wanted contiguous along M / N
};

namespace plan_detail {

inline bool has(const std::string& s, char c) { return s.find(c) != std::string::npos; }
inline std::string pick(const std::string& from, const std::string& set) {
    std::string r;
    for (char c : from)
        if (has(set, c)) r.push_back(c);
    return r;
}
inline void require(bool ok, const char* msg) {
    if (!ok) throw std::invalid_argument(msg);
}

}  // namespace plan_detail

// ext(letter) -> extent.  sizeA / sizeB: element counts of the operands (for the cheaper-permutation choice).
template <class ExtFn>
ContractPlan plan_contraction(const char* ia, const char* ib, const char* ic, ExtFn ext, bool batch_last) {
    using namespace plan_detail;
    std::string sa(ia), sb(ib), sc(ic);
    ContractPlan p;
    char bl = 0;
    bool blA = false, blB = false;
    if (batch_last) {
        require(!sc.empty(), "contract: batch_last with a scalar result");
        bl = sc.back();
        blA = !sa.empty() && sa.back() == bl;
        blB = !sb.empty() && sb.back() == bl;
        require(blA || blB, "contract: the batch letter must be the last index of A or B");
        sc.pop_back();
        if (blA) sa.pop_back();
        if (blB) sb.pop_back();
        require(!has(sa, bl) && !has(sb, bl) && !has(sc, bl), "contract: the batch letter appears twice");
        p.batch = ext(bl);
    }
    std::string Mset, Nset, Kset;
    for (char c : sa) {
        const bool inB = has(sb, c), inC = has(sc, c);
        require(inB != inC, "contract: every index of A must appear in exactly one of B, C");
        (inC ? Mset : Kset).push_back(c);
    }
    for (char c : sb) {
        const bool inA = has(sa, c), inC = has(sc, c);
        require(inA != inC, "contract: every index of B must appear in exactly one of A, C");
        if (inC) Nset.push_back(c);
    }
    require(Mset.size() + Nset.size() == sc.size(), "contract: C has indices found in neither A nor B");
    require(!Kset.empty(), "contract: no summed index");
    auto extent = [&](const std::string& s) {
        int64_t n = 1;
        for (char c : s) n *= ext(c);
        return n;
    };
    // ---- orders of the M, N and K groups -----------------------------------------------------------
    const std::string mC = pick(sc, Mset), nC = pick(sc, Nset);
    const bool c_mn = sc == mC + nC;  // C = [M..., N...]
    const bool c_nm = sc == nC + mC;  // C = [N..., M...]
    std::string mord, nord;
    if (c_mn || c_nm) { mord = mC; nord = nC; }
    else { mord = pick(sa, Mset); nord = pick(sb, Nset); }
    // K order: prefer the order that leaves the LARGER operand unpermuted
    const std::string kA = pick(sa, Kset), kB = pick(sb, Kset);
    auto conforms = [&](const std::string& idx, const std::string& r, const std::string& k) {
        return idx == r + k || idx == k + r;
    };
    std::string kord = kA;
    if (kA != kB) {
        const bool a_ok_kA = conforms(sa, mord, kA), b_ok_kB = conforms(sb, nord, kB);
        const bool a_ok_kB = conforms(sa, mord, kB), b_ok_kA = conforms(sb, nord, kA);
        const int64_t szA = extent(sa), szB = extent(sb);   // per batch member: the ratio is what matters
        const int64_t cost_kA = (a_ok_kA ? 0 : szA) + (b_ok_kA ? 0 : szB);
        const int64_t cost_kB = (a_ok_kB ? 0 : szA) + (b_ok_kB ? 0 : szB);
        kord = cost_kB < cost_kA ? kB : kA;
    }
    const int64_t M = extent(mord), N = extent(nord), K = extent(kord);
    // A product with a huge K and small M, N is bandwidth bound: the streaming kernel (skinny.cu) wants both
    // operands contiguous along their SMALL index, so an operand that must be re-ordered anyway is put there
    p.khuge = p.batch == 1 && M <= 128 && N <= 128 && M * N <= 4096 && K >= 8192;
    bool a_t, b_t;
    if (sa == mord + kord) a_t = false;            // stored M x K
    else if (sa == kord + mord) a_t = true;        // stored K x M
    else {
        p.permA = (p.khuge ? mord + kord : kord + mord) + (blA ? std::string(1, bl) : std::string());
        a_t = !p.khuge;
    }
    if (sb == kord + nord) b_t = false;            // stored K x N
    else if (sb == nord + kord) b_t = true;        // stored N x K
    else {
        p.permB = (p.khuge ? nord + kord : kord + nord) + (blB ? std::string(1, bl) : std::string());
        b_t = p.khuge;
    }
    const int64_t lda = a_t ? K : M, ldb = b_t ? N : K;
    const int64_t sA = blA ? M * K : 0, sB = blB ? K * N : 0;
    p.K = K;
    if (c_mn || (!c_mn && !c_nm)) {
        p.swapped = false;
        p.transX = a_t; p.transY = b_t;
        p.M = M; p.N = N;
        p.ldx = lda; p.ldy = ldb;
        p.strideX = sA; p.strideY = sB;
    } else {
        // C stored [N..., M...]:  C^T(N x M) = op(B)^T op(A)^T
        p.swapped = true;
        p.transX = !b_t; p.transY = !a_t;
        p.M = N; p.N = M;
        p.ldx = ldb; p.ldy = lda;
        p.strideX = sB; p.strideY = sA;
    }
    p.ldc = p.M;
    p.strideC = M * N;
    if (!(c_mn || c_nm)) p.tempC = mord + nord + (batch_last ? std::string(1, bl) : std::string());
    return p;
}

}  // namespace jues
