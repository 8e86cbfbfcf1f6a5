// 4-index AO->MO transformation on the device.  Internal.
#pragma once
#include "tensor_ops.h"

namespace jues {

// Where the AO integrals come from.  `slab(lo, n)` returns a device pointer to
// gao[:, :, :, lo:lo+n] (padded extents np^3 x n), valid until the next slab() call.
struct GaoSource {
    int64_t n = 0;    // logical nao
    int64_t np = 0;   // padded (even) nao
    bool phys = false; // non-resident sources: produce slabs as g'[mu,lam,nu,sig] = g[mu,nu,lam,sig]
    // non-resident sources: restrict the streamed sigma range to [sig_lo, sig_hi) (hi < 0: all).  The
    // transform is linear in gao, so ranks that each stream a disjoint sigma range obtain partial
    // results whose sum (one all-reduce) is the full transform.
    int64_t sig_lo = 0, sig_hi = -1;
    virtual ~GaoSource() {}
    virtual bool resident() const = 0;          // whole tensor addressable on the device
    virtual const double* base() const { return nullptr; }
    virtual const double* slab(jues_ctx* ctx, int64_t lo, int64_t cnt) = 0;
    // Optional: a block of the THIRD index, gao[:, :, lo:lo+cnt, :] as a dense (np, np, cnt, np) array.
    // With it the first quarter contracts the complete sigma range in one GEMM per block (K = N, no
    // read-modify-write of the N^3 x d4 accumulator per sigma slab).
    virtual bool has_block3() const { return false; }
    virtual const double* block3(jues_ctx*, int64_t, int64_t) { return nullptr; }
    // Every source: dst[mu,lam,n,sig] = gao[mu, lo+n, lam, sig], n < cnt -- a block of the SECOND chemists'
    // index laid out in physicists' order g'[mu,lam,nu,sig] (dense, np x np x cnt x np, zero padding).
    // This is the unit the sharded transform distributes over ranks and streams through the device.
    virtual void phys_block(jues_ctx* ctx, int64_t lo, int64_t cnt, double* dst) = 0;
};

// dense device tensor (np^4)
struct DeviceGao : GaoSource {
    const double* p;
    DeviceGao(const double* p_, int64_t n_, int64_t np_) : p(p_) { n = n_; np = np_; }
    bool resident() const override { return true; }
    const double* base() const override { return p; }
    const double* slab(jues_ctx*, int64_t lo, int64_t) override { return p + lo * np * np * np; }
    void phys_block(jues_ctx* ctx, int64_t lo, int64_t cnt, double* dst) override;
};

// host (caller-owned, unpadded column-major n^4) streamed through pinned staging in sigma slabs
struct HostGao : GaoSource {
    const double* h;
    DBuf stage, stage2;
    int64_t stage_cnt = 0;
    HostGao(const double* h_, int64_t n_, int64_t np_) : h(h_) { n = n_; np = np_; }
    bool resident() const override { return false; }
    const double* slab(jues_ctx* ctx, int64_t lo, int64_t cnt) override;
    void phys_block(jues_ctx* ctx, int64_t lo, int64_t cnt, double* dst) override;
};

// counter-based synthetic ERIs generated per slab
struct SynthGao : GaoSource {
    unsigned long long seed;
    double scale;
    DBuf stage;
    SynthGao(int64_t n_, int64_t np_, unsigned long long s, double sc) : seed(s), scale(sc) { n = n_; np = np_; }
    bool resident() const override { return false; }
    const double* slab(jues_ctx* ctx, int64_t lo, int64_t cnt) override;
    bool has_block3() const override { return true; }
    const double* block3(jues_ctx* ctx, int64_t lo, int64_t cnt) override;
    void phys_block(jues_ctx* ctx, int64_t lo, int64_t cnt, double* dst) override;
};

// Upload host matrix C (n x d, column-major, ld = n) into a zero-padded device matrix (np x dp).
void upload_padded_matrix(jues_ctx* ctx, DBuf& dst, const double* host, int64_t n, int64_t d, int64_t np,
                          int64_t dp);
// Upload host gao (n^4) into a zero-padded dense device tensor (np^4).
void upload_padded_gao(jues_ctx* ctx, double* dst, const double* host, int64_t n, int64_t np);

// out[p0,p1,p2,p3] (padded extents dp[0..3], chemists' order: C1 on slot 1 ...) =
//   sum C1[mu,p0] C2[nu,p1] C3[lam,p2] C4[sig,p3] gao[mu,nu,lam,sig].
// Ck are device matrices (np x dp[k]).  The order in which the four indices are contracted is
// chosen to minimise the flop count (the last AO index first when gao is streamed).
// If `reference_order` is set the reference's fixed order (sigma, lambda, nu, mu;
// Transformation.jl:68-91) is used.
// Ping/pong buffers for the three intermediate quarter-transformed tensors; pass the same workspace
// to successive transforms to avoid re-allocating multi-GB blocks for every integral class.
struct TransformWorkspace {
    DBuf buf[2];
};
void tei_transform_dev(jues_ctx* ctx, GaoSource& gao, const double* const Cm[4], const int64_t dp[4],
                       double* out, bool reference_order = false, TransformWorkspace* ws = nullptr);

// ---- the sharded one-pass transform (physicists' order) ----------------------------------------------
// out[p,q,r,s] = <pq|rs> = sum Cp[mu,p] Cq[lam,q] Cr[nu,r] Cs[sig,s] gao[mu,nu,lam,sig]   (= (pr|qs))
// for ALL p, q, r of the given column sets and the LAST index restricted to this rank's share of the
// columns of Cs_all: rank d owns the s_counts[d] columns that follow those of ranks 0..d-1.
//
// One pass over the AO integrals, shared by the ranks (Transformation.jl:39-93 contracts sigma, lambda,
// nu, mu in turn; ParCCD.jl:31-40 is the reference's only multi-process transform):
//   1. rank r reads / generates / uploads only the blocks g'[mu,lam,nu in B_r,sig] of ITS share B_r of the
//      AO index nu (1/P of the tensor), in sub-blocks that are streamed through the device;
//   2. three local quarter transforms per sub-block -- mu->p, lam->q, sig->s for the columns of EVERY rank;
//   3. one personalised exchange: the (p,q,nu in B_r,s in S_d) block goes to rank d (NCCL send/recv over
//      NVLink; a plain copy with one rank);
//   4. the last quarter nu->r over the complete, re-assembled nu range.
// Flops 8 N^5 / P per rank for full column sets, AO traffic N^4 / P per rank, one exchange of N^4 / P.
// `out` is allocated here: dense (dP, dQ, dR, s_counts[rank]).
void tei_transform_sharded(jues_ctx* ctx, GaoSource& gao, const double* Cp, int64_t dP, const double* Cq,
                           int64_t dQ, const double* Cr, int64_t dR, const double* Cs_all,
                           const std::vector<int64_t>& s_counts, DBuf& out);
// this rank's share [lo, lo+cnt) of the padded AO index in the sharded transform
void ao_share(const jues_ctx* ctx, int64_t np, int rank, int64_t* lo, int64_t* cnt);

// testing hook: called after every quarter transform with (step 0..3, output pointer, elements)
typedef void (*QuarterProbe)(void* user, jues_ctx* ctx, int step, const double* out, size_t n);
void set_quarter_probe(QuarterProbe fn, void* user);

// flops of the order tei_transform_dev would pick (for reporting)
double tei_transform_flops(int64_t np, const int64_t dp[4], bool reference_order, bool streamed);

}  // namespace jues
