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
};

// dense device tensor (np^4)
struct DeviceGao : GaoSource {
    const double* p;
    DeviceGao(const double* p_, int64_t n_, int64_t np_) : p(p_) { n = n_; np = np_; }
    bool resident() const override { return true; }
    const double* base() const override { return p; }
    const double* slab(jues_ctx*, int64_t lo, int64_t) override { return p + lo * np * np * np; }
};

// host (caller-owned, unpadded column-major n^4) streamed through pinned staging in sigma slabs
struct HostGao : GaoSource {
    const double* h;
    DBuf stage, stage2;
    int64_t stage_cnt = 0;
    HostGao(const double* h_, int64_t n_, int64_t np_) : h(h_) { n = n_; np = np_; }
    bool resident() const override { return false; }
    const double* slab(jues_ctx* ctx, int64_t lo, int64_t cnt) override;
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

// testing hook: called after every quarter transform with (step 0..3, output pointer, elements)
typedef void (*QuarterProbe)(void* user, jues_ctx* ctx, int step, const double* out, size_t n);
void set_quarter_probe(QuarterProbe fn, void* user);

// flops of the order tei_transform_dev would pick (for reporting)
double tei_transform_flops(int64_t np, const int64_t dp[4], bool reference_order, bool streamed);

}  // namespace jues
