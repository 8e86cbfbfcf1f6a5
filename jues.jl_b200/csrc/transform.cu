// 4-index AO->MO transformation (Transformation.jl:39-93, IntegralTransformation.jl:94) as four
// quarter transforms, each ONE (batched) launch of the sm_100a DGEMM with no permutation pass:
//   axis 0      : out[d, rest]        = C^T (d x N) * in (N x rest)                 'T','N'
//   axis 1,2    : out[pre, d, post_b] = in[pre, N, post_b] * C (N x d), batch = post 'N','N'
//   axis 3      : out[pre, d]         = in[pre, N] * C (N x d)                       'N','N'
// The contraction order is chosen to minimise flops (the reference always contracts sigma, lambda,
// nu, mu -- Transformation.jl:68-91 -- which costs 2 N^4 d4 for the first quarter whatever d4 is).
#include "transform.h"
#include "dgemm.h"
#include "dist.h"

#include <algorithm>
#include <exception>

namespace jues {

static QuarterProbe g_probe = nullptr;
static void* g_probe_user = nullptr;
void set_quarter_probe(QuarterProbe fn, void* user) { g_probe = fn; g_probe_user = user; }

const double* SynthGao::slab(jues_ctx* ctx, int64_t lo, int64_t cnt) {
    const int64_t plane = np * np * np;
    if ((int64_t)stage.n < plane * cnt) stage.alloc(ctx, (size_t)(plane * cnt));
    synth_eri_fill(ctx, stage.p, n, np, lo, cnt, seed, scale, phys);
    return stage.p;
}

const double* SynthGao::block3(jues_ctx* ctx, int64_t lo, int64_t cnt) {
    const int64_t need = np * np * cnt * np;
    if ((int64_t)stage.n < need) stage.alloc(ctx, (size_t)need);
    synth_eri_block(ctx, stage.p, n, np, lo, cnt, 0, np, seed, scale, phys);
    return stage.p;
}

void upload_padded_matrix(jues_ctx* ctx, DBuf& dst, const double* host, int64_t n, int64_t d, int64_t np,
                          int64_t dp) {
    dst.alloc(ctx, (size_t)(np * dp));
    if (np != n || dp != d) dst.zero();
    if (n > 0 && d > 0)
        JUES_CUDA(cudaMemcpy2DAsync(dst.p, np * 8, host, n * 8, n * 8, d, cudaMemcpyHostToDevice, ctx->stream));
}

// copy sigma-planes [lo, lo+cnt) of a host n^4 tensor into a padded (np^3 x cnt) device block
static void upload_gao_planes(jues_ctx* ctx, double* dst, const double* host, int64_t n, int64_t np,
                              int64_t lo, int64_t cnt) {
    const int64_t plane = np * np * np;
    if (n == np) {
        JUES_CUDA(cudaMemcpyAsync(dst, host + lo * n * n * n, (size_t)(plane * cnt) * sizeof(double),
                                  cudaMemcpyHostToDevice, ctx->stream));
        return;
    }
    JUES_CUDA(cudaMemsetAsync(dst, 0, (size_t)(plane * cnt) * sizeof(double), ctx->stream));
    for (int64_t s = 0; s < cnt && lo + s < n; ++s) {
        cudaMemcpy3DParms p = {};
        p.srcPtr = make_cudaPitchedPtr((void*)(host + (lo + s) * n * n * n), n * 8, n, n);
        p.dstPtr = make_cudaPitchedPtr((void*)(dst + s * plane), np * 8, np, np);
        p.extent = make_cudaExtent(n * 8, n, n);
        p.kind = cudaMemcpyHostToDevice;
        JUES_CUDA(cudaMemcpy3DAsync(&p, ctx->stream));
    }
}

const double* HostGao::slab(jues_ctx* ctx, int64_t lo, int64_t cnt) {
    const int64_t plane = np * np * np;
    if ((int64_t)stage.n < plane * cnt) stage.alloc(ctx, (size_t)(plane * cnt));
    upload_gao_planes(ctx, stage.p, h, n, np, lo, cnt);
    if (!phys) return stage.p;
    if ((int64_t)stage2.n < plane * cnt) stage2.alloc(ctx, (size_t)(plane * cnt));
    Ten a(stage.p, np, np, np, cnt), b(stage2.p, np, np, np, cnt);
    permute_axpby(ctx, 1.0, a, "mnls", 0.0, b, "mlns");
    return stage2.p;
}

// ---- blocks of the second chemists' index in physicists' order -----------------------------------------
void DeviceGao::phys_block(jues_ctx* ctx, int64_t lo, int64_t cnt, double* dst) {
    // view in[mu, n, lam, sig] = p[mu + np*(lo+n) + np^2*lam + np^3*sig]  ->  dst[mu, lam, n, sig]
    Ten in(const_cast<double*>(p) + lo * np, np, cnt, np, np), outv(dst, np, np, cnt, np);
    const int64_t st[4] = {1, np, np * np, np * np * np};
    permute_axpby_strided(ctx, 1.0, in, st, "mnls", 0.0, outv, "mlns");
}

void SynthGao::phys_block(jues_ctx* ctx, int64_t lo, int64_t cnt, double* dst) {
    synth_eri_block(ctx, dst, n, np, lo, cnt, 0, np, seed, scale, true);
}

void HostGao::phys_block(jues_ctx* ctx, int64_t lo, int64_t cnt, double* dst) {
    // host block h[mu, lo+n, lam, sig] -> device chem-ordered block [np, cnt, np, np] (zero padded), then
    // one device permutation into [mu, lam, n, sig].  For every (lam, sig) the host run (mu, n) is contiguous.
    const size_t need = (size_t)(np * cnt * np * np);
    if (stage.n < need) stage.alloc(ctx, need);
    const int64_t c_log = std::max<int64_t>(0, std::min(cnt, n - lo));   // rows that exist on the host
    if (n != np || c_log != cnt) JUES_CUDA(cudaMemsetAsync(stage.p, 0, need * sizeof(double), ctx->stream));
    if (c_log > 0) {
        if (n == np) {
            // rows of n*c_log doubles, one per (lam, sig): pitch n^2 on the host, n*cnt on the device
            JUES_CUDA(cudaMemcpy2DAsync(stage.p, (size_t)(np * cnt) * 8, h + lo * n, (size_t)(n * n) * 8,
                                        (size_t)(n * c_log) * 8, (size_t)(n * n), cudaMemcpyHostToDevice,
                                        ctx->stream));
        } else {
            // padded mu: one 3-D copy per sigma plane (x = mu, y = n, z = lam)
            for (int64_t sg = 0; sg < n; ++sg) {
                cudaMemcpy3DParms q = {};
                q.srcPtr = make_cudaPitchedPtr((void*)(h + sg * n * n * n), n * 8, n, n);
                q.srcPos = make_cudaPos(0, (size_t)lo, 0);
                q.dstPtr = make_cudaPitchedPtr((void*)(stage.p + sg * np * cnt * np), np * 8, np, cnt);
                q.extent = make_cudaExtent(n * 8, c_log, n);
                q.kind = cudaMemcpyHostToDevice;
                JUES_CUDA(cudaMemcpy3DAsync(&q, ctx->stream));
            }
        }
    }
    Ten in(stage.p, np, cnt, np, np), outv(dst, np, np, cnt, np);
    permute_axpby(ctx, 1.0, in, "mnls", 0.0, outv, "mlns");
}

void upload_padded_gao(jues_ctx* ctx, double* dst, const double* host, int64_t n, int64_t np) {
    if (!dst) return;
    upload_gao_planes(ctx, dst, host, n, np, 0, np);
}

namespace {

double quarter_flops(const int64_t e[4], int axis, int64_t d) {
    return 2.0 * (double)e[0] * (double)e[1] * (double)e[2] * (double)e[3] * (double)d;
}

// choose the contraction order (a permutation of axes 0..3) with the fewest flops
void best_order(int64_t np, const int64_t dp[4], bool ref_order, bool streamed, int order[4], double* flops) {
    int perm[4] = {0, 1, 2, 3};
    double best = 1e300;
    int best_perm[4] = {3, 2, 1, 0};
    if (ref_order) {
        int64_t e[4] = {np, np, np, np};
        double f = 0;
        for (int s = 0; s < 4; ++s) { f += quarter_flops(e, best_perm[s], dp[best_perm[s]]); e[best_perm[s]] = dp[best_perm[s]]; }
        best = f;
    } else {
        do {
            if (streamed && perm[0] != 3) continue;
            int64_t e[4] = {np, np, np, np};
            double f = 0;
            for (int s = 0; s < 4; ++s) { f += quarter_flops(e, perm[s], dp[perm[s]]); e[perm[s]] = dp[perm[s]]; }
            // tie-break towards the reference order (last axis first)
            if (f < best * (1.0 - 1e-12)) { best = f; std::copy(perm, perm + 4, best_perm); }
        } while (std::next_permutation(perm, perm + 4));
    }
    std::copy(best_perm, best_perm + 4, order);
    if (flops) *flops = best;
}

// one quarter transform of a dense device tensor `in` with extents e[4] along `axis`
void quarter(jues_ctx* ctx, const double* in, const int64_t e[4], int axis, const double* Cm, int64_t np,
             int64_t d, double* out, double beta = 0.0, int64_t k_lo = 0, int64_t k_cnt = -1) {
    int64_t pre = 1, post = 1;
    for (int q = 0; q < axis; ++q) pre *= e[q];
    for (int q = axis + 1; q < 4; ++q) post *= e[q];
    const int64_t nk = k_cnt < 0 ? e[axis] : k_cnt;
    GemmCall g;
    g.K = nk;
    g.beta = beta;
    if (axis == 0) {
        g.transA = true; g.transB = false;
        g.M = d; g.N = post;
        g.A = Cm + k_lo; g.lda = np;
        g.B = in; g.ldb = e[0];
        g.C = out; g.ldc = d;
    } else {
        g.transA = false; g.transB = false;
        g.M = pre; g.N = d;
        g.A = in; g.lda = pre; g.strideA = pre * e[axis];
        g.B = Cm + k_lo; g.ldb = np; g.strideB = 0;
        g.C = out; g.ldc = pre; g.strideC = pre * d;
        g.batch = post;
    }
    dgemm(ctx, g);
}

}  // namespace

double tei_transform_flops(int64_t np, const int64_t dp[4], bool reference_order, bool streamed) {
    int order[4];
    double f;
    best_order(np, dp, reference_order, streamed, order, &f);
    return f;
}

void tei_transform_dev(jues_ctx* ctx, GaoSource& gao, const double* const Cm[4], const int64_t dp[4],
                       double* out, bool reference_order, TransformWorkspace* ws) {
    const int64_t np = gao.np;
    int order[4];
    best_order(np, dp, reference_order, !gao.resident(), order, nullptr);
    int64_t e[4] = {np, np, np, np};
    TransformWorkspace local_ws;
    TransformWorkspace* w = ws ? ws : &local_ws;   // ping/pong buffers re-used by successive transforms
    const double* src = gao.base();
    for (int s = 0; s < 4; ++s) {
        const int ax = order[s];
        int64_t e2[4] = {e[0], e[1], e[2], e[3]};
        e2[ax] = dp[ax];
        const size_t nout = (size_t)(e2[0] * e2[1] * e2[2] * e2[3]);
        double* dst;
        if (s == 3) dst = out;
        else {
            DBuf& b = w->buf[s & 1];
            if (b.n < nout) b.alloc(ctx, nout);
            dst = b.p;
        }
        if (s == 0 && !gao.resident()) {
            // stream sigma slabs:  tmp[mu nu lam, b] += g[mu nu lam, slab] * C4[slab, b]
            JUES_REQUIRE(ax == 3, "internal: streamed transform must contract the last index first");
            const int64_t plane = np * np * np;
            // Every slab GEMM re-reads and re-writes the whole N^3 x d4 accumulator, so K (= slab
            // thickness) should be >= ~64 for the pass to stay compute-bound (2*cnt/16 flop per byte
            // of accumulator traffic vs a ridge of ~5.7 flop/B); spend up to 20 % of the free HBM.
            size_t free_b = 0, total_b = 0;
            cudaMemGetInfo(&free_b, &total_b);
            free_b += ctx->big_cached_bytes;   // cached blocks are ours to re-use
            int64_t cnt = (int64_t)(std::min(0.20 * (double)free_b, 24.0e9) / 8.0 / (double)plane);
            cnt = std::max<int64_t>(2, std::min<int64_t>(np, std::min<int64_t>(cnt, 128)));
            if (getenv("JUES_B200_FORCE_STREAM")) cnt = std::min<int64_t>(cnt, 6);  // testing: several slabs
            cnt &= ~int64_t(1);
            const int64_t s_begin = gao.sig_lo, s_end = gao.sig_hi < 0 ? np : gao.sig_hi;
            if (s_begin >= s_end) JUES_CUDA(cudaMemsetAsync(dst, 0, nout * sizeof(double), ctx->stream));
            if (gao.has_block3() && s_begin < s_end) {
                // blocks of the third index: out[(mu,nu,lam in block), b] = g[(mu,nu,lam), sigma] C4[sigma, b]
                // with the whole sigma range as K -- disjoint row blocks of the output, no accumulation
                for (int64_t lo = 0; lo < np; lo += cnt) {
                    const int64_t c = std::min(cnt, np - lo);
                    const double* blk;
                    {
                        TraceTimer tt(ctx, "tei.slab");
                        blk = gao.block3(ctx, lo, c);
                    }
                    TraceTimer tq(ctx, "tei.q1");
                    GemmCall g;
                    g.M = np * np * c; g.N = dp[3]; g.K = s_end - s_begin;
                    g.A = blk + s_begin * (np * np * c); g.lda = np * np * c;
                    g.B = Cm[3] + s_begin; g.ldb = np;
                    g.C = dst + lo * np * np; g.ldc = plane;
                    dgemm(ctx, g);
                }
            } else
            for (int64_t lo = s_begin; lo < s_end; lo += cnt) {
                const int64_t c = std::min(cnt, s_end - lo);
                const double* sl;
                {
                    TraceTimer tt(ctx, "tei.slab");
                    sl = gao.slab(ctx, lo, c);
                }
                int64_t es[4] = {np, np, np, c};
                TraceTimer tq(ctx, "tei.q1");
                quarter(ctx, sl, es, 3, Cm[3], np, dp[3], dst, lo == s_begin ? 0.0 : 1.0, lo, c);
            }
        } else {
            TraceTimer tq(ctx, s == 0 ? "tei.q1" : s == 1 ? "tei.q2" : s == 2 ? "tei.q3" : "tei.q4");
            quarter(ctx, src, e, ax, Cm[ax], np, dp[ax], dst);
        }
        if (g_probe) g_probe(g_probe_user, ctx, s, dst, nout);
        if (s < 3) src = dst;
        e[ax] = dp[ax];
    }
}

// ---------------------------------------------------------------------------------------------
// sharded one-pass transform (see transform.h)
// ---------------------------------------------------------------------------------------------
void ao_share(const jues_ctx* ctx, int64_t np, int rank, int64_t* lo, int64_t* cnt) {
    const int64_t per = (np + ctx->nranks - 1) / ctx->nranks;
    *lo = std::min<int64_t>(np, per * rank);
    *cnt = std::min<int64_t>(np, per * (rank + 1)) - *lo;
}

void tei_transform_sharded(jues_ctx* ctx, GaoSource& gao, const double* Cp, int64_t dP, const double* Cq,
                           int64_t dQ, const double* Cr, int64_t dR, const double* Cs_all,
                           const std::vector<int64_t>& s_counts, DBuf& out) {
    const int P = ctx->nranks, me = ctx->rank;
    const int64_t np = gao.np;
    JUES_REQUIRE((int)s_counts.size() == P, "sharded transform: one column count per rank");
    JUES_REQUIRE((dP & 1) == 0 && dP > 0 && dQ > 0 && dR > 0, "sharded transform: dP must be even");
    int64_t dS = 0, ns_max = 0;
    std::vector<int64_t> s_off(P);
    for (int d = 0; d < P; ++d) { s_off[d] = dS; dS += s_counts[d]; ns_max = std::max(ns_max, s_counts[d]); }
    const int64_t ns = s_counts[me];
    int64_t n0, nb;
    ao_share(ctx, np, me, &n0, &nb);
    const int64_t pq = dP * dQ;
    const int64_t per = (np + P - 1) / P;                            // the largest share of nu

    // Sub-block thickness: the same on every rank (the exchange rounds must line up), so it is a function
    // of the shapes only.  Per plane of nu: the AO block and its staging twin for host sources (2 np^3),
    // the two intermediates, and -- with several ranks -- two send and two receive buffers.
    const double per_plane = 8.0 * ((double)np * np * np * 2.0 + (double)dP * np * np + (double)pq * np +
                                    (P > 1 ? 2.0 * (double)pq * dS + 2.0 * (double)pq * ns_max * P : 0.0));
    int64_t cnt = (int64_t)(12.0e9 / per_plane);
    cnt = std::max<int64_t>(1, std::min<int64_t>(cnt, per));
    if (getenv("JUES_B200_FORCE_STREAM")) cnt = std::min<int64_t>(cnt, 3);   // testing: several sub-blocks
    const int64_t rounds = (per + cnt - 1) / cnt;

    DBuf yfull(ctx, (size_t)std::max<int64_t>(1, pq * np * ns));     // [p, q, nu (all), s in S_me]
    {
        DBuf ao(ctx, (size_t)(np * np * cnt * np)), x1(ctx, (size_t)(dP * np * cnt * np)),
            x2(ctx, (size_t)(pq * cnt * np));
        DBuf ysend[2], yrecv[2];
        cudaEvent_t ready[2] = {nullptr, nullptr}, done[2] = {nullptr, nullptr};
        struct EvGuard {
            jues_ctx* ctx; cudaEvent_t* a; cudaEvent_t* b;
            ~EvGuard() {
                // on an error path work may still be queued on the second stream: drain it before the buffers go
                if (std::uncaught_exceptions() > 0) cudaStreamSynchronize(ctx->comm_stream);
                for (int k = 0; k < 2; ++k) { if (a[k]) cudaEventDestroy(a[k]); if (b[k]) cudaEventDestroy(b[k]); }
            }
        } guard{ctx, ready, done};
        if (P > 1) {
            for (int k = 0; k < 2; ++k) {
                ysend[k].alloc(ctx, (size_t)(pq * cnt * dS));
                yrecv[k].alloc(ctx, (size_t)(pq * cnt * ns * P));
                JUES_CUDA(cudaEventCreateWithFlags(&ready[k], cudaEventDisableTiming));
                JUES_CUDA(cudaEventCreateWithFlags(&done[k], cudaEventDisableTiming));
            }
            // the second stream starts behind everything already queued on the first (allocations above)
            JUES_CUDA(cudaEventRecord(ready[0], ctx->stream));
            JUES_CUDA(cudaStreamWaitEvent(ctx->comm_stream, ready[0], 0));
        }
        for (int64_t k = 0; k < rounds; ++k) {
            const int64_t lo = k * cnt;
            const int64_t c = std::max<int64_t>(0, std::min(cnt, nb - lo));      // this rank's planes in round k
            const int b = (int)(k & 1);
            // ---- steps 1 + 2: three local quarters of this sub-block ---------------------------------
            if (P > 1 && k >= 2) JUES_CUDA(cudaStreamWaitEvent(ctx->stream, done[b], 0));   // send buffer free again
            if (c > 0) {
                {
                    TraceTimer tt(ctx, "tei.block");
                    gao.phys_block(ctx, n0 + lo, c, ao.p);
                }
                {   // mu -> p:  x1[p, (lam, n, sig)] = Cp^T (dP x np) * ao (np x np*c*np)
                    TraceTimer tq(ctx, "tei.q1");
                    GemmCall g;
                    g.transA = true;
                    g.M = dP; g.N = np * c * np; g.K = np;
                    g.A = Cp; g.lda = np;
                    g.B = ao.p; g.ldb = np;
                    g.C = x1.p; g.ldc = dP;
                    dgemm(ctx, g);
                }
                {   // lam -> q:  x2[p, q, (n, sig)] = x1[p, lam, (n, sig)] * Cq (np x dQ), batch over (n, sig)
                    TraceTimer tq(ctx, "tei.q2");
                    GemmCall g;
                    g.M = dP; g.N = dQ; g.K = np; g.batch = c * np;
                    g.A = x1.p; g.lda = dP; g.strideA = dP * np;
                    g.B = Cq; g.ldb = np; g.strideB = 0;
                    g.C = x2.p; g.ldc = dP; g.strideC = pq;
                    dgemm(ctx, g);
                }
                {   // sig -> s:  y[(p, q, n), s] = x2[(p, q, n), sig] * Cs_all (np x dS), all ranks' columns;
                    // with one rank straight into its place in yfull
                    TraceTimer tq(ctx, "tei.q3");
                    GemmCall g;
                    g.M = pq * c; g.N = dS; g.K = np;
                    g.A = x2.p; g.lda = pq * c;
                    g.B = Cs_all; g.ldb = np;
                    if (P == 1) { g.C = yfull.p + pq * lo; g.ldc = pq * np; }
                    else { g.C = ysend[b].p; g.ldc = pq * c; }
                    dgemm(ctx, g);
                }
            }
            if (P == 1) continue;
            // ---- step 3: (p, q, planes of round k, s in S_d) -> rank d, on the second stream so that it
            //      overlaps the quarters of the next sub-block --------------------------------------------
            JUES_CUDA(cudaEventRecord(ready[b], ctx->stream));
            {
                StreamScope sc(ctx, ctx->comm_stream);
                JUES_CUDA(cudaStreamWaitEvent(ctx->stream, ready[b], 0));
                TraceTimer tx(ctx, "tei.exchange");
                std::vector<size_t> so(P), scn(P), ro(P), rc(P);
                std::vector<int64_t> c_of(P), lo_of(P);
                size_t racc = 0;
                for (int d = 0; d < P; ++d) {
                    int64_t lo_d, nb_d;
                    ao_share(ctx, np, d, &lo_d, &nb_d);
                    lo_of[d] = lo_d;
                    c_of[d] = std::max<int64_t>(0, std::min(cnt, nb_d - lo));
                    so[d] = (size_t)(pq * c * s_off[d]);
                    scn[d] = (size_t)(pq * c * s_counts[d]);
                    ro[d] = racc;
                    rc[d] = (size_t)(pq * c_of[d] * ns);
                    racc += rc[d];
                }
                all_to_all_v(ctx, ysend[b].p, so.data(), scn.data(), yrecv[b].p, ro.data(), rc.data());
                for (int d = 0; d < P; ++d) {
                    if (c_of[d] == 0 || ns == 0) continue;
                    const int64_t sd[4] = {pq * c_of[d], ns, 1, 1}, dd[4] = {pq * np, ns, 1, 1};
                    block_copy(ctx, yrecv[b].p + ro[d], sd, yfull.p + pq * (lo_of[d] + lo), dd, sd);
                }
                JUES_CUDA(cudaEventRecord(done[b], ctx->stream));
            }
        }
        if (P > 1) {
            JUES_CUDA(cudaStreamWaitEvent(ctx->stream, done[0], 0));
            JUES_CUDA(cudaStreamWaitEvent(ctx->stream, done[1], 0));
            // the buffers of this scope are released in stream order behind these waits
        }
    }
    // ---- step 4: nu -> r over the complete nu range ------------------------------------------------------
    out.alloc(ctx, (size_t)std::max<int64_t>(1, pq * dR * ns));
    if (ns > 0) {
        TraceTimer tq(ctx, "tei.q4");
        GemmCall g;
        g.M = pq; g.N = dR; g.K = np; g.batch = ns;
        g.A = yfull.p; g.lda = pq; g.strideA = pq * np;
        g.B = Cr; g.ldb = np; g.strideB = 0;
        g.C = out.p; g.ldc = pq; g.strideC = pq * dR;
        dgemm(ctx, g);
    }
}

}  // namespace jues
