// extern "C" entry points for the transformation, RMP2, RCCD, RCCSD and the device-resident
// rank-4 tensor handle (include/jues_b200.h).
#include "api_util.h"
#include "cc.h"
#include "dist.h"

#include <memory>

using namespace jues;

namespace {

void begin_call(jues_ctx* ctx) {
    resolve_timers(ctx);
    ctx->timings.clear();
    ctx->alloc_host_s = 0.0;
    ctx->alloc_calls = 0;
    ctx->stats.reset();
    ctx->bytes_peak = ctx->bytes_allocated;
}

size_t free_device_bytes() {
    size_t f = 0, t = 0;
    cudaMemGetInfo(&f, &t);
    return f;
}

// A GaoSource for a host tensor: resident copy when it fits comfortably, streamed otherwise.
struct HostGaoHolder {
    DBuf dense;
    std::unique_ptr<GaoSource> src;
};

// stream == true: the consumer is the sharded one-pass transform, which reads each AO block exactly once
// (RMP2, RCCD, RCCSD, mRCCD): blocks are uploaded straight from the caller's array, 1/P of it per rank.
// stream == false: the AO tensor is read several times (tei_transform in arbitrary contraction order,
// get_fock + coupled cluster): keep a device copy when it fits comfortably.
void make_host_gao(jues_ctx* ctx, HostGaoHolder& h, const double* gao, int64_t nao, bool stream) {
    JUES_REQUIRE(gao != nullptr && nao > 0, "null or empty gao");
    const int64_t np = round_up(nao, 2);
    const double bytes = (double)np * np * np * np * 8.0;
    const bool force_stream = getenv("JUES_B200_FORCE_STREAM") != nullptr;  // testing hook
    if (stream || force_stream) {
        h.src.reset(new HostGao(gao, nao, np));
        return;
    }
    flush_big_cache(ctx);                       // make room for the resident copy
    const double avail = (double)free_device_bytes() + 0.0;
    if (bytes < 0.45 * avail) {
        Timer t(ctx, "h2d.gao");
        h.dense.alloc(ctx, (size_t)(np * np * np * np));
        upload_padded_gao(ctx, h.dense.p, gao, nao, np);
        h.src.reset(new DeviceGao(h.dense.p, nao, np));
    } else {
        h.src.reset(new HostGao(gao, nao, np));
    }
}

void check_t4_is_gao(const jues_t4* g) {
    JUES_REQUIRE(g && (g->p || g->virtual_synth), "null tensor handle");
    JUES_REQUIRE(g->d[0] == g->d[1] && g->d[1] == g->d[2] && g->d[2] == g->d[3],
                 "gao handle must be (nao,nao,nao,nao)");
}

// GaoSource of a tensor handle: dense storage or on-demand synthetic slabs
struct T4Source {
    std::unique_ptr<GaoSource> src;
    explicit T4Source(const jues_t4* g) {
        if (g->virtual_synth) src.reset(new SynthGao(g->d[0], g->dp[0], g->seed, g->scale));
        else src.reset(new DeviceGao(g->p, g->d[0], g->dp[0]));
    }
};

// download a padded device tensor (dp) into an unpadded host array (d)
void download_block(jues_ctx* ctx, const double* dev, const int64_t dp[4], double* host, const int64_t d[4]) {
    const size_t n = (size_t)(d[0] * d[1] * d[2] * d[3]);
    if (!n) return;
    if (d[0] == dp[0] && d[1] == dp[1] && d[2] == dp[2]) {
        JUES_CUDA(cudaMemcpyAsync(host, dev, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    } else {
        DBuf tmp(ctx, n);
        block_copy(ctx, dev, dp, tmp.p, d, d);
        JUES_CUDA(cudaMemcpyAsync(host, tmp.p, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    }
    JUES_CUDA(cudaStreamSynchronize(ctx->stream));
}

void transform_common(jues_ctx* ctx, GaoSource& src, const double* const Ch[4], const int64_t d[4],
                      int phys_order, DBuf& out, int64_t dp_out[4]) {
    const int64_t nao = src.n, np = src.np;
    DBuf Cd[4];
    const double* Cm[4];
    int64_t dp[4];
    for (int q = 0; q < 4; ++q) {
        JUES_REQUIRE(Ch[q] != nullptr && d[q] > 0, "null or empty coefficient matrix");
        // multi-GPU: the LAST MO index is split into equal even slabs, one per rank
        dp[q] = round_up(d[q], q == 3 ? 2 * (int64_t)ctx->nranks : 2);
        upload_padded_matrix(ctx, Cd[q], Ch[q], nao, d[q], np, dp[q]);
        Cm[q] = Cd[q].p;
    }
    const size_t n = (size_t)(dp[0] * dp[1] * dp[2] * dp[3]);
    DBuf chem(ctx, n);
    {
        Timer t(ctx, "tei.transform");
        int64_t b0, vs;
        slab_of(ctx, dp[3], &b0, &vs);
        const double* Cs[4] = {Cm[0], Cm[1], Cm[2], Cm[3] + b0 * np};
        const int64_t ds[4] = {dp[0], dp[1], dp[2], vs};
        const size_t per_rank = (size_t)(dp[0] * dp[1] * dp[2] * vs);
        tei_transform_dev(ctx, src, Cs, ds, chem.p + (size_t)ctx->rank * per_rank);
        all_gather_inplace(ctx, chem.p, per_rank);   // every rank ends with the whole tensor
    }
    if (phys_order) {
        Timer t(ctx, "tei.permute");
        out.alloc(ctx, n);
        Ten a(chem.p, dp[0], dp[1], dp[2], dp[3]), b(out.p, dp[0], dp[2], dp[1], dp[3]);
        permute_axpby(ctx, 1.0, a, "iajb", 0.0, b, "ijab");
        dp_out[0] = dp[0]; dp_out[1] = dp[2]; dp_out[2] = dp[1]; dp_out[3] = dp[3];
    } else {
        out = std::move(chem);
        for (int q = 0; q < 4; ++q) dp_out[q] = dp[q];
    }
}

}  // namespace

extern "C" int jues_b200_set_amplitude_callback(jues_ctx* ctx, jues_b200_amp_cb cb, void* user) {
    if (!ctx) return JUES_B200_EINVAL;
    ctx->amp_cb = cb;
    ctx->amp_user = user;
    return JUES_B200_OK;
}

// ---------------------------------------------------------------------------------------------
// tei_transform
// ---------------------------------------------------------------------------------------------
extern "C" int jues_b200_tei_transform(jues_ctx* ctx, const double* gao, int64_t nao, const double* C1,
                                       int64_t d1, const double* C2, int64_t d2, const double* C3,
                                       int64_t d3, const double* C4, int64_t d4, int phys_order, double* out) {
    JUES_API_BEGIN(ctx)
    begin_call(ctx);
    JUES_REQUIRE(out != nullptr, "null output");
    Timer total(ctx, "total");
    HostGaoHolder h;
    make_host_gao(ctx, h, gao, nao, false);
    const double* Ch[4] = {C1, C2, C3, C4};
    const int64_t d[4] = {d1, d2, d3, d4};
    DBuf res;
    int64_t dp[4];
    transform_common(ctx, *h.src, Ch, d, phys_order, res, dp);
    const int64_t dl[4] = {d1, phys_order ? d3 : d2, phys_order ? d2 : d3, d4};
    Timer t(ctx, "d2h.out");
    download_block(ctx, res.p, dp, out, dl);
    JUES_API_END(ctx)
}

extern "C" int jues_b200_tei_transform_t4(jues_ctx* ctx, const jues_t4* gao, const double* C1, int64_t d1,
                                          const double* C2, int64_t d2, const double* C3, int64_t d3,
                                          const double* C4, int64_t d4, int phys_order, jues_t4** out) {
    JUES_API_BEGIN(ctx)
    begin_call(ctx);
    check_t4_is_gao(gao);
    JUES_REQUIRE(out != nullptr, "null output");
    Timer total(ctx, "total");
    T4Source holder(gao);
    GaoSource& src = *holder.src;
    const double* Ch[4] = {C1, C2, C3, C4};
    const int64_t d[4] = {d1, d2, d3, d4};
    DBuf res;
    int64_t dp[4];
    transform_common(ctx, src, Ch, d, phys_order, res, dp);
    const int64_t dl[4] = {d1, phys_order ? d3 : d2, phys_order ? d2 : d3, d4};
    jues_t4* t = nullptr;
    int rc = jues_b200_t4_create(ctx, dl[0], dl[1], dl[2], dl[3], &t);
    if (rc) return rc;
    block_copy(ctx, res.p, dp, t->p, t->dp, dl);   // padded extents may differ (multi-GPU slab padding)
    JUES_CUDA(cudaStreamSynchronize(ctx->stream));
    *out = t;
    JUES_API_END(ctx)
}

// Determinism stress of the transform (test hook): the same transform is run reps+1 times inside one call;
// the output of every quarter of every repetition is compared on the device with the first run's.
// bad[4] / worst[4]: per quarter, repetitions that differ and the largest sum of squared differences.
namespace {
struct ProbeState {
    DBuf ref[4];
    DBuf res;     // per (rep, quarter): [count, first index, last index, max |diff|]
    int rep = -1, reps = 0;
};
__global__ void diffstat_kernel(const double* __restrict__ a, const double* __restrict__ b, size_t n,
                                double* __restrict__ out) {
    unsigned long long cnt = 0, first = ~0ull, last = 0;
    double mx = 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const double d = fabs(a[i] - b[i]);
        if (d != 0.0 || a[i] != a[i] || b[i] != b[i]) {
            ++cnt;
            if (i < first) first = i;
            if (i > last) last = i;
            if (d > mx) mx = d;
        }
    }
    if (cnt) {
        atomicAdd(reinterpret_cast<unsigned long long*>(out), cnt);
        atomicMin(reinterpret_cast<unsigned long long*>(out + 1), first);
        atomicMax(reinterpret_cast<unsigned long long*>(out + 2), last);
        atomicMax(reinterpret_cast<unsigned long long*>(out + 3), (unsigned long long)__double_as_longlong(mx));
    }
}
void stress_probe(void* user, jues_ctx* ctx, int step, const double* out, size_t n) {
    ProbeState* st = static_cast<ProbeState*>(user);
    if (st->rep < 0) {
        st->ref[step].alloc(ctx, n);
        cudaMemcpyAsync(st->ref[step].p, out, n * 8, cudaMemcpyDeviceToDevice, ctx->stream);
    } else {
        diffstat_kernel<<<148 * 8, 256, 0, ctx->stream>>>(st->ref[step].p, out, n, st->res.p + 4 * (4 * st->rep + step));
    }
}
}  // namespace

extern "C" int jues_b200_transform_stress(jues_ctx* ctx, const jues_t4* gao, const double* C1, int64_t d1,
                                          const double* C2, int64_t d2, const double* C3, int64_t d3,
                                          const double* C4, int64_t d4, int reps, double* stats) {
    JUES_API_BEGIN(ctx)
    begin_call(ctx);
    check_t4_is_gao(gao);
    JUES_REQUIRE(reps > 0 && reps <= 1024 && stats, "bad arguments");
    T4Source holder(gao);
    const double* Ch[4] = {C1, C2, C3, C4};
    const int64_t d[4] = {d1, d2, d3, d4};
    ProbeState st;
    st.reps = reps;
    st.res.alloc(ctx, (size_t)reps * 16);
    std::vector<unsigned long long> init((size_t)reps * 16, 0ull);
    for (int k = 0; k < reps * 4; ++k) init[(size_t)4 * k + 1] = ~0ull;
    JUES_CUDA(cudaMemcpyAsync(st.res.p, init.data(), init.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
    JUES_CUDA(cudaStreamSynchronize(ctx->stream));
    struct Guard { ~Guard() { set_quarter_probe(nullptr, nullptr); } } guard;
    set_quarter_probe(stress_probe, &st);
    for (int r = -1; r < reps; ++r) {
        st.rep = r;
        DBuf res;
        int64_t dp[4];
        transform_common(ctx, *holder.src, Ch, d, 0, res, dp);
    }
    std::vector<unsigned long long> h((size_t)reps * 16);
    JUES_CUDA(cudaMemcpyAsync(h.data(), st.res.p, h.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
    JUES_CUDA(cudaStreamSynchronize(ctx->stream));
    for (size_t k = 0; k < h.size(); ++k) {
        if (k % 4 == 3) { double x; memcpy(&x, &h[k], 8); stats[k] = x; }
        else stats[k] = (h[k] == ~0ull) ? -1.0 : (double)h[k];
    }
    JUES_API_END(ctx)
}

// ---------------------------------------------------------------------------------------------
// RMP2
// ---------------------------------------------------------------------------------------------
extern "C" int jues_b200_rmp2(jues_ctx* ctx, const double* gao, int64_t nao, const double* Cao, int64_t nocc,
                              const double* Cav, int64_t nvir, const double* eps, double* e_mp2) {
    if (is_group_call(ctx))
        return group_run(ctx, [&](jues_ctx* m, bool lead) {
            double e_;
            return jues_b200_rmp2(m, gao, nao, Cao, nocc, Cav, nvir, eps, lead ? e_mp2 : &e_);
        });
    JUES_API_BEGIN(ctx)
    begin_call(ctx);
    JUES_REQUIRE(e_mp2 != nullptr, "null output");
    Timer total(ctx, "total");
    Problem P;
    setup_problem(ctx, P, nao, Cao, nocc, Cav, nvir, eps);
    HostGaoHolder h;
    make_host_gao(ctx, h, gao, nao, true);
    *e_mp2 = rmp2_dev(ctx, P, *h.src);
    JUES_API_END(ctx)
}

extern "C" int jues_b200_rmp2_t4(jues_ctx* ctx, const jues_t4* gao, const double* Cao, int64_t nocc,
                                 const double* Cav, int64_t nvir, const double* eps, double* e_mp2) {
    if (is_group_call(ctx) && gao && !gao->virtual_synth) {
        ctx->last_error = "a dense device tensor lives on one GPU: pass the host array or a generated tensor to a multi-GPU group";
        return JUES_B200_ESTATE;
    }
    if (is_group_call(ctx))
        return group_run(ctx, [&](jues_ctx* m, bool lead) {
            double e_;
            return jues_b200_rmp2_t4(m, gao, Cao, nocc, Cav, nvir, eps, lead ? e_mp2 : &e_);
        });
    JUES_API_BEGIN(ctx)
    begin_call(ctx);
    check_t4_is_gao(gao);
    JUES_REQUIRE(e_mp2 != nullptr, "null output");
    Timer total(ctx, "total");
    Problem P;
    setup_problem(ctx, P, gao->d[0], Cao, nocc, Cav, nvir, eps);
    T4Source holder(gao);
    GaoSource& src = *holder.src;
    *e_mp2 = rmp2_dev(ctx, P, src);
    JUES_API_END(ctx)
}

// ---------------------------------------------------------------------------------------------
// RCCD / RCCSD
// ---------------------------------------------------------------------------------------------
namespace {
void run_cc(jues_ctx* ctx, GaoSource& src, const double* Cao, int64_t nocc, const double* Cav, int64_t nvir,
            const double* eps, bool singles, int maxit, int guess_mode, double* e, double* e_hist,
            double* T1_out, double* T2_out) {
    JUES_REQUIRE(e != nullptr, "null energy output");
    Problem P;
    setup_problem(ctx, P, src.n, Cao, nocc, Cav, nvir, eps);
    CCResult r = cc_dev(ctx, P, src, singles, maxit, guess_mode, T1_out, T2_out, ctx->amp_cb, ctx->amp_user);
    *e = r.energy;
    ctx->timings.emplace_back("alloc.host_ms", (float)(ctx->alloc_host_s * 1e3));
    ctx->timings.emplace_back("alloc.calls", (float)ctx->alloc_calls);
    if (e_hist)
        for (int k = 0; k <= maxit; ++k) e_hist[k] = r.e_hist[k];
}
}  // namespace

extern "C" int jues_b200_rccd(jues_ctx* ctx, const double* gao, int64_t nao, const double* Cao, int64_t nocc,
                              const double* Cav, int64_t nvir, const double* eps, int maxit, int guess_mode,
                              double* e_ccd, double* e_hist, double* T2_out) {
    if (is_group_call(ctx))
        return group_run(ctx, [&](jues_ctx* m, bool lead) {
            double e_;
            return jues_b200_rccd(m, gao, nao, Cao, nocc, Cav, nvir, eps, maxit, guess_mode, lead ? e_ccd : &e_,
                                  lead ? e_hist : nullptr, lead ? T2_out : nullptr);
        });
    JUES_API_BEGIN(ctx)
    begin_call(ctx);
    JUES_REQUIRE(guess_mode == 0 || guess_mode == 1, "guess_mode must be 0 (reference) or 1 (MP2)");
    Timer total(ctx, "total");
    HostGaoHolder h;
    make_host_gao(ctx, h, gao, nao, true);
    run_cc(ctx, *h.src, Cao, nocc, Cav, nvir, eps, false, maxit, guess_mode, e_ccd, e_hist, nullptr, T2_out);
    JUES_API_END(ctx)
}

extern "C" int jues_b200_rccd_t4(jues_ctx* ctx, const jues_t4* gao, const double* Cao, int64_t nocc,
                                 const double* Cav, int64_t nvir, const double* eps, int maxit,
                                 int guess_mode, double* e_ccd, double* e_hist, double* T2_out) {
    if (is_group_call(ctx) && gao && !gao->virtual_synth) {
        ctx->last_error = "a dense device tensor lives on one GPU: pass the host array or a generated tensor to a multi-GPU group";
        return JUES_B200_ESTATE;
    }
    if (is_group_call(ctx))
        return group_run(ctx, [&](jues_ctx* m, bool lead) {
            double e_;
            return jues_b200_rccd_t4(m, gao, Cao, nocc, Cav, nvir, eps, maxit, guess_mode, lead ? e_ccd : &e_,
                                     lead ? e_hist : nullptr, lead ? T2_out : nullptr);
        });
    JUES_API_BEGIN(ctx)
    begin_call(ctx);
    check_t4_is_gao(gao);
    JUES_REQUIRE(guess_mode == 0 || guess_mode == 1, "guess_mode must be 0 (reference) or 1 (MP2)");
    Timer total(ctx, "total");
    T4Source holder(gao);
    GaoSource& src = *holder.src;
    run_cc(ctx, src, Cao, nocc, Cav, nvir, eps, false, maxit, guess_mode, e_ccd, e_hist, nullptr, T2_out);
    JUES_API_END(ctx)
}

extern "C" int jues_b200_rccsd(jues_ctx* ctx, const double* gao, int64_t nao, const double* Cao, int64_t nocc,
                               const double* Cav, int64_t nvir, const double* eps, int maxit, double* e_ccsd,
                               double* e_hist, double* T1_out, double* T2_out) {
    if (is_group_call(ctx))
        return group_run(ctx, [&](jues_ctx* m, bool lead) {
            double e_;
            return jues_b200_rccsd(m, gao, nao, Cao, nocc, Cav, nvir, eps, maxit, lead ? e_ccsd : &e_,
                                   lead ? e_hist : nullptr, lead ? T1_out : nullptr, lead ? T2_out : nullptr);
        });
    JUES_API_BEGIN(ctx)
    begin_call(ctx);
    Timer total(ctx, "total");
    HostGaoHolder h;
    make_host_gao(ctx, h, gao, nao, true);
    run_cc(ctx, *h.src, Cao, nocc, Cav, nvir, eps, true, maxit, 1, e_ccsd, e_hist, T1_out, T2_out);
    JUES_API_END(ctx)
}

extern "C" int jues_b200_rccsd_t4(jues_ctx* ctx, const jues_t4* gao, const double* Cao, int64_t nocc,
                                  const double* Cav, int64_t nvir, const double* eps, int maxit,
                                  double* e_ccsd, double* e_hist, double* T1_out, double* T2_out) {
    if (is_group_call(ctx) && gao && !gao->virtual_synth) {
        ctx->last_error = "a dense device tensor lives on one GPU: pass the host array or a generated tensor to a multi-GPU group";
        return JUES_B200_ESTATE;
    }
    if (is_group_call(ctx))
        return group_run(ctx, [&](jues_ctx* m, bool lead) {
            double e_;
            return jues_b200_rccsd_t4(m, gao, Cao, nocc, Cav, nvir, eps, maxit, lead ? e_ccsd : &e_,
                                      lead ? e_hist : nullptr, lead ? T1_out : nullptr, lead ? T2_out : nullptr);
        });
    JUES_API_BEGIN(ctx)
    begin_call(ctx);
    check_t4_is_gao(gao);
    Timer total(ctx, "total");
    T4Source holder(gao);
    GaoSource& src = *holder.src;
    run_cc(ctx, src, Cao, nocc, Cav, nvir, eps, true, maxit, 1, e_ccsd, e_hist, T1_out, T2_out);
    JUES_API_END(ctx)
}

// ---------------------------------------------------------------------------------------------
// jues_t4: device-resident rank-4 tensor (DiskFourTensor replacement, DiskFourTensors.jl:5-95)
// ---------------------------------------------------------------------------------------------
extern "C" int jues_b200_t4_create(jues_ctx* ctx, int64_t d1, int64_t d2, int64_t d3, int64_t d4, jues_t4** out) {
    JUES_API_BEGIN(ctx)
    JUES_REQUIRE(out != nullptr, "null output");
    JUES_REQUIRE(d1 > 0 && d2 > 0 && d3 > 0 && d4 > 0, "extents must be positive");
    std::unique_ptr<jues_t4> t(new jues_t4());
    t->ctx = ctx;
    const int64_t d[4] = {d1, d2, d3, d4};
    size_t n = 1;
    for (int q = 0; q < 4; ++q) { t->d[q] = d[q]; t->dp[q] = round_up(d[q], 2); n *= (size_t)t->dp[q]; }
    t->bytes = (n * 8 + 255) & ~size_t(255);
    cudaError_t e = cudaMalloc(&t->p, t->bytes);
    if (e == cudaErrorMemoryAllocation) {
        cudaGetLastError();
        flush_big_cache(ctx);
        e = cudaMalloc(&t->p, t->bytes);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        char buf[160];
        snprintf(buf, sizeof buf, "device allocation of %.3f GB for a rank-4 tensor failed", t->bytes / 1e9);
        throw Error(JUES_B200_ENOMEM, buf);
    }
    JUES_CUDA(cudaMemsetAsync(t->p, 0, t->bytes, ctx->stream));
    JUES_CUDA(cudaStreamSynchronize(ctx->stream));
    *out = t.release();
    JUES_API_END(ctx)
}

extern "C" int jues_b200_t4_create_synth(jues_ctx* ctx, int64_t nao, uint64_t seed, double scale, jues_t4** out) {
    JUES_API_BEGIN(ctx)
    JUES_REQUIRE(out != nullptr && nao > 0, "bad arguments");
    std::unique_ptr<jues_t4> t(new jues_t4());
    t->ctx = ctx;
    for (int q = 0; q < 4; ++q) { t->d[q] = nao; t->dp[q] = round_up(nao, 2); }
    t->virtual_synth = true;
    t->seed = seed;
    t->scale = scale;
    *out = t.release();
    JUES_API_END(ctx)
}

extern "C" int jues_b200_get_comm_counters(jues_ctx* ctx, int64_t* collectives, double* bytes_received) {
    if (!ctx) return JUES_B200_EINVAL;
    if (collectives) *collectives = ctx->stats.collectives;
    if (bytes_received) *bytes_received = ctx->stats.collective_bytes;
    return JUES_B200_OK;
}

extern "C" int jues_b200_t4_destroy(jues_t4* t) {
    if (!t) return JUES_B200_EINVAL;
    if (t->ctx) cudaSetDevice(t->ctx->device);
    if (t->p) cudaFree(t->p);
    delete t;
    return JUES_B200_OK;
}

extern "C" int jues_b200_t4_dims(const jues_t4* t, int64_t dims_out[4]) {
    if (!t || !dims_out) return JUES_B200_EINVAL;
    for (int q = 0; q < 4; ++q) dims_out[q] = t->d[q];
    return JUES_B200_OK;
}

extern "C" int jues_b200_t4_fill(jues_t4* t, double value) {
    if (!t) return JUES_B200_EINVAL;
    jues_ctx* ctx = t->ctx;
    JUES_API_BEGIN(ctx)
    // pads stay zero: fill the logical block only
    const size_t n = (size_t)(t->dp[0] * t->dp[1] * t->dp[2] * t->dp[3]);
    if (t->d[0] == t->dp[0] && t->d[1] == t->dp[1] && t->d[2] == t->dp[2] && t->d[3] == t->dp[3]) {
        fill(ctx, t->p, n, value);
    } else {
        const size_t nl = (size_t)(t->d[0] * t->d[1] * t->d[2] * t->d[3]);
        DBuf tmp(ctx, nl);
        fill(ctx, tmp.p, nl, value);
        block_copy(ctx, tmp.p, t->d, t->p, t->dp, t->d);
    }
    JUES_CUDA(cudaStreamSynchronize(ctx->stream));
    JUES_API_END(ctx)
}

namespace {
void check_range(const jues_t4* t, const int64_t lo[4], const int64_t hi[4], int64_t ext[4]) {
    JUES_REQUIRE(t && (t->p || t->virtual_synth) && lo && hi, "null argument");
    for (int q = 0; q < 4; ++q) {
        JUES_REQUIRE(lo[q] >= 0 && hi[q] <= t->d[q] && lo[q] <= hi[q], "slice out of range");
        ext[q] = hi[q] - lo[q];
    }
}
}  // namespace

extern "C" int jues_b200_t4_set_slice(jues_t4* t, const int64_t lo[4], const int64_t hi[4], const double* host) {
    if (!t) return JUES_B200_EINVAL;
    jues_ctx* ctx = t->ctx;
    JUES_API_BEGIN(ctx)
    int64_t ext[4];
    check_range(t, lo, hi, ext);
    JUES_REQUIRE(host != nullptr, "null host buffer");
    if (t->virtual_synth) throw Error(JUES_B200_ESTATE, "a synthetic (generated) tensor is read-only");
    const size_t n = (size_t)(ext[0] * ext[1] * ext[2] * ext[3]);
    if (n) {
        DBuf tmp(ctx, n);
        JUES_CUDA(cudaMemcpyAsync(tmp.p, host, n * 8, cudaMemcpyHostToDevice, ctx->stream));
        double* dst = t->p + lo[0] + t->dp[0] * (lo[1] + t->dp[1] * (lo[2] + t->dp[2] * lo[3]));
        block_copy(ctx, tmp.p, ext, dst, t->dp, ext);
        JUES_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    JUES_API_END(ctx)
}

extern "C" int jues_b200_t4_get_slice(const jues_t4* t, const int64_t lo[4], const int64_t hi[4], double* host) {
    if (!t) return JUES_B200_EINVAL;
    jues_ctx* ctx = t->ctx;
    JUES_API_BEGIN(ctx)
    int64_t ext[4];
    check_range(t, lo, hi, ext);
    JUES_REQUIRE(host != nullptr, "null host buffer");
    const size_t n = (size_t)(ext[0] * ext[1] * ext[2] * ext[3]);
    if (n) {
        DBuf tmp(ctx, n), gen;
        const double* src;
        if (t->virtual_synth) {
            // generate the sigma planes the slice touches, then cut the block out of them
            const int64_t np = t->dp[0];
            gen.alloc(ctx, (size_t)(np * np * np * ext[3]));
            synth_eri_fill(ctx, gen.p, t->d[0], np, lo[3], ext[3], t->seed, t->scale);
            src = gen.p + lo[0] + np * (lo[1] + np * lo[2]);
        } else {
            src = t->p + lo[0] + t->dp[0] * (lo[1] + t->dp[1] * (lo[2] + t->dp[2] * lo[3]));
        }
        block_copy(ctx, src, t->dp, tmp.p, ext, ext);
        JUES_CUDA(cudaMemcpyAsync(host, tmp.p, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
        JUES_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    JUES_API_END(ctx)
}

extern "C" int jues_b200_t4_synth_eri(jues_t4* t, uint64_t seed, double scale) {
    if (!t) return JUES_B200_EINVAL;
    jues_ctx* ctx = t->ctx;
    JUES_API_BEGIN(ctx)
    check_t4_is_gao(t);
    if (t->virtual_synth) throw Error(JUES_B200_ESTATE, "a synthetic (generated) tensor is read-only");
    synth_eri_fill(ctx, t->p, t->d[0], t->dp[0], 0, t->dp[3], seed, scale);
    JUES_CUDA(cudaStreamSynchronize(ctx->stream));
    JUES_API_END(ctx)
}

// ---------------------------------------------------------------------------------------------
// get_fock, AutoRCCSD.do_rccsd, compute_pT
// ---------------------------------------------------------------------------------------------
#include "pt.h"

extern "C" int jues_b200_cc_default_options(jues_b200_cc_options* opt) {
    if (!opt) return JUES_B200_EINVAL;
    // CoupledCluster.defaults (CoupledCluster.jl:36-43)
    opt->cc_max_iter = 50;
    opt->cc_e_conv = 1e-10;
    opt->cc_max_rms = 1e-10;
    opt->do_pT = 0;
    opt->fcn = 0;
    opt->diis = 0;
    return JUES_B200_OK;
}

extern "C" int jues_b200_get_fock(jues_ctx* ctx, const double* gao, int64_t nao, const double* hao,
                                  const double* C, int64_t nmo, const double* Co, int64_t nocc, double* f_out) {
    JUES_API_BEGIN(ctx)
    begin_call(ctx);
    Timer total(ctx, "total");
    HostGaoHolder h;
    make_host_gao(ctx, h, gao, nao, false);
    Timer t(ctx, "fock.build");
    fock_dev(ctx, *h.src, hao, C, nmo, Co, nocc, f_out);
    JUES_API_END(ctx)
}

extern "C" int jues_b200_get_fock_t4(jues_ctx* ctx, const jues_t4* gao, const double* hao, const double* C,
                                     int64_t nmo, const double* Co, int64_t nocc, double* f_out) {
    JUES_API_BEGIN(ctx)
    begin_call(ctx);
    check_t4_is_gao(gao);
    Timer total(ctx, "total");
    T4Source holder(gao);
    Timer t(ctx, "fock.build");
    fock_dev(ctx, *holder.src, hao, C, nmo, Co, nocc, f_out);
    JUES_API_END(ctx)
}

namespace {
void run_auto(jues_ctx* ctx, GaoSource& src, const double* hao, const double* Ca, int64_t nmo, int64_t ndocc,
              const jues_b200_cc_options* opt_in, double* e_cc, double* e_pt, int* iterations, int* converged,
              double* e_hist, double* rms_hist, double* T1_out, double* T2_out) {
    jues_b200_cc_options opt;
    jues_b200_cc_default_options(&opt);
    if (opt_in) opt = *opt_in;
    JUES_REQUIRE(e_cc != nullptr, "null energy output");
    JUES_REQUIRE(hao && Ca, "null hao / Ca");
    JUES_REQUIRE(ndocc > 0 && nmo > ndocc, "need 0 < ndocc < nmo");
    JUES_REQUIRE(opt.fcn >= 0 && opt.fcn < ndocc, "fcn must satisfy 0 <= fcn < ndocc");
    JUES_REQUIRE(opt.cc_max_iter >= 0, "cc_max_iter must be non-negative");
    JUES_REQUIRE(!opt.do_pT || e_pt, "do_pT needs an e_pt output");
    const int64_t nao = src.n, fcn = opt.fcn;
    const int64_t no = ndocc - fcn, nv = nmo - ndocc;
    // Fock matrix from ALL doubly occupied orbitals (AutoRCCSD.jl:219; IntegralTransformation.jl:122-138)
    std::vector<double> f((size_t)(nmo * nmo));
    {
        Timer t(ctx, "fock.build");
        fock_dev(ctx, src, hao, Ca, nmo, Ca, ndocc, f.data());
    }
    // resolvents from the diagonal (:222-224,242-244), off-diagonal blocks (:227-231)
    std::vector<double> eps((size_t)(no + nv)), foo((size_t)(no * no)), fov((size_t)(no * nv)), fvv((size_t)(nv * nv));
    auto F = [&](int64_t p, int64_t q) { return p == q ? 0.0 : f[(size_t)(p + nmo * q)]; };
    for (int64_t i = 0; i < no; ++i) eps[i] = f[(size_t)((fcn + i) * (nmo + 1))];
    for (int64_t a = 0; a < nv; ++a) eps[no + a] = f[(size_t)((ndocc + a) * (nmo + 1))];
    for (int64_t k = 0; k < no; ++k)
        for (int64_t i = 0; i < no; ++i) foo[i + no * k] = F(fcn + i, fcn + k);
    for (int64_t c = 0; c < nv; ++c)
        for (int64_t k = 0; k < no; ++k) fov[k + no * c] = F(fcn + k, ndocc + c);
    for (int64_t a = 0; a < nv; ++a)
        for (int64_t c = 0; c < nv; ++c) fvv[c + nv * a] = F(ndocc + c, ndocc + a);
    Problem P;
    setup_problem(ctx, P, nao, Ca + nao * fcn, no, Ca + nao * ndocc, nv, eps.data());
    AutoOptions ao;
    ao.max_iter = opt.cc_max_iter; ao.e_conv = opt.cc_e_conv; ao.max_rms = opt.cc_max_rms; ao.do_pT = opt.do_pT != 0;
    AutoResult r = auto_rccsd_dev(ctx, P, src, foo.data(), fov.data(), fvv.data(), ao, T1_out, T2_out,
                                  ctx->amp_cb, ctx->amp_user);
    *e_cc = r.ecc;
    if (e_pt && r.has_pt) *e_pt = r.ept;
    if (iterations) *iterations = r.iterations;
    if (converged) *converged = r.converged ? 1 : 0;
    for (int k = 0; k <= r.iterations; ++k) {
        if (e_hist) e_hist[k] = r.e_hist[k];
        if (rms_hist) rms_hist[k] = r.rms_hist[k];
    }
    ctx->timings.emplace_back("alloc.host_ms", (float)(ctx->alloc_host_s * 1e3));
    ctx->timings.emplace_back("alloc.calls", (float)ctx->alloc_calls);
}
}  // namespace

extern "C" int jues_b200_auto_rccsd(jues_ctx* ctx, const double* gao, int64_t nao, const double* hao,
                                    const double* Ca, int64_t nmo, int64_t ndocc,
                                    const jues_b200_cc_options* opt, double* e_cc, double* e_pt,
                                    int* iterations, int* converged, double* e_hist, double* rms_hist,
                                    double* T1_out, double* T2_out) {
    if (is_group_call(ctx))
        return group_run(ctx, [&](jues_ctx* m, bool lead) {
            double e_, p_;
            int i_, c_;
            return jues_b200_auto_rccsd(m, gao, nao, hao, Ca, nmo, ndocc, opt, lead ? e_cc : &e_, lead ? e_pt : &p_,
                                        lead ? iterations : &i_, lead ? converged : &c_, lead ? e_hist : nullptr,
                                        lead ? rms_hist : nullptr, lead ? T1_out : nullptr, lead ? T2_out : nullptr);
        });
    JUES_API_BEGIN(ctx)
    begin_call(ctx);
    Timer total(ctx, "total");
    HostGaoHolder h;
    make_host_gao(ctx, h, gao, nao, false);
    run_auto(ctx, *h.src, hao, Ca, nmo, ndocc, opt, e_cc, e_pt, iterations, converged, e_hist, rms_hist,
             T1_out, T2_out);
    JUES_API_END(ctx)
}

extern "C" int jues_b200_auto_rccsd_t4(jues_ctx* ctx, const jues_t4* gao, const double* hao, const double* Ca,
                                       int64_t nmo, int64_t ndocc, const jues_b200_cc_options* opt,
                                       double* e_cc, double* e_pt, int* iterations, int* converged,
                                       double* e_hist, double* rms_hist, double* T1_out, double* T2_out) {
    if (is_group_call(ctx) && gao && !gao->virtual_synth) {
        ctx->last_error = "a dense device tensor lives on one GPU: pass the host array or a generated tensor to a multi-GPU group";
        return JUES_B200_ESTATE;
    }
    if (is_group_call(ctx))
        return group_run(ctx, [&](jues_ctx* m, bool lead) {
            double e_, p_;
            int i_, c_;
            return jues_b200_auto_rccsd_t4(m, gao, hao, Ca, nmo, ndocc, opt, lead ? e_cc : &e_, lead ? e_pt : &p_,
                                           lead ? iterations : &i_, lead ? converged : &c_, lead ? e_hist : nullptr,
                                           lead ? rms_hist : nullptr, lead ? T1_out : nullptr, lead ? T2_out : nullptr);
        });
    JUES_API_BEGIN(ctx)
    begin_call(ctx);
    check_t4_is_gao(gao);
    Timer total(ctx, "total");
    T4Source holder(gao);
    run_auto(ctx, *holder.src, hao, Ca, nmo, ndocc, opt, e_cc, e_pt, iterations, converged, e_hist, rms_hist,
             T1_out, T2_out);
    JUES_API_END(ctx)
}

namespace {
// host (d0,d1,d2,d3) column-major -> zero-padded device tensor (dp0..dp3)
void upload_padded_t4(jues_ctx* ctx, DTen& dst, const double* host, const int64_t d[4], const int64_t dp[4]) {
    const size_t n = (size_t)(d[0] * d[1] * d[2] * d[3]);
    dst.alloc(ctx, dp[0], dp[1], dp[2], dp[3]);
    dst.buf.zero();
    DBuf raw(ctx, n);
    JUES_CUDA(cudaMemcpyAsync(raw.p, host, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    block_copy(ctx, raw.p, d, dst.p(), dp, d);
}
}  // namespace

extern "C" int jues_b200_compute_pt(jues_ctx* ctx, const double* T1, const double* T2, const double* Vvvvo,
                                    const double* Vvooo, const double* Vvovo, const double* fo,
                                    const double* fv, int64_t nocc, int64_t nvir, double* e_pt) {
    JUES_API_BEGIN(ctx)
    begin_call(ctx);
    JUES_REQUIRE(T1 && T2 && Vvvvo && Vvooo && Vvovo && fo && fv && e_pt, "null argument");
    JUES_REQUIRE(nocc > 0 && nvir > 0, "nocc and nvir must be positive");
    Timer total(ctx, "total");
    const int64_t o = round_up(nocc, 2), v = round_up(nvir, 2);
    const int64_t K = v + o;
    DTen t1, Acat(ctx, v, v, o, K), Bq(ctx, v, o, K, o), Br(ctx, v, o, K, o), Vv(ctx, v, v, o, o);
    {
        Timer t(ctx, "pt.upload");
        DTen t2, a, b, c, OAp(ctx, v, v, o, v), ooov(ctx, o, o, o, v);
        upload_padded_matrix(ctx, t1.buf, T1, nocc, nvir, o, v);
        t1.t = Ten(t1.buf.p, o, v);
        const int64_t d2[4] = {nocc, nocc, nvir, nvir}, p2[4] = {o, o, v, v};
        upload_padded_t4(ctx, t2, T2, d2, p2);
        const int64_t da[4] = {nvir, nvir, nvir, nocc}, pa[4] = {v, v, v, o};
        upload_padded_t4(ctx, a, Vvvvo, da, pa);                       // Vvvvo[b,d,a,p] = <pd|ab>
        permute_axpby(ctx, 1.0, a, "bdap", 0.0, OAp, "abpd");
        a.release();
        const int64_t db[4] = {nvir, nocc, nocc, nocc}, pb[4] = {v, o, o, o};
        upload_padded_t4(ctx, b, Vvooo, db, pb);                       // Vvooo[c,r,q,l] = <qr|lc> = ooov[q,r,l,c]
        permute_axpby(ctx, 1.0, b, "crql", 0.0, ooov, "qrlc");
        const int64_t dc[4] = {nvir, nocc, nvir, nocc}, pc[4] = {v, o, v, o};
        upload_padded_t4(ctx, c, Vvovo, dc, pc);                       // Vvovo[a,i,b,j] = <ij|ab>
        permute_axpby(ctx, 1.0, c, "aibj", 0.0, Vv, "abij");
        pt_build_operands(ctx, o, v, OAp.p(), t2.p(), ooov.p(), Acat.p(), Bq.p(), Br.p());
    }
    double emin = fo[0], emax = fv[0];
    for (int64_t k = 0; k < nocc; ++k) emin = std::min(emin, fo[k]);
    for (int64_t k = 0; k < nvir; ++k) emax = std::max(emax, fv[k]);
    std::vector<double> eo((size_t)o, emin - 1.0e3), ev((size_t)v, emax + 1.0e3);
    std::copy(fo, fo + nocc, eo.begin());
    std::copy(fv, fv + nvir, ev.begin());
    DBuf eod(ctx, (size_t)o), evd(ctx, (size_t)v);
    JUES_CUDA(cudaMemcpyAsync(eod.p, eo.data(), o * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    JUES_CUDA(cudaMemcpyAsync(evd.p, ev.data(), v * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    JUES_CUDA(cudaStreamSynchronize(ctx->stream));
    PtInputs in;
    in.o = o; in.v = v; in.nocc = nocc;
    in.Acat = Acat.p(); in.Bq = Bq.p(); in.Br = Br.p(); in.Vv = Vv.p(); in.t1 = t1.p();
    in.eo = eod.p; in.ev = evd.p;
    Timer t(ctx, "pt.energy");
    *e_pt = pt_dev(ctx, in);
    JUES_API_END(ctx)
}

// ---------------------------------------------------------------------------------------------
// mRCCD.do_rccd (DIIS)
// ---------------------------------------------------------------------------------------------
namespace {
void run_mrccd(jues_ctx* ctx, GaoSource& src, const double* Cao, int64_t nocc, const double* Cav, int64_t nvir,
               const double* eps, int maxit, double* e, int* iterations, double* rms_hist, double* e_hist,
               double* T2_out) {
    JUES_REQUIRE(e != nullptr, "null energy output");
    Problem P;
    setup_problem(ctx, P, src.n, Cao, nocc, Cav, nvir, eps);
    MrccdResult r = mrccd_dev(ctx, P, src, maxit, T2_out, ctx->amp_cb, ctx->amp_user);
    *e = r.energy;
    if (iterations) *iterations = r.iterations;
    for (int k = 0; k < r.iterations; ++k) {
        if (rms_hist) rms_hist[k] = r.rms_hist[k];
        if (e_hist) e_hist[k] = r.e_hist[k];
    }
}
}  // namespace

extern "C" int jues_b200_mrccd(jues_ctx* ctx, const double* gao, int64_t nao, const double* Cao, int64_t nocc,
                               const double* Cav, int64_t nvir, const double* eps, int maxit, double* e_ccd,
                               int* iterations, double* rms_hist, double* e_hist, double* T2_out) {
    if (is_group_call(ctx))
        return group_run(ctx, [&](jues_ctx* m, bool lead) {
            double e_;
            int i_;
            return jues_b200_mrccd(m, gao, nao, Cao, nocc, Cav, nvir, eps, maxit, lead ? e_ccd : &e_,
                                   lead ? iterations : &i_, lead ? rms_hist : nullptr, lead ? e_hist : nullptr,
                                   lead ? T2_out : nullptr);
        });
    JUES_API_BEGIN(ctx)
    begin_call(ctx);
    Timer total(ctx, "total");
    HostGaoHolder h;
    make_host_gao(ctx, h, gao, nao, true);
    run_mrccd(ctx, *h.src, Cao, nocc, Cav, nvir, eps, maxit, e_ccd, iterations, rms_hist, e_hist, T2_out);
    JUES_API_END(ctx)
}

extern "C" int jues_b200_mrccd_t4(jues_ctx* ctx, const jues_t4* gao, const double* Cao, int64_t nocc,
                                  const double* Cav, int64_t nvir, const double* eps, int maxit, double* e_ccd,
                                  int* iterations, double* rms_hist, double* e_hist, double* T2_out) {
    if (is_group_call(ctx) && gao && !gao->virtual_synth) {
        ctx->last_error = "a dense device tensor lives on one GPU: pass the host array or a generated tensor to a multi-GPU group";
        return JUES_B200_ESTATE;
    }
    if (is_group_call(ctx))
        return group_run(ctx, [&](jues_ctx* m, bool lead) {
            double e_;
            int i_;
            return jues_b200_mrccd_t4(m, gao, Cao, nocc, Cav, nvir, eps, maxit, lead ? e_ccd : &e_,
                                      lead ? iterations : &i_, lead ? rms_hist : nullptr, lead ? e_hist : nullptr,
                                      lead ? T2_out : nullptr);
        });
    JUES_API_BEGIN(ctx)
    begin_call(ctx);
    check_t4_is_gao(gao);
    Timer total(ctx, "total");
    T4Source holder(gao);
    run_mrccd(ctx, *holder.src, Cao, nocc, Cav, nvir, eps, maxit, e_ccd, iterations, rms_hist, e_hist, T2_out);
    JUES_API_END(ctx)
}

// ---------------------------------------------------------------------------------------------
// Density-fitted variants (DF-RMP2.jl:1-46, DF-RCCD.jl:11-54)
// ---------------------------------------------------------------------------------------------
extern "C" int jues_b200_df_rmp2(jues_ctx* ctx, const double* pqP, int64_t nao, int64_t naux, const double* Jpqh,
                                 const double* Cao, int64_t nocc, const double* Cav, int64_t nvir,
                                 const double* eps, double* e_mp2) {
    if (is_group_call(ctx))
        return group_run(ctx, [&](jues_ctx* m, bool lead) {
            double e_;
            return jues_b200_df_rmp2(m, pqP, nao, naux, Jpqh, Cao, nocc, Cav, nvir, eps, lead ? e_mp2 : &e_);
        });
    JUES_API_BEGIN(ctx)
    begin_call(ctx);
    JUES_REQUIRE(e_mp2 != nullptr, "null energy output");
    Timer total(ctx, "total");
    Problem P;
    setup_problem(ctx, P, nao, Cao, nocc, Cav, nvir, eps);
    *e_mp2 = df_rmp2_dev(ctx, P, pqP, naux, Jpqh);
    JUES_API_END(ctx)
}

extern "C" int jues_b200_df_rccd(jues_ctx* ctx, const double* pqP, int64_t nao, int64_t naux, const double* Jpqh,
                                 const double* Cao, int64_t nocc, const double* Cav, int64_t nvir,
                                 const double* eps, int maxit, double* e_ccd, double* e_hist, double* T2_out) {
    if (is_group_call(ctx))
        return group_run(ctx, [&](jues_ctx* m, bool lead) {
            double e_;
            return jues_b200_df_rccd(m, pqP, nao, naux, Jpqh, Cao, nocc, Cav, nvir, eps, maxit, lead ? e_ccd : &e_,
                                     lead ? e_hist : nullptr, lead ? T2_out : nullptr);
        });
    JUES_API_BEGIN(ctx)
    begin_call(ctx);
    JUES_REQUIRE(e_ccd != nullptr, "null energy output");
    Timer total(ctx, "total");
    Problem P;
    setup_problem(ctx, P, nao, Cao, nocc, Cav, nvir, eps);
    CCResult r = df_rccd_dev(ctx, P, pqP, naux, Jpqh, maxit, T2_out, ctx->amp_cb, ctx->amp_user);
    *e_ccd = r.energy;
    if (e_hist)
        for (int k = 0; k <= maxit; ++k) e_hist[k] = r.e_hist[k];
    JUES_API_END(ctx)
}

// ---------------------------------------------------------------------------------------------
// Kernel-level check of the packed (symmetric / antisymmetric) particle-particle ladder
// ---------------------------------------------------------------------------------------------
#include "dgemm.h"

// out[i,j,a,b] = sum_ef tau[i,j,e,f] W4[e,f,a,b] evaluated exactly as `nslabs` ranks of the CC driver
// would (cc.cu: sa_ladder): per slab of the last index pack [W+|W-] from that slab only, one batched
// GEMM into the slab's column block of [L+|L-], then -- all blocks being in one buffer here instead of
// all-gathered -- unpack every slab.  Lets a single GPU exercise the b0 != 0 code paths of the kernels.
extern "C" int jues_b200_sa_ladder(jues_ctx* ctx, const double* tau, const double* W4, int64_t nocc,
                                   int64_t nvir, int nslabs, double* out) {
    JUES_API_BEGIN(ctx)
    begin_call(ctx);
    JUES_REQUIRE(tau && W4 && out, "null argument");
    JUES_REQUIRE(nocc > 0 && nvir > 0 && nslabs >= 1 && nslabs <= 64, "bad extents / slab count");
    const int64_t o = round_up(nocc, 2), v = round_up(nvir, 2 * (int64_t)nslabs), vs = v / nslabs;
    const int64_t oo = o * o, np = sa_pairs(v), ld = round_up(np, 2), nq = vs * sa_slots(v), nq_all = v * sa_slots(v);
    DTen taud, W4d;
    const int64_t dt[4] = {nocc, nocc, nvir, nvir}, pt[4] = {o, o, v, v};
    upload_padded_t4(ctx, taud, tau, dt, pt);
    const int64_t dw[4] = {nvir, nvir, nvir, nvir}, pw[4] = {v, v, v, v};
    upload_padded_t4(ctx, W4d, W4, dw, pw);
    DBuf Tpm(ctx, (size_t)(2 * oo * ld)), Lpm(ctx, (size_t)(2 * oo * nq_all)), outd(ctx, (size_t)(oo * v * v));
    pack_tau_sa(ctx, taud.p(), oo, v, ld, Tpm.p);
    for (int64_t r = 0; r < nslabs; ++r) {
        DBuf Wsa(ctx, (size_t)(2 * ld * nq));
        Wsa.zero();
        // the slab W4[e,f,w,z in S_r] is a contiguous block (last index slowest)
        pack_vvvv_sa(ctx, W4d.p() + v * v * v * (r * vs), v, r * vs, vs, ld, Wsa.p);
        GemmCall g;
        g.M = oo; g.N = nq; g.K = np; g.batch = 2;
        g.A = Tpm.p; g.lda = oo; g.strideA = oo * ld;
        g.B = Wsa.p; g.ldb = ld; g.strideB = ld * nq;
        g.C = Lpm.p + oo * nq * r; g.ldc = oo; g.strideC = oo * nq_all;
        dgemm(ctx, g);
    }
    for (int64_t r = 0; r < nslabs; ++r)
        unpack_ladder_sa(ctx, Lpm.p, oo, v, r * vs, vs, outd.p + oo * v * (r * vs));
    download_block(ctx, outd.p, pt, out, dt);
    JUES_API_END(ctx)
}
