// Context management and the BLAS.gemm! replacement entry point of the C ABI.
#include "jues_common.h"
#include "dgemm.h"
#include "api_util.h"
#include "tensor_ops.h"

#include <mutex>

namespace jues {
std::string g_init_error;  // message of a failed jues_b200_init (no context exists yet)
}

using namespace jues;

extern "C" const char* jues_b200_version(void) { return "jues_b200 0.1.0 sm_100a"; }

extern "C" int jues_b200_init(jues_ctx** out, int device) {
    if (!out) return JUES_B200_EINVAL;
    *out = nullptr;
    jues_ctx* ctx = nullptr;
    try {
        int ndev = 0;
        cudaError_t e = cudaGetDeviceCount(&ndev);
        if (e != cudaSuccess || ndev == 0) {
            cudaGetLastError();
            throw Error(JUES_B200_ECUDA,
                        std::string("no CUDA device available (") + cudaGetErrorString(e) +
                            "); jues_b200 has no CPU fallback");
        }
        JUES_REQUIRE(device >= 0 && device < ndev, "device index out of range");
        JUES_CUDA(cudaSetDevice(device));
        cudaDeviceProp prop;
        JUES_CUDA(cudaGetDeviceProperties(&prop, device));
        if (prop.major != 10) {
            char buf[256];
            snprintf(buf, sizeof buf,
                     "device %d (%s) is sm_%d%d; jues_b200 is built for sm_100a only", device, prop.name,
                     prop.major, prop.minor);
            throw Error(JUES_B200_ECUDA, buf);
        }
        ctx = new jues_ctx();
        ctx->device = device;
        ctx->sm_count = prop.multiProcessorCount;
        if (const char* mb = getenv("JUES_B200_BIG_MB")) ctx->big_bytes = (size_t)std::max(1, atoi(mb)) << 20;
        ctx->sync_comm = getenv("JUES_B200_SYNC_COMM") != nullptr;
        if (const char* tr = getenv("JUES_B200_TRACE")) ctx->trace = std::max(1, atoi(tr));
        JUES_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        JUES_CUDA(cudaStreamCreateWithFlags(&ctx->comm_stream, cudaStreamNonBlocking));
        {   // keep freed blocks in the stream-ordered pool (all work of a context is on one stream)
            cudaMemPool_t pool;
            JUES_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
            unsigned long long thr = ~0ull;
            JUES_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
        }
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        JUES_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (!fn || qres != cudaDriverEntryPointSuccess)
            throw Error(JUES_B200_ECUDA, "cuTensorMapEncodeTiled not available from the driver");
        ctx->encode = reinterpret_cast<PFN_encodeTiled>(fn);
        ctx->red_cap = 1 << 16;
        JUES_CUDA(cudaMalloc(&ctx->red_dev, ctx->red_cap * sizeof(double)));
        JUES_CUDA(cudaMallocHost(&ctx->red_host, ctx->red_cap * sizeof(double)));
        *out = ctx;
        return JUES_B200_OK;
    } catch (const Error& e) {
        g_init_error = e.what();
        if (ctx) jues_b200_finalize(ctx);
        return e.code;
    } catch (const std::exception& e) {
        g_init_error = e.what();
        if (ctx) jues_b200_finalize(ctx);
        return JUES_B200_ECUDA;
    }
}

extern "C" void jues_b200_finalize(jues_ctx* ctx) {
    if (!ctx) return;
    if (ctx->group) {                       // leader of a single-process multi-GPU group: members first
        std::vector<jues_ctx*>* g = ctx->group;
        ctx->group = nullptr;
        for (jues_ctx* m : *g)
            if (m != ctx) jues_b200_finalize(m);
        delete g;
    }
    cudaSetDevice(ctx->device);
    resolve_timers(ctx);
    jues::dist_teardown(ctx);
    for (auto& kv : ctx->big_free) cudaFree(kv.second);
    ctx->big_free.clear();
    if (ctx->red_dev) cudaFree(ctx->red_dev);
    if (ctx->red_host) cudaFreeHost(ctx->red_host);
    if (ctx->comm_stream) cudaStreamDestroy(ctx->comm_stream);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" const char* jues_b200_last_error(jues_ctx* ctx) {
    if (!ctx) return g_init_error.c_str();
    return ctx->last_error.c_str();
}

extern "C" int jues_b200_set_trace(jues_ctx* ctx, int level) {
    if (!ctx || level < 0 || level > 2) return JUES_B200_EINVAL;
    ctx->trace = level;
    return JUES_B200_OK;
}

extern "C" int jues_b200_get_phases(jues_ctx* ctx, jues_b200_phase* out, int cap) {
    if (!ctx) return JUES_B200_EINVAL;
    cudaSetDevice(ctx->device);
    resolve_timers(ctx);
    int n = 0;
    for (auto& kv : ctx->timings) {
        if (out && n < cap) {
            strncpy(out[n].name, kv.first.c_str(), sizeof(out[n].name) - 1);
            out[n].name[sizeof(out[n].name) - 1] = 0;
            out[n].ms = kv.second;
        }
        ++n;
    }
    return n;
}

extern "C" int jues_b200_get_counters(jues_ctx* ctx, double* gemm_flops, int64_t* gemm_launches,
                                      int64_t* aux_launches, int64_t* bytes_peak) {
    if (!ctx) return JUES_B200_EINVAL;
    if (gemm_flops) *gemm_flops = ctx->stats.gemm_flops;
    if (gemm_launches) *gemm_launches = ctx->stats.gemm_launches;
    if (aux_launches) *aux_launches = ctx->stats.aux_launches;
    if (bytes_peak) *bytes_peak = (int64_t)ctx->bytes_peak;
    return JUES_B200_OK;
}

// ---------------------------------------------------------------------------------------------
// BLAS.gemm! replacement with host operands
// ---------------------------------------------------------------------------------------------
namespace {
// copy a column-major host matrix (rows x cols, leading dimension ld) into a zero-padded device
// matrix with an even leading dimension
void upload_matrix(jues_ctx* ctx, DBuf& d, int64_t& ld_dev, const double* host, int64_t rows,
                   int64_t cols, int64_t ld) {
    ld_dev = round_up(rows, 2);
    d.alloc(ctx, (size_t)ld_dev * cols);
    if (ld_dev != rows) d.zero();
    JUES_CUDA(cudaMemcpy2DAsync(d.p, ld_dev * 8, host, ld * 8, rows * 8, cols, cudaMemcpyHostToDevice,
                                ctx->stream));
}
}  // namespace

extern "C" int jues_b200_dgemm(jues_ctx* ctx, char transA, char transB, int64_t M, int64_t N, int64_t K,
                               double alpha, const double* A, int64_t lda, const double* B, int64_t ldb,
                               double beta, double* C, int64_t ldc) {
    JUES_API_BEGIN(ctx)
    const bool tA = (transA == 'T' || transA == 't');
    const bool tB = (transB == 'T' || transB == 't');
    JUES_REQUIRE(tA || transA == 'N' || transA == 'n', "transA must be 'N' or 'T'");
    JUES_REQUIRE(tB || transB == 'N' || transB == 'n', "transB must be 'N' or 'T'");
    JUES_REQUIRE(M >= 0 && N >= 0 && K >= 0, "negative dimension");
    if (M == 0 || N == 0) return JUES_B200_OK;
    JUES_REQUIRE(A && B && C, "null matrix");
    const int64_t ar = tA ? K : M, ac = tA ? M : K;
    const int64_t br = tB ? N : K, bc = tB ? K : N;
    JUES_REQUIRE(lda >= ar && ldb >= br && ldc >= M, "leading dimension too small");
    JUES_REQUIRE(K > 0, "K must be positive");
    DBuf dA, dB, dC;
    int64_t la, lb, lc;
    upload_matrix(ctx, dA, la, A, ar, ac, lda);
    upload_matrix(ctx, dB, lb, B, br, bc, ldb);
    lc = round_up(M, 2);
    dC.alloc(ctx, (size_t)lc * N);
    if (beta != 0.0)
        JUES_CUDA(cudaMemcpy2DAsync(dC.p, lc * 8, C, ldc * 8, M * 8, N, cudaMemcpyHostToDevice, ctx->stream));
    GemmCall g;
    g.transA = tA; g.transB = tB;
    g.M = M; g.N = N; g.K = K;
    g.A = dA.p; g.lda = la; g.B = dB.p; g.ldb = lb; g.C = dC.p; g.ldc = lc;
    g.alpha = alpha; g.beta = beta;
    g.force_cfg = ctx_force_cfg(ctx);
    dgemm(ctx, g);
    JUES_CUDA(cudaMemcpy2DAsync(C, ldc * 8, dC.p, lc * 8, M * 8, N, cudaMemcpyDeviceToHost, ctx->stream));
    JUES_CUDA(cudaStreamSynchronize(ctx->stream));
    JUES_API_END(ctx)
}

extern "C" int jues_b200_dgemm_bench(jues_ctx* ctx, char transA, char transB, int64_t M, int64_t N,
                                     int64_t K, int reps, double* ms_avg) {
    JUES_API_BEGIN(ctx)
    const bool tA = (transA == 'T' || transA == 't');
    const bool tB = (transB == 'T' || transB == 't');
    JUES_REQUIRE(M > 0 && N > 0 && K > 0 && reps > 0 && ms_avg, "bad arguments");
    const int64_t ar = round_up(tA ? K : M, 2), ac = tA ? M : K;
    const int64_t br = round_up(tB ? N : K, 2), bc = tB ? K : N;
    const int64_t lc = round_up(M, 2);
    DBuf dA(ctx, (size_t)ar * ac), dB(ctx, (size_t)br * bc), dC(ctx, (size_t)lc * N);
    fill_pattern(ctx, dA.p, dA.n, 1);
    fill_pattern(ctx, dB.p, dB.n, 2);
    GemmCall g;
    g.transA = tA; g.transB = tB;
    g.M = M; g.N = N; g.K = K;
    g.A = dA.p; g.lda = ar; g.B = dB.p; g.ldb = br; g.C = dC.p; g.ldc = lc;
    g.force_cfg = ctx_force_cfg(ctx);
    dgemm(ctx, g);  // warm-up
    dgemm(ctx, g);
    JUES_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0, ctx->stream);
    for (int r = 0; r < reps; ++r) dgemm(ctx, g);
    cudaEventRecord(e1, ctx->stream);
    JUES_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *ms_avg = ms / reps;
    JUES_API_END(ctx)
}

// Determinism / self-consistency stress of the DGEMM kernel at a given shape: the same product is
// launched `reps`+1 times on device-resident pseudo-random operands; every repetition is compared with
// the first on the device (sum of squared differences, exact 0 expected).  batch > 1 uses the layout
// of the quarter transforms (A [M,K,batch], B shared, C [M,N,batch]).
extern "C" int jues_b200_dgemm_stress(jues_ctx* ctx, char transA, char transB, int64_t M, int64_t N,
                                      int64_t K, int64_t batch, int reps, int* n_bad, double* worst_sqdiff) {
    JUES_API_BEGIN(ctx)
    const bool tA = (transA == 'T' || transA == 't');
    const bool tB = (transB == 'T' || transB == 't');
    JUES_REQUIRE(M > 0 && N > 0 && K > 0 && batch > 0 && reps > 0 && reps <= 4096 && n_bad && worst_sqdiff,
                 "bad arguments");
    JUES_REQUIRE(batch == 1 || (!tA && !tB && (M & 1) == 0), "batched stress: 'N','N' with even M only");
    const int64_t ar = round_up(tA ? K : M, 2), ac = tA ? M : K;
    const int64_t br = round_up(tB ? N : K, 2), bc = tB ? K : N;
    const int64_t lc = round_up(M, 2);
    DBuf dA(ctx, (size_t)ar * ac * batch), dB(ctx, (size_t)br * bc), dC0(ctx, (size_t)lc * N * batch),
        dC1(ctx, (size_t)lc * N * batch), res(ctx, (size_t)reps);
    fill_pattern(ctx, dA.p, dA.n, 1);
    fill_pattern(ctx, dB.p, dB.n, 2);
    GemmCall g;
    g.transA = tA; g.transB = tB;
    g.M = M; g.N = N; g.K = K; g.batch = batch;
    g.A = dA.p; g.lda = ar; g.strideA = ar * ac;
    g.B = dB.p; g.ldb = br; g.strideB = 0;
    g.C = dC0.p; g.ldc = lc; g.strideC = lc * N;
    g.force_cfg = ctx_force_cfg(ctx);
    dC0.zero();
    dgemm(ctx, g);
    g.C = dC1.p;
    for (int r = 0; r < reps; ++r) {
        dC1.zero();
        dgemm(ctx, g);
        sqdiff_async(ctx, dC0.n, dC0.p, dC1.p, res.p + r);
    }
    std::vector<double> h((size_t)reps);
    JUES_CUDA(cudaMemcpyAsync(h.data(), res.p, (size_t)reps * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    JUES_CUDA(cudaStreamSynchronize(ctx->stream));
    *n_bad = 0; *worst_sqdiff = 0.0;
    for (double x : h) {
        if (x != 0.0) ++*n_bad;
        if (x > *worst_sqdiff || x != x) *worst_sqdiff = x;
    }
    JUES_API_END(ctx)
}
