// Host launcher of the sm_100a TMA + DMMA FP64 GEMM: builds the TMA tensor maps, picks a tile
// configuration, launches on the context's stream.
#include "dgemm.h"
#include <algorithm>
#include <cmath>
#include "dgemm_sm100.cuh"
#include "tensor_ops.h"

namespace jues {

using namespace gemm;

namespace {

struct Cfg {
    int BM, BN, WM, WN, STAGES;
    const char* name;
    double cost;  // relative cost per tile-flop (smaller tiles re-read operands more often)
};
// all configurations use 8 consumer warps + 1 producer warp
static const Cfg kCfgs[] = {
    {128, 128, 64, 32, 4, "128x128x16_w64x32_s4", 1.00},
    {64, 128, 32, 32, 6, "64x128x16_w32x32_s6", 1.12},
    {128, 64, 64, 16, 6, "128x64x16_w64x16_s6", 1.12},
    {64, 64, 32, 16, 8, "64x64x16_w32x16_s8", 1.30},
    // 1 x 8 warp layouts: 16-row granularity in M against tile / wave quantisation (e.g. M = 400 = 5 x 80,
    // M = 2000 -> 18 x 112 x 16 tiles = 288 CTAs = 1.95 waves of 148 SMs)
    {112, 128, 112, 16, 5, "112x128x16_w112x16_s5", 1.03},
    {96, 128, 96, 16, 5, "96x128x16_w96x16_s5", 1.05},
    {80, 128, 80, 16, 6, "80x128x16_w80x16_s6", 1.07},
    {48, 128, 48, 16, 8, "48x128x16_w48x16_s8", 1.20},
    // narrow tiles for the skinny products of the sweep (N or M = nocc ~ 20: a 64-wide tile is 3x padding)
    {128, 32, 32, 16, 8, "128x32x16_w32x16_s8", 1.45},
    {32, 128, 16, 32, 8, "32x128x16_w16x32_s8", 1.45},
};
constexpr int kNumCfgs = sizeof(kCfgs) / sizeof(kCfgs[0]);

typedef void (*KernelFn)(const CUtensorMap, const CUtensorMap, const Params);

template <bool A_KC, bool B_KC>
KernelFn kernel_for(int cfg, int* smem) {
    switch (cfg) {
        case 0: *smem = SmemLayout<128, 128, 4>::TOTAL; return dgemm_tma_dmma<A_KC, B_KC, 128, 128, 64, 32, 4>;
        case 1: *smem = SmemLayout<64, 128, 6>::TOTAL;  return dgemm_tma_dmma<A_KC, B_KC, 64, 128, 32, 32, 6>;
        case 2: *smem = SmemLayout<128, 64, 6>::TOTAL;  return dgemm_tma_dmma<A_KC, B_KC, 128, 64, 64, 16, 6>;
        case 3: *smem = SmemLayout<64, 64, 8>::TOTAL;   return dgemm_tma_dmma<A_KC, B_KC, 64, 64, 32, 16, 8>;
        case 4: *smem = SmemLayout<112, 128, 5>::TOTAL; return dgemm_tma_dmma<A_KC, B_KC, 112, 128, 112, 16, 5>;
        case 5: *smem = SmemLayout<96, 128, 5>::TOTAL;  return dgemm_tma_dmma<A_KC, B_KC, 96, 128, 96, 16, 5>;
        case 6: *smem = SmemLayout<80, 128, 6>::TOTAL;  return dgemm_tma_dmma<A_KC, B_KC, 80, 128, 80, 16, 6>;
        case 7: *smem = SmemLayout<48, 128, 8>::TOTAL;  return dgemm_tma_dmma<A_KC, B_KC, 48, 128, 48, 16, 8>;
        case 8: *smem = SmemLayout<128, 32, 8>::TOTAL;  return dgemm_tma_dmma<A_KC, B_KC, 128, 32, 32, 16, 8>;
        default: *smem = SmemLayout<32, 128, 8>::TOTAL; return dgemm_tma_dmma<A_KC, B_KC, 32, 128, 16, 32, 8>;
    }
}

KernelFn pick_kernel(bool a_kc, bool b_kc, int cfg, int* smem) {
    if (a_kc) return b_kc ? kernel_for<true, true>(cfg, smem) : kernel_for<true, false>(cfg, smem);
    return b_kc ? kernel_for<false, true>(cfg, smem) : kernel_for<false, false>(cfg, smem);
}

void make_map(jues_ctx* ctx, CUtensorMap* map, const double* base, int64_t d0, int64_t d1,
              int64_t ld, int64_t batch, int64_t bstride, int box0, int box1) {
    JUES_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "GEMM operand must be 16-byte aligned");
    JUES_REQUIRE((ld & 1) == 0 && ld >= d0, "GEMM operand leading dimension must be even and >= rows");
    JUES_REQUIRE(batch == 1 || (bstride & 1) == 0, "GEMM batch stride must be even");
    cuuint64_t dims[3] = {(cuuint64_t)d0, (cuuint64_t)d1, (cuuint64_t)batch};
    cuuint64_t bs = batch == 1 ? (cuuint64_t)ld * (cuuint64_t)d1 * 8ull : (cuuint64_t)bstride * 8ull;
    if (bs == 0) bs = 16;
    cuuint64_t strides[2] = {(cuuint64_t)ld * 8ull, bs};
    cuuint32_t box[3] = {(cuuint32_t)box0, (cuuint32_t)box1, 1u};
    cuuint32_t estr[3] = {1u, 1u, 1u};
    CUresult r = ctx->encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double*>(base), dims,
                             strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char buf[256];
        snprintf(buf, sizeof buf,
                 "cuTensorMapEncodeTiled failed (%d) dims=(%lld,%lld,%lld) ld=%lld bstride=%lld box=(%d,%d)",
                 (int)r, (long long)d0, (long long)d1, (long long)batch, (long long)ld,
                 (long long)bstride, box0, box1);
        throw Error(JUES_B200_ECUDA, buf);
    }
}

// Pick the tile configuration and split-K factor that minimise a simple time model:
// waves(tiles*split / SMs) * (stages per CTA + fixed prologue/epilogue cost) * tile area * tile cost.
// (Two refinements were measured at BASELINE config 3 and rejected: a cycle-level model with an HBM floor
// and up to 256 splits -- 6.99 ms of GEMM time per sweep against 6.85 ms, profiles/r02/session_r02f.log --
// and counting only the computed sub-tiles of partial tiles -- 7.41 ms, profiles/r02/gemm_list_c3_r02h.txt.
// The skinny products of the sweep are bandwidth bound and want a different kernel, not a different tile.)
void choose_cfg(jues_ctx* ctx, int64_t M, int64_t N, int64_t K, int64_t batch, bool allow_split,
                int* cfg_out, int* split_out) {
    int best = 0, best_split = 1;
    double best_t = 1e300;
    const double sms = ctx->sm_count > 0 ? ctx->sm_count : 148;
    const int KT = (int)((K + BK - 1) / BK);
    for (int c = 0; c < kNumCfgs; ++c) {
        const double tiles = (double)((M + kCfgs[c].BM - 1) / kCfgs[c].BM) *
                             (double)((N + kCfgs[c].BN - 1) / kCfgs[c].BN) * (double)batch;
        const double area = (double)kCfgs[c].BM * kCfgs[c].BN;
        int max_split = 1;
        // (a product of one or two tiles may split 128-way: one CTA per SM)
        if (allow_split && tiles < sms) max_split = (int)std::min<double>(tiles * 64 <= sms ? 128.0 : 64.0, std::max(1.0, KT / 8.0));
        for (int sp = 1; sp <= max_split; sp = sp < 4 ? sp + 1 : sp * 2) {
            const int ktp = (KT + sp - 1) / sp;
            const double waves = ceil(tiles * sp / sms);
            // per-CTA cost in "k-stage" units: mainloop stages + ~6 stages of prologue/epilogue,
            // plus the reduction pass over the split workspace
            double t = waves * (ktp + 6.0) * area * kCfgs[c].cost;
            if (sp > 1) t += 0.05 * sp * (double)M * (double)N * (double)batch / sms * 16.0;
            if (t < best_t * 0.999) {
                best_t = t;
                best = c;
                best_split = sp;
            }
        }
    }
    *cfg_out = best;
    *split_out = best_split;
}

bool persistent_gemm() {
    static const bool off = getenv("JUES_B200_GEMM_NONPERSISTENT") != nullptr;   // A/B switch for measurements
    return !off;
}

}  // namespace

int dgemm_num_configs() { return kNumCfgs; }
const char* dgemm_config_name(int cfg) { return (cfg >= 0 && cfg < kNumCfgs) ? kCfgs[cfg].name : "?"; }

void dgemm(jues_ctx* ctx, const GemmCall& g) {
    if (g.M <= 0 || g.N <= 0 || g.batch <= 0) return;
    JUES_REQUIRE(g.K > 0, "GEMM with K == 0");
    JUES_REQUIRE(g.A && g.B && g.C, "GEMM null operand");
    JUES_REQUIRE(g.M < (1ll << 31) && g.N < (1ll << 31) && g.K < (1ll << 31), "GEMM dimension too large");
    JUES_REQUIRE(g.ldc >= g.M, "GEMM ldc < M");
    Timer* tk = nullptr;
    if (ctx->trace >= 2) {
        char nm[48];
        snprintf(nm, sizeof nm, "gemm %lldx%lldx%lldx%lld %c%c%s", (long long)g.M, (long long)g.N, (long long)g.K,
                 (long long)g.batch, g.transA ? 'T' : 'N', g.transB ? 'T' : 'N', g.beta != 0.0 ? "+" : "");
        tk = new Timer(ctx, nm);
    }
    // bandwidth-bound skinny products go to the streaming FMA kernels (skinny.cu)
    if (skinny_gemm(ctx, g)) {
        delete tk;
        ctx->stats.gemm_flops += 2.0 * (double)g.M * (double)g.N * (double)g.K;
        ctx->stats.gemm_launches += 1;
        return;
    }
    const bool a_kc = g.transA;   // A stored K x M
    const bool b_kc = !g.transB;  // B stored K x N
    int cfg = 0, ksplit = 1;
    choose_cfg(ctx, g.M, g.N, g.K, g.batch, true, &cfg, &ksplit);
    if (g.force_cfg >= 0) {
        cfg = g.force_cfg % kNumCfgs;
        if (g.force_cfg >= kNumCfgs) ksplit = g.force_cfg / kNumCfgs + 1;  // testing: cfg + 4*(split-1)
    }
    const Cfg& c = kCfgs[cfg];
    const int KT_all = (int)((g.K + BK - 1) / BK);
    if (ksplit > KT_all) ksplit = KT_all;
    int kt_per_split = (KT_all + ksplit - 1) / ksplit;
    ksplit = (KT_all + kt_per_split - 1) / kt_per_split;  // no empty splits

    CUtensorMap mapA, mapB;
    const int bmulA = (g.batch > 1 && g.strideA != 0) ? 1 : 0;
    const int bmulB = (g.batch > 1 && g.strideB != 0) ? 1 : 0;
    const int64_t nbA = bmulA ? g.batch : 1, nbB = bmulB ? g.batch : 1;
    if (a_kc) make_map(ctx, &mapA, g.A, g.K, g.M, g.lda, nbA, g.strideA, BK, c.BM);
    else      make_map(ctx, &mapA, g.A, g.M, g.K, g.lda, nbA, g.strideA, 16, BK);
    if (b_kc) make_map(ctx, &mapB, g.B, g.K, g.N, g.ldb, nbB, g.strideB, BK, c.BN);
    else      make_map(ctx, &mapB, g.B, g.N, g.K, g.ldb, nbB, g.strideB, 16, BK);

    Params p;
    p.M = (int)g.M; p.N = (int)g.N; p.K = (int)g.K;
    p.tilesM = (int)((g.M + c.BM - 1) / c.BM);
    p.tilesN = (int)((g.N + c.BN - 1) / c.BN);
    p.tiles_per_batch = (long long)p.tilesM * p.tilesN;
    p.batch = (int)g.batch;
    p.ksplit = ksplit;
    p.bmulA = bmulA; p.bmulB = bmulB;
    p.kt_per_split = kt_per_split;
    p.raster_n_fast = p.tilesM >= p.tilesN ? 1 : 0;
    DBuf work;
    if (ksplit > 1) {
        // each split writes its own dense M x N slice; a second kernel sums the slices in a fixed
        // order (deterministic) and applies alpha/beta
        work.alloc(ctx, (size_t)g.M * g.N * ksplit * g.batch);
        p.C = work.p; p.ldc = g.M; p.strideSplit = g.M * g.N; p.strideC = g.M * g.N * ksplit;
        p.alpha = 1.0; p.beta = 0.0;
        p.Cin = nullptr;
    } else {
        p.C = g.C; p.ldc = g.ldc; p.strideC = g.strideC; p.strideSplit = 0;
        p.alpha = g.alpha; p.beta = g.beta;
        p.Cin = g.beta != 0.0 ? g.Cin : nullptr;
    }
    const long long total = p.tiles_per_batch * g.batch * ksplit;
    JUES_REQUIRE(total < (1ll << 31), "GEMM grid too large");

    int smem = 0;
    KernelFn fn = pick_kernel(a_kc, b_kc, cfg, &smem);
    // the attribute is per (device, kernel): remembered in the context, not in a process-wide static
    if (ctx->smem_attr_done.insert((const void*)fn).second)
        JUES_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int threads = (c.BM / c.WM) * (c.BN / c.WN) * 32 + 32;
    p.total_tiles = total;
    // persistent CTAs (one per SM) once there is more than ~a wave of tiles; otherwise one CTA per tile
    const long long sms = ctx->sm_count > 0 ? ctx->sm_count : 148;
    // Persistent CTAs pay off when tiles are short (few k-stages: the per-tile launch / barrier-init /
    // first-load latency is a visible fraction; measured +10 % on the quarter transforms, K = N) and
    // cost ~1-3 % on long-K tiles (static tile assignment, extra live registers), so: short K only.
    const bool persistent = persistent_gemm() && kt_per_split <= 32 && total > sms;
    const long long grid = persistent ? sms : total;
    fn<<<(unsigned)grid, threads, smem, ctx->stream>>>(mapA, mapB, p);
    delete tk;
    JUES_CUDA(cudaGetLastError());
    ctx->stats.gemm_flops += 2.0 * (double)g.M * (double)g.N * (double)g.K * (double)g.batch;
    ctx->stats.gemm_launches += 1;
    if (ksplit > 1)
        splitk_reduce(ctx, work.p, ksplit, g.M, g.N, g.batch, g.alpha, g.beta, g.C, g.ldc, g.strideC, g.Cin);
}

}  // namespace jues
