// Streaming kernels for the "skinny" products of the path: C = alpha * op(A) op(B) + beta * C where one
// operand is large and the flop / byte ratio is far below the FP64 ridge (K <= ~100 or min(M, N) <= ~20; the
// t1-dressed integral terms of RCCSD.jl:187-259 and the o^2 v^3-sized intermediates).  The TMA + DMMA tile
// kernel is 3-4x above the HBM floor on these whatever its configuration (profiles/r02/gemm_autotune.log:
// tile padding makes it DMMA-issue bound, and few, short tiles cannot keep enough bytes in flight).  Here the
// large operand is read exactly once, coalesced, by plain FP64 FMA threads; the small operand lives in shared
// memory.  Results are deterministic (fixed summation order, no atomics).
//
//   rows kernels : out(r, c) = sum_k L(r, k) S(k, c),  r huge, c <= 256, K moderate.  One thread per row and
//                  group of columns.  L is either r-contiguous ("N": loaded straight, coalesced over rows) or
//                  k-contiguous ("T": 32-deep k tiles transposed through shared memory).  Generic strides for
//                  S and out make the same kernels serve C = A B with M huge and (transposed) with N huge.
//   K-huge kernel: C (M x N, both <= 128) = A (M x K, M-contiguous) B^T (N x K, N-contiguous), K >= 8192,
//                  split over ~2 CTAs per SM; partial tiles are summed by splitk_reduce in a fixed order.
#include "dgemm.h"
#include "tensor_ops.h"

#include <algorithm>

namespace jues {

namespace {

struct RowsArgs {
    long long R;            // rows of the large operand / of the output
    int K, nc;              // inner extent, number of output columns
    const double* L; long long ldl;       // large operand: "N": L[r + k*ldl], "T": L[k + r*ldl]
    const double* S; long long ssk, ssc;  // small operand S(k, c) at S[k*ssk + c*ssc]
    double* O; long long sor, soc;        // out(r, c) at O[r*sor + c*soc]
    double alpha, beta;
    int kc;                 // k-chunk held in shared memory at a time
};

constexpr int kTX = 64;     // rows per block

// NCG columns per thread; blockDim = (kTX, TY), TY * NCG >= nc.
template <int NCG, bool L_KCONTIG>
__global__ void __launch_bounds__(kTX * 8) skinny_rows_kernel(RowsArgs a) {
    extern __shared__ double sm[];
    const int tx = threadIdx.x, ty = threadIdx.y, TY = blockDim.y;
    const int ncp = TY * NCG;                       // padded column count of the shared S chunk
    double* Ss = sm;                                // [kc][ncp]
    double* Lt = sm + (size_t)a.kc * ncp;           // [32][kTX + 1]   (L_KCONTIG only)
    const long long r0 = (long long)blockIdx.x * kTX;
    const long long r = r0 + tx;
    const bool live = r < a.R;
    const int c0 = ty * NCG;
    double acc[NCG];
#pragma unroll
    for (int c = 0; c < NCG; ++c) acc[c] = 0.0;
    const int nthreads = kTX * TY, tid = ty * kTX + tx;
    for (int k0 = 0; k0 < a.K; k0 += a.kc) {
        const int kn = min(a.kc, a.K - k0);
        __syncthreads();
        for (int e = tid; e < kn * ncp; e += nthreads) {
            const int c = e % ncp, k = e / ncp;
            Ss[e] = c < a.nc ? a.S[(long long)(k0 + k) * a.ssk + (long long)c * a.ssc] : 0.0;
        }
        __syncthreads();
        if (!L_KCONTIG) {
            if (live) {
                const double* lp = a.L + r + (long long)k0 * a.ldl;
#pragma unroll 4
                for (int k = 0; k < kn; ++k) {
                    const double x = lp[(long long)k * a.ldl];
                    const double* s = Ss + k * ncp + c0;
#pragma unroll
                    for (int c = 0; c < NCG; ++c) acc[c] = fma(x, s[c], acc[c]);
                }
            }
        } else {
            for (int kt = 0; kt < kn; kt += 32) {
                const int kw = min(32, kn - kt);
                __syncthreads();
                // tile [kTX rows][32 k]: lanes along k (contiguous), warps over rows
                for (int e = tid; e < kTX * 32; e += nthreads) {
                    const int kk = e & 31, rr = e >> 5;
                    const long long row = r0 + rr;
                    Lt[kk * (kTX + 1) + rr] = (kk < kw && row < a.R) ? a.L[(long long)(k0 + kt + kk) + row * a.ldl] : 0.0;
                }
                __syncthreads();
#pragma unroll 4
                for (int kk = 0; kk < kw; ++kk) {
                    const double x = Lt[kk * (kTX + 1) + tx];
                    const double* s = Ss + (kt + kk) * ncp + c0;
#pragma unroll
                    for (int c = 0; c < NCG; ++c) acc[c] = fma(x, s[c], acc[c]);
                }
            }
        }
    }
    if (live) {
#pragma unroll
        for (int c = 0; c < NCG; ++c) {
            if (c0 + c < a.nc) {
                double* o = a.O + r * a.sor + (long long)(c0 + c) * a.soc;
                const double v = a.alpha * acc[c];
                *o = a.beta == 0.0 ? v : v + a.beta * *o;
            }
        }
    }
}

template <int NCG>
void launch_rows(jues_ctx* ctx, const RowsArgs& a, bool kcontig, int TY, size_t smem) {
    const dim3 block(kTX, TY);
    const unsigned grid = (unsigned)((a.R + kTX - 1) / kTX);
    if (kcontig) {
        if (ctx->smem_attr_done.insert((const void*)skinny_rows_kernel<NCG, true>).second)
            JUES_CUDA(cudaFuncSetAttribute(skinny_rows_kernel<NCG, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        skinny_rows_kernel<NCG, true><<<grid, block, smem, ctx->stream>>>(a);
    } else {
        if (ctx->smem_attr_done.insert((const void*)skinny_rows_kernel<NCG, false>).second)
            JUES_CUDA(cudaFuncSetAttribute(skinny_rows_kernel<NCG, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        skinny_rows_kernel<NCG, false><<<grid, block, smem, ctx->stream>>>(a);
    }
    JUES_CUDA(cudaGetLastError());
}

void rows_product(jues_ctx* ctx, RowsArgs a, bool kcontig) {
    // columns per thread and column groups per block: NCG * TY >= nc, TY <= 8
    int NCG = a.nc <= 8 ? (a.nc <= 1 ? 1 : (a.nc <= 4 ? 4 : 8)) : (a.nc <= 64 ? 8 : (a.nc <= 128 ? 16 : 32));
    int TY = (a.nc + NCG - 1) / NCG;
    if (TY > 8) { NCG = 32; TY = (a.nc + 31) / 32; }
    const int ncp = TY * NCG;
    // shared memory: S chunk [kc][ncp] (<= 40 KB so that several blocks share an SM) + the transposed L tile
    const size_t tile = kcontig ? (size_t)32 * (kTX + 1) * 8 : 0;
    int kc = (int)std::max<size_t>(32, (size_t(40) << 10) / ((size_t)ncp * 8));
    kc = std::min(a.K, kc & ~31);
    if (kc <= 0) kc = std::min(a.K, 32);
    a.kc = kc;
    const size_t smem = (size_t)kc * ncp * 8 + tile;
    switch (NCG) {
        case 1: launch_rows<1>(ctx, a, kcontig, TY, smem); break;
        case 4: launch_rows<4>(ctx, a, kcontig, TY, smem); break;
        case 8: launch_rows<8>(ctx, a, kcontig, TY, smem); break;
        case 16: launch_rows<16>(ctx, a, kcontig, TY, smem); break;
        default: launch_rows<32>(ctx, a, kcontig, TY, smem); break;
    }
    ctx->stats.aux_launches += 1;
}

// ---- K huge, M and N small:  W[z] (M x N) = sum_{k in chunk z} A[:, k] B[:, k]^T ------------------------
struct KHugeArgs {
    int M, N;
    long long K, kper;      // k range per CTA
    const double* A; long long lda;   // A[m + k*lda]
    const double* B; long long ldb;   // B[n + k*ldb]
    double* W;              // [gridDim.x][M*N]
};

constexpr int kKC = 32;     // k-rows staged per step

template <int TM, int TN>
__global__ void __launch_bounds__(256) skinny_khuge_kernel(KHugeArgs a) {
    extern __shared__ double sm[];
    const int Mp = (a.M + TM - 1) / TM * TM, Np = (a.N + TN - 1) / TN * TN;
    double* As = sm;                       // [kKC][Mp]
    double* Bs = sm + kKC * Mp;            // [kKC][Np]
    const int gm = Mp / TM, gn = Np / TN;
    const int tid = threadIdx.x;
    const bool worker = tid < gm * gn;
    const int m0 = (tid % gm) * TM, n0 = (tid / gm) * TN;
    double acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.0;
    const long long kb = (long long)blockIdx.x * a.kper, ke = min(a.K, kb + a.kper);
    for (long long k0 = kb; k0 < ke; k0 += kKC) {
        const int kn = (int)min((long long)kKC, ke - k0);
        __syncthreads();
        for (int e = tid; e < kKC * Mp; e += blockDim.x) {
            const int m = e % Mp, kk = e / Mp;
            As[e] = (m < a.M && kk < kn) ? a.A[m + (k0 + kk) * a.lda] : 0.0;
        }
        for (int e = tid; e < kKC * Np; e += blockDim.x) {
            const int n = e % Np, kk = e / Np;
            Bs[e] = (n < a.N && kk < kn) ? a.B[n + (k0 + kk) * a.ldb] : 0.0;
        }
        __syncthreads();
        if (worker) {
#pragma unroll 4
            for (int kk = 0; kk < kKC; ++kk) {
                double x[TM], y[TN];
#pragma unroll
                for (int i = 0; i < TM; ++i) x[i] = As[kk * Mp + m0 + i];
#pragma unroll
                for (int j = 0; j < TN; ++j) y[j] = Bs[kk * Np + n0 + j];
#pragma unroll
                for (int i = 0; i < TM; ++i)
#pragma unroll
                    for (int j = 0; j < TN; ++j) acc[i][j] = fma(x[i], y[j], acc[i][j]);
            }
        }
    }
    if (worker) {
        double* w = a.W + (long long)blockIdx.x * a.M * a.N;
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
            for (int j = 0; j < TN; ++j)
                if (m0 + i < a.M && n0 + j < a.N) w[(m0 + i) + (long long)(n0 + j) * a.M] = acc[i][j];
    }
}

template <int TM, int TN>
void launch_khuge(jues_ctx* ctx, const KHugeArgs& a, int ctas) {
    const int Mp = (a.M + TM - 1) / TM * TM, Np = (a.N + TN - 1) / TN * TN;
    const size_t smem = (size_t)kKC * (Mp + Np) * 8;
    if (ctx->smem_attr_done.insert((const void*)skinny_khuge_kernel<TM, TN>).second)
        JUES_CUDA(cudaFuncSetAttribute(skinny_khuge_kernel<TM, TN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    skinny_khuge_kernel<TM, TN><<<ctas, 256, smem, ctx->stream>>>(a);
    JUES_CUDA(cudaGetLastError());
}

}  // namespace

bool skinny_gemm(jues_ctx* ctx, const GemmCall& g) {
    static const bool off = getenv("JUES_B200_NO_SKINNY") != nullptr;
    if (off || g.batch != 1 || g.force_cfg >= 0) return false;
    const double M = (double)g.M, N = (double)g.N, K = (double)g.K;
    const double intensity = 2.0 * M * N * K / (8.0 * (M * K + K * N + M * N));
    // ---- K huge, M and N small ('N','T': both operands contiguous along their small index) -----------------
    if (!g.transA && g.transB && g.M <= 128 && g.N <= 128 && g.K >= 8192 && intensity < 13.0) {
        const int sms = ctx->sm_count > 0 ? ctx->sm_count : 148;
        int ctas = (int)std::min<int64_t>(2 * sms, (g.K + 4 * kKC - 1) / (4 * kKC));
        KHugeArgs a;
        a.M = (int)g.M; a.N = (int)g.N; a.K = g.K;
        a.kper = ((g.K + ctas - 1) / ctas + kKC - 1) / kKC * kKC;
        ctas = (int)((g.K + a.kper - 1) / a.kper);
        a.A = g.A; a.lda = g.lda; a.B = g.B; a.ldb = g.ldb;
        DBuf work(ctx, (size_t)ctas * g.M * g.N);
        a.W = work.p;
        const int64_t big = std::max(g.M, g.N), small = std::min(g.M, g.N);
        if (big <= 32) launch_khuge<2, 2>(ctx, a, ctas);
        else if (small <= 32) launch_khuge<4, 4>(ctx, a, ctas);
        else launch_khuge<8, 8>(ctx, a, ctas);
        ctx->stats.aux_launches += 1;
        splitk_reduce(ctx, work.p, ctas, g.M, g.N, 1, g.alpha, g.beta, g.C, g.ldc, 0);
        return true;
    }
    if (g.K > 4096 || intensity >= 10.0) return false;
    // ---- M huge, N small:  C(m, n) = sum_k A(m, k) B(k, n) ---------------------------------------------------
    if (g.N <= 256 && g.M >= 2048 && g.M >= 16 * g.N) {
        RowsArgs a;
        a.R = g.M; a.K = (int)g.K; a.nc = (int)g.N;
        a.L = g.A; a.ldl = g.lda;
        a.S = g.B; a.ssk = g.transB ? g.ldb : 1; a.ssc = g.transB ? 1 : g.ldb;
        a.O = g.C; a.sor = 1; a.soc = g.ldc;
        a.alpha = g.alpha; a.beta = g.beta;
        rows_product(ctx, a, g.transA);
        return true;
    }
    // ---- N huge, M small:  C^T(n, m) = sum_k B^T(n, k) A^T(k, m) ------------------------------------------------
    if (g.M <= 256 && g.N >= 2048 && g.N >= 16 * g.M) {
        RowsArgs a;
        a.R = g.N; a.K = (int)g.K; a.nc = (int)g.M;
        a.L = g.B; a.ldl = g.ldb;
        a.S = g.A; a.ssk = g.transA ? 1 : g.lda; a.ssc = g.transA ? g.lda : 1;
        a.O = g.C; a.sor = g.ldc; a.soc = 1;
        a.alpha = g.alpha; a.beta = g.beta;
        rows_product(ctx, a, !g.transB);      // B stored K x N: k-contiguous rows of B^T
        return true;
    }
    return false;
}

}  // namespace jues
