// Streaming kernels for two families of bandwidth-bound products of the sweep, where the TMA + DMMA tile
// kernel is 3-4x above the HBM floor whatever its configuration (profiles/r02/gemm_autotune.log: tile padding
// makes it DMMA-issue bound and few short tiles cannot keep enough bytes in flight):
//
//   * matrix-vector products  y[r] = alpha * sum_k L[k + r*ld] x[k] + beta * y[r]   (N == 1, op(A) = A^T):
//     the t1-contractions of the ov^3 / o^2v^2 integrals (RCCSD.jl:187-210, 248-259).  One warp per row,
//     16-byte loads along k, x in shared memory, shuffle reduction: one pass over L at HBM speed.
//   * K-huge products  C (M x N, both <= 128) = A (M x K) B^T (N x K), K >= 8192, both operands dense and
//     contiguous along their small index (Fae / Fmi / T1 residual terms: K = o v^2 or o^2 v).  K is split over
//     ~2 CTAs per SM; each streams its contiguous slices of A and B through a double-buffered cp.async ring
//     and accumulates a register micro-tile per thread; the partial tiles are summed by splitk_reduce in a
//     fixed order.
//
// Deterministic (fixed summation order, no atomics).  A first, general "one thread per row" family for
// M >> N, K products was measured slower than the tile kernel and is not kept
// (profiles/r02/gemm_list_c3_r02i.txt).
#include "dgemm.h"
#include "tensor_ops.h"

#include <algorithm>

namespace jues {

namespace {

// ---- matrix-vector ------------------------------------------------------------------------------------------
struct GemvArgs {
    long long R;            // rows
    int K;
    const double* L; long long ldl;   // L[k + r*ldl]
    const double* x; long long sx;    // x[k*sx]
    double* y; long long sy;          // y[r*sy]
    double alpha, beta;
};

template <bool VEC>
__global__ void __launch_bounds__(256) skinny_gemv_kernel(GemvArgs a) {
    extern __shared__ double xs[];    // [K]
    for (int k = threadIdx.x; k < a.K; k += blockDim.x) xs[k] = a.x[(long long)k * a.sx];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    for (long long r = (long long)blockIdx.x * wpb + warp; r < a.R; r += (long long)gridDim.x * wpb) {
        const double* row = a.L + r * a.ldl;
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
        if (VEC) {
            const double2* row2 = reinterpret_cast<const double2*>(row);
            const double2* x2 = reinterpret_cast<const double2*>(xs);
            const int K2 = a.K >> 1;
            int k = lane;
            for (; k + 96 < K2; k += 128) {     // four independent 16-byte loads in flight per lane
                const double2 v0 = row2[k], v1 = row2[k + 32], v2 = row2[k + 64], v3 = row2[k + 96];
                const double2 w0 = x2[k], w1 = x2[k + 32], w2 = x2[k + 64], w3 = x2[k + 96];
                s0 = fma(v0.x, w0.x, fma(v0.y, w0.y, s0));
                s1 = fma(v1.x, w1.x, fma(v1.y, w1.y, s1));
                s2 = fma(v2.x, w2.x, fma(v2.y, w2.y, s2));
                s3 = fma(v3.x, w3.x, fma(v3.y, w3.y, s3));
            }
            for (; k < K2; k += 32) {
                const double2 v = row2[k], w = x2[k];
                s0 = fma(v.x, w.x, fma(v.y, w.y, s0));
            }
        } else {
            for (int k = lane; k < a.K; k += 32) s0 = fma(row[k], xs[k], s0);
        }
        double s = (s0 + s1) + (s2 + s3);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) s += __shfl_down_sync(0xffffffffu, s, off);
        if (lane == 0) {
            double* o = a.y + r * a.sy;
            const double v = a.alpha * s;
            *o = a.beta == 0.0 ? v : v + a.beta * *o;
        }
    }
}

// y[r] = alpha * sum_k L[r + k*ldl] x[k] + beta * y[r]: the matrix is contiguous along r.  Block = 32 rows x 8
// k-groups: a warp reads 32 consecutive rows of one k (256 bytes), the eight groups stride through k, their
// partial sums are added in a fixed order through shared memory.
__global__ void __launch_bounds__(256) skinny_gemv_n_kernel(GemvArgs a) {
    extern __shared__ double xs[];    // [K] + [8][32]
    double* part = xs + a.K;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int k = threadIdx.x; k < a.K; k += blockDim.x) xs[k] = a.x[(long long)k * a.sx];
    __syncthreads();
    const long long r = (long long)blockIdx.x * 32 + tx;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    if (r < a.R) {
        const double* col = a.L + r;
        int k = ty;
        for (; k + 24 < a.K; k += 32) {
            const double v0 = col[(long long)k * a.ldl], v1 = col[(long long)(k + 8) * a.ldl];
            const double v2 = col[(long long)(k + 16) * a.ldl], v3 = col[(long long)(k + 24) * a.ldl];
            s0 = fma(v0, xs[k], s0); s1 = fma(v1, xs[k + 8], s1);
            s2 = fma(v2, xs[k + 16], s2); s3 = fma(v3, xs[k + 24], s3);
        }
        for (; k < a.K; k += 8) s0 = fma(col[(long long)k * a.ldl], xs[k], s0);
    }
    part[ty * 32 + tx] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    if (ty == 0 && r < a.R) {
        double s = 0.0;
#pragma unroll
        for (int q = 0; q < 8; ++q) s += part[q * 32 + tx];
        double* o = a.y + r * a.sy;
        const double v = a.alpha * s;
        *o = a.beta == 0.0 ? v : v + a.beta * *o;
    }
}

// ---- K huge, M and N small:  W[z] (M x N) = sum_{k in chunk z} A[:, k] B[:, k]^T --------------------------
struct KHugeArgs {
    int M, N;               // both even
    long long K, kper;      // k range per CTA (a multiple of kKC)
    const double* A;        // dense M x K (lda == M), 16-byte aligned
    const double* B;        // dense N x K (ldb == N), 16-byte aligned
    double* W;              // [M*N][ldz]: element e of CTA z at W[e*ldz + z] (slices contiguous per element)
    int ldz;
};

constexpr int kKC = 32;     // k-columns staged per step

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N_>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N_) : "memory"); }

// The slices A[:, k0 .. k0+kn) and B[:, k0 .. k0+kn) are contiguous runs of kn*M and kn*N doubles.
template <int TM, int TN, int THREADS>
__global__ void __launch_bounds__(THREADS) skinny_khuge_kernel(KHugeArgs a) {
    extern __shared__ __align__(16) double sm[];
    const int M = a.M, N = a.N;
    const int stage = kKC * (M + N);               // doubles per stage: [A chunk | B chunk]
    const int gm = (M + TM - 1) / TM, gn = (N + TN - 1) / TN;
    const int tid = threadIdx.x;
    const bool worker = tid < gm * gn;
    const int m0 = (tid % gm) * TM, n0 = (tid / gm) * TN;
    double acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.0;
    const long long kb = (long long)blockIdx.x * a.kper, ke = min(a.K, kb + a.kper);
    const int nsteps = (int)((ke - kb + kKC - 1) / kKC);

    auto issue = [&](int step) {
        const long long k0 = kb + (long long)step * kKC;
        const int kn = (int)min((long long)kKC, ke - k0);
        double* As = sm + (step & 1) * stage;
        double* Bs = As + kKC * M;
        const double* ga = a.A + k0 * M;
        const double* gb = a.B + k0 * N;
        const int na = kn * M / 2, nb = kn * N / 2;            // 16-byte pieces (M, N even)
        for (int e = tid; e < na; e += THREADS) cp_async16(As + 2 * e, ga + 2 * e);
        for (int e = tid; e < nb; e += THREADS) cp_async16(Bs + 2 * e, gb + 2 * e);
        // a short last step: the rest of the stage must read as zero
        for (int e = kn * M + tid; e < kKC * M; e += THREADS) As[e] = 0.0;
        for (int e = kn * N + tid; e < kKC * N; e += THREADS) Bs[e] = 0.0;
        cp_async_commit();
    };

    if (nsteps > 0) issue(0);
    for (int step = 0; step < nsteps; ++step) {
        if (step + 1 < nsteps) {
            issue(step + 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        if (worker) {
            const double* As = sm + (step & 1) * stage;
            const double* Bs = As + kKC * M;
#pragma unroll 4
            for (int kk = 0; kk < kKC; ++kk) {
                double x[TM], y[TN];
#pragma unroll
                for (int i = 0; i < TM; ++i) x[i] = (m0 + i < M) ? As[kk * M + m0 + i] : 0.0;
#pragma unroll
                for (int j = 0; j < TN; ++j) y[j] = (n0 + j < N) ? Bs[kk * N + n0 + j] : 0.0;
#pragma unroll
                for (int i = 0; i < TM; ++i)
#pragma unroll
                    for (int j = 0; j < TN; ++j) acc[i][j] = fma(x[i], y[j], acc[i][j]);
            }
        }
        __syncthreads();      // the stage is refilled two steps later
    }
    if (worker) {
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
            for (int j = 0; j < TN; ++j)
                if (m0 + i < M && n0 + j < N)
                    a.W[((long long)(m0 + i) + (long long)(n0 + j) * M) * a.ldz + blockIdx.x] = acc[i][j];
    }
}

// C[m + n*ldc] = alpha * sum_z W[e*ldz + z] + beta * C: one warp per element, lanes over z in a fixed order
__global__ void __launch_bounds__(256) skinny_slices_reduce_kernel(const double* __restrict__ W, int ldz, int nz,
                                                                   int M, int N, double alpha, double beta,
                                                                   double* __restrict__ C, long long ldc) {
    const int lane = threadIdx.x & 31;
    const long long e = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (e >= (long long)M * N) return;
    const double* w = W + e * ldz;
    double s = 0.0;
    for (int z = lane; z < nz; z += 32) s += w[z];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_down_sync(0xffffffffu, s, off);
    if (lane == 0) {
        double* c = C + (e % M) + (e / M) * ldc;
        *c = beta == 0.0 ? alpha * s : alpha * s + beta * *c;
    }
}

template <int TM, int TN, int THREADS>
void launch_khuge(jues_ctx* ctx, const KHugeArgs& a, int ctas) {
    const size_t smem = (size_t)2 * kKC * (a.M + a.N) * 8;
    if (ctx->smem_attr_done.insert((const void*)skinny_khuge_kernel<TM, TN, THREADS>).second)
        JUES_CUDA(cudaFuncSetAttribute(skinny_khuge_kernel<TM, TN, THREADS>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, 132 * 1024));
    skinny_khuge_kernel<TM, TN, THREADS><<<ctas, THREADS, smem, ctx->stream>>>(a);
    JUES_CUDA(cudaGetLastError());
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

bool skinny_gemm(jues_ctx* ctx, const GemmCall& g) {
    static const bool off = getenv("JUES_B200_NO_SKINNY") != nullptr;
    if (off || g.batch != 1 || g.force_cfg >= 0 || g.Cin != nullptr) return false;
    const int sms = ctx->sm_count > 0 ? ctx->sm_count : 148;
    // ---- matrix-vector: N == 1, A stored K x M (k-contiguous rows) ------------------------------------------
    if (g.N == 1 && g.transA && g.M >= 256 && g.K >= 64 && g.K <= 8192) {
        GemvArgs a;
        a.R = g.M; a.K = (int)g.K;
        a.L = g.A; a.ldl = g.lda;
        a.x = g.B; a.sx = g.transB ? g.ldb : 1;       // B is K x 1 (or 1 x K)
        a.y = g.C; a.sy = 1;
        a.alpha = g.alpha; a.beta = g.beta;
        const bool vec = aligned16(g.A) && (g.lda & 1) == 0 && (g.K & 1) == 0;
        const unsigned grid = (unsigned)std::min<int64_t>((g.M + 7) / 8, (int64_t)sms * 8);
        const size_t smem = (size_t)g.K * 8;
        if (ctx->smem_attr_done.insert((const void*)skinny_gemv_kernel<true>).second) {
            JUES_CUDA(cudaFuncSetAttribute(skinny_gemv_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
            JUES_CUDA(cudaFuncSetAttribute(skinny_gemv_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        }
        if (vec) skinny_gemv_kernel<true><<<grid, 256, smem, ctx->stream>>>(a);
        else skinny_gemv_kernel<false><<<grid, 256, smem, ctx->stream>>>(a);
        JUES_CUDA(cudaGetLastError());
        ctx->stats.aux_launches += 1;
        return true;
    }
    // ---- matrix-vector: N == 1, A stored M x K (contiguous along the rows) ---------------------------------
    if (g.N == 1 && !g.transA && g.M >= 2048 && g.K >= 16 && g.K <= 7900) {
        GemvArgs a;
        a.R = g.M; a.K = (int)g.K;
        a.L = g.A; a.ldl = g.lda;
        a.x = g.B; a.sx = g.transB ? g.ldb : 1;
        a.y = g.C; a.sy = 1;
        a.alpha = g.alpha; a.beta = g.beta;
        if (ctx->smem_attr_done.insert((const void*)skinny_gemv_n_kernel).second)
            JUES_CUDA(cudaFuncSetAttribute(skinny_gemv_n_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        skinny_gemv_n_kernel<<<(unsigned)((g.M + 31) / 32), 256, (size_t)(g.K + 256) * 8, ctx->stream>>>(a);
        JUES_CUDA(cudaGetLastError());
        ctx->stats.aux_launches += 1;
        return true;
    }
    // ---- K huge, M and N small ('N','T', dense operands) -----------------------------------------------------
    const double M = (double)g.M, N = (double)g.N, K = (double)g.K;
    const double intensity = 2.0 * M * N * K / (8.0 * (M * K + K * N + M * N));
    // (above ~1000 output elements the tile kernel with its narrow 32x128 tile and a 128-way split is faster:
    // 20 x 100 x 200000 in 79 us against 98 us, profiles/r02/session_r02n_epilogue_narrow_tiles.log)
    if (!g.transA && g.transB && g.M <= 128 && g.N <= 128 && g.M * g.N <= 1024 && g.K >= 8192 && intensity < 13.0 &&
        g.lda == g.M && g.ldb == g.N && (g.M & 1) == 0 && (g.N & 1) == 0 && aligned16(g.A) && aligned16(g.B)) {
        int ctas = (int)std::min<int64_t>(2 * sms, (g.K + 4 * kKC - 1) / (4 * kKC));
        KHugeArgs a;
        a.M = (int)g.M; a.N = (int)g.N; a.K = g.K;
        a.kper = ((g.K + ctas - 1) / ctas + kKC - 1) / kKC * kKC;
        ctas = (int)((g.K + a.kper - 1) / a.kper);
        a.A = g.A; a.B = g.B;
        a.ldz = (ctas + 1) & ~1;
        DBuf work(ctx, (size_t)a.ldz * g.M * g.N);
        a.W = work.p;
        if (std::max(g.M, g.N) <= 32) launch_khuge<2, 2, 256>(ctx, a, ctas);
        else launch_khuge<4, 4, 256>(ctx, a, ctas);
        skinny_slices_reduce_kernel<<<(unsigned)((g.M * g.N + 7) / 8), 256, 0, ctx->stream>>>(
            work.p, a.ldz, ctas, (int)g.M, (int)g.N, g.alpha, g.beta, g.C, g.ldc);
        JUES_CUDA(cudaGetLastError());
        ctx->stats.aux_launches += 2;
        return true;
    }
    return false;
}

}  // namespace jues
