// FP64 GEMM for sm_100a: DMMA (mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4) register tiles fed by
// TMA (cp.async.bulk.tensor, 128B swizzle) through an mbarrier full/empty ring, with a
// dedicated producer warp.  tcgen05 has no FP64 kind, so the FP64 tensor path on Blackwell is the
// warp-level DMMA; all wider PTX f64 shapes (m16n8k4/8/16) lower to the same DMMA.8x8x4 on sm_100a.
//
// Operand layouts in shared memory (one pipeline stage holds a BMx16 slice of op(A) and a BNx16
// slice of op(B)^T; BK = 16 doubles = one 128-byte swizzle row):
//   KC ("K-contiguous": A stored KxM, i.e. transA='T'; B stored KxN, i.e. transB='N')
//       one TMA box {16 k, R rows}; element (r,k) at  r*128 + (((k>>1) ^ (r&7))<<4) + (k&1)*8
//   RC ("row-contiguous": A stored MxK, transA='N'; B stored NxK, transB='T')
//       R/16 TMA boxes {16 r, 16 k} of 2 KB; element (r,k) at
//       (r>>4)*2048 + k*128 + ((((r&15)>>1) ^ (k&7))<<4) + (r&1)*8
// DMMA wants, from lane (g = lane>>2, t = lane&3), A[row g][k t] and B[k t][col g].  The mapping of
// "MMA row/col g" to tile rows and of "MMA k t" to tile k is free as long as A, B and the
// accumulator agree, and is chosen so that every LDS.64 is bank-conflict-free for BOTH layouts:
//   k(t, s)  = 2t + ((t&1) ^ s)     s = 0,1 : the two DMMAs that cover an 8-k block
//                                    ({0,3,4,7} and {1,2,5,6})
//   row(g)   = g                     for an RC operand
//   row(g)   = {0,1,4,5,2,3,6,7}[g]  for a KC operand
// (verified exhaustively by tests/test_smem_layout.py).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace jues {
namespace gemm {

constexpr int BK = 16;  // doubles per pipeline stage along K (= 128 B)

struct Params {
    int M, N, K;
    int tilesM, tilesN;
    long long tiles_per_batch;
    long long total_tiles;   // tiles_per_batch * batch * ksplit; CTAs loop over them (persistent)
    int batch;
    int ksplit;         // split-K factor (1 = off); split z handles k-stages [z*kt_per_split, ...)
    int kt_per_split;
    long long strideSplit;  // element stride between split slices of C (workspace) when ksplit > 1
    int bmulA, bmulB;   // 0: operand is shared by all batches (batch stride 0), 1: batched
    int raster_n_fast;  // 1: consecutive CTAs walk N first (A tile shared), 0: M first
    double* C;
    long long ldc;
    long long strideC;
    double alpha, beta;
    // C = alpha*acc + beta*Cin: Cin has the layout of C (ldc, strideC); nullptr = C itself.  Lets a product
    // start from a static tensor without a copy pass (WJ = <mj|eb> + ..., WE = -<mb|je> + ... in the CC sweep).
    const double* Cin;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok;
}
// Bounded wait: a mis-programmed pipeline traps (-> CUDA error on the host) instead of hanging
// the GPU.  ~10 s at 2 GHz.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 20000000000LL) __trap();
    }
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar,
                                            int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ double lds64(uint32_t addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c0), "+d"(c1)
        : "d"(a), "d"(b));
}

// byte offset of element (r, k) of an operand tile, r in [0,R), k in [0,16)
template <bool KC>
__host__ __device__ __forceinline__ int frag_off(int r, int k) {
    if (KC) return r * 128 + ((((k >> 1) ^ (r & 7)) << 4) | ((k & 1) << 3));
    return (r >> 4) * 2048 + k * 128 + (((((r & 15) >> 1) ^ (k & 7)) << 4) | ((r & 1) << 3));
}
template <bool KC>
__host__ __device__ __forceinline__ int frag_row(int g) {  // MMA row/col g -> row inside an 8-row tile
    if (KC) return (g & 1) | ((g & 2) << 1) | ((g & 4) >> 1);
    return g;
}
__host__ __device__ __forceinline__ int frag_k(int t, int s) { return 2 * t + ((t & 1) ^ s); }

template <int BM, int BN, int STAGES>
struct SmemLayout {
    static constexpr int A_BYTES = BM * BK * 8;
    static constexpr int B_BYTES = BN * BK * 8;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
    static constexpr int TOTAL = BAR_OFF + 2 * STAGES * 8 + 1024;  // + alignment slack
};

template <bool A_KC, bool B_KC, int BM, int BN, int WM, int WN, int STAGES>
__global__ void __launch_bounds__((BM / WM) * (BN / WN) * 32 + 32, 1)
dgemm_tma_dmma(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
               const Params p) {
    constexpr int WARPS_M = BM / WM;
    constexpr int WARPS_N = BN / WN;
    constexpr int NCONS = WARPS_M * WARPS_N;  // consumer warps
    constexpr int TI = WM / 8;                // 8x8 accumulator tiles per warp along M
    constexpr int TJ = WN / 8;
    using L = SmemLayout<BM, BN, STAGES>;
    static_assert(WM % 16 == 0 && WN % 16 == 0, "warp tile must be a multiple of 16");
    static_assert(BM % 16 == 0 && BN % 16 == 0 && BM <= 256 && BN <= 256, "bad CTA tile");

    extern __shared__ unsigned char smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_full = smem_base + L::BAR_OFF;
    const uint32_t bar_empty = bar_full + STAGES * 8;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    // Persistent CTAs: CTA b works on tiles b, b + gridDim.x, ...  The TMA producer runs ahead across
    // tile boundaries (one continuous stage ring), so the loads of the next tile are in flight while the
    // consumers store the current one: no per-tile launch, barrier-init or first-load latency.
    const int KT_all = (p.K + BK - 1) / BK;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, NCONS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == NCONS) {
        // ================================ TMA producer ==========================================
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB) : "memory");
            unsigned it = 0;  // stage counter over all tiles of this CTA
            for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                const long long bzs = tile / p.tiles_per_batch;  // batch * ksplit + split
                const int rt = (int)(tile - bzs * p.tiles_per_batch);
                const int bz = (int)(bzs / p.ksplit);
                const int zs = (int)(bzs - (long long)bz * p.ksplit);
                int tm, tn;
                if (p.raster_n_fast) { tm = rt / p.tilesN; tn = rt - tm * p.tilesN; }
                else { tn = rt / p.tilesM; tm = rt - tn * p.tilesM; }
                const int m0 = tm * BM, n0 = tn * BN;
                const int kt_begin = zs * p.kt_per_split;
                const int KT = min(KT_all - kt_begin, p.kt_per_split);
                for (int kt = 0; kt < KT; ++kt, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (uint32_t)((it / STAGES) & 1);
                    mbar_wait(bar_empty + 8 * s, ph ^ 1u);
                    const uint32_t full = bar_full + 8 * s;
                    mbar_expect_tx(full, (uint32_t)L::STAGE_BYTES);
                    const uint32_t sa = smem_base + s * L::STAGE_BYTES;
                    const uint32_t sb = sa + L::A_BYTES;
                    const int k0 = (kt_begin + kt) * BK;
                    if (A_KC) {
                        tma_load_3d(sa, &mapA, full, k0, m0, bz * p.bmulA);
                    } else {
#pragma unroll
                        for (int q = 0; q < BM / 16; ++q)
                            tma_load_3d(sa + q * 2048, &mapA, full, m0 + 16 * q, k0, bz * p.bmulA);
                    }
                    if (B_KC) {
                        tma_load_3d(sb, &mapB, full, k0, n0, bz * p.bmulB);
                    } else {
#pragma unroll
                        for (int q = 0; q < BN / 16; ++q)
                            tma_load_3d(sb + q * 2048, &mapB, full, n0 + 16 * q, k0, bz * p.bmulB);
                    }
                }
            }
        }
        return;
    }

    // ================================== DMMA consumers ==========================================
    const int wm = warp % WARPS_M;
    const int wn = warp / WARPS_M;
    const int g = lane >> 2;
    const int t = lane & 3;
    const int ra = frag_row<A_KC>(g);
    const int rb = frag_row<B_KC>(g);

    // byte offsets of this lane's fragment element for (s, kb, i&1); tile i adds (i>>1)*2048
    int offA[2][2][2], offB[2][2][2];
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
        for (int kb = 0; kb < 2; ++kb)
#pragma unroll
            for (int ip = 0; ip < 2; ++ip) {
                const int k = kb * 8 + frag_k(t, s);
                offA[s][kb][ip] = frag_off<A_KC>(wm * WM + ip * 8 + ra, k);
                offB[s][kb][ip] = L::A_BYTES + frag_off<B_KC>(wn * WN + ip * 8 + rb, k);
            }

    const double alpha = p.alpha, beta = p.beta;
    unsigned it = 0;
    for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const long long bzs = tile / p.tiles_per_batch;
        const int rt = (int)(tile - bzs * p.tiles_per_batch);
        const int bz = (int)(bzs / p.ksplit);
        const int zs = (int)(bzs - (long long)bz * p.ksplit);
        int tm, tn;
        if (p.raster_n_fast) { tm = rt / p.tilesN; tn = rt - tm * p.tilesN; }
        else { tn = rt / p.tilesM; tm = rt - tn * p.tilesM; }
        const int m0 = tm * BM, n0 = tn * BN;
        const int kt_begin = zs * p.kt_per_split;
        const int KT = min(KT_all - kt_begin, p.kt_per_split);

        double acc[TI][TJ][2];
#pragma unroll
        for (int i = 0; i < TI; ++i)
#pragma unroll
            for (int j = 0; j < TJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

        // Release protocol.  A stage may go back to the TMA producer only when every LDS that reads it
        // has RETURNED its data.  Neither `asm volatile` nor the release semantics of mbarrier.arrive
        // make ptxas keep the DMMAs (which are what waits for the loaded registers) in front of the
        // arrive: it sinks them below it, so an arrive at the end of the k-step is issued right behind
        // the youngest LDS of the stage.  Measured on B200 (profiles/gemm_release_race_r02.md): those
        // youngest loads then occasionally return the NEXT fill of the stage -- 8x16 blocks of C wrong in
        // ~1 of 10^4 tiles on a cold GPU.  So the arrive for stage s is issued at the top of the NEXT
        // k-step, behind the spin-wait on the next full barrier: a control-flow boundary the DMMAs of
        // step s cannot cross, and by then they have all issued, i.e. consumed every loaded register.
        int s_prev = -1;
        for (int kt = 0; kt < KT; ++kt, ++it) {
            const int s = it % STAGES;
            const uint32_t ph = (uint32_t)((it / STAGES) & 1);
            mbar_wait(bar_full + 8 * s, ph);
            if (s_prev >= 0) {
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_empty + 8 * s_prev);
            }
            s_prev = s;
            const uint32_t st = smem_base + s * L::STAGE_BYTES;
#pragma unroll
            for (int kb = 0; kb < 2; ++kb) {
#pragma unroll
                for (int ss = 0; ss < 2; ++ss) {
                    double a[TI], b[TJ];
#pragma unroll
                    for (int i = 0; i < TI; ++i)
                        a[i] = lds64(st + offA[ss][kb][i & 1] + (i >> 1) * 2048);
#pragma unroll
                    for (int j = 0; j < TJ; ++j)
                        b[j] = lds64(st + offB[ss][kb][j & 1] + (j >> 1) * 2048);
#pragma unroll
                    for (int i = 0; i < TI; ++i)
#pragma unroll
                        for (int j = 0; j < TJ; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
                }
            }
        }

        // ================================= epilogue of this tile ================================
        const long long c_off = (long long)bz * p.strideC + (long long)zs * p.strideSplit;
        double* __restrict__ Cb = p.C + c_off;
        const double* Sb = (p.Cin ? p.Cin : p.C) + c_off;      // where the old values come from
        // Accumulating epilogue: a column's TI old values are requested back to back, and (where the
        // accumulators leave room under the 168-register limit of a 9-warp CTA) the old values of column q+1
        // BEFORE column q is stored.  Written as load, store, load, ... every load waits for a full memory
        // round trip behind the previous store (the compiler must assume the stores alias) -- measured
        // +50 us on a 2000^3 product and 3x on the skinny accumulating products of the sweep.
        constexpr bool PIPE = TI * TJ * 2 + 2 * TI <= 56;
        if (PIPE && beta != 0.0) {
            double old[TI], nxt[TI];
            auto column = [&](int q) -> long long {      // element offset of column q, -1 beyond N
                const int n = n0 + wn * WN + (q >> 1) * 8 + frag_row<B_KC>(2 * t + (q & 1));
                return n < p.N ? (long long)n * p.ldc : -1;
            };
            auto fetch = [&](long long off, double* dst) {
#pragma unroll
                for (int i = 0; i < TI; ++i) {
                    const int m = m0 + wm * WM + i * 8 + ra;
                    dst[i] = (off >= 0 && m < p.M) ? __ldcg(Sb + off + m) : 0.0;
                }
            };
            fetch(column(0), old);
#pragma unroll
            for (int q = 0; q < 2 * TJ; ++q) {
                const long long off = column(q);
                if (q + 1 < 2 * TJ) fetch(column(q + 1), nxt);
                if (off >= 0) {
                    double* col = Cb + off;
#pragma unroll
                    for (int i = 0; i < TI; ++i) {
                        const int m = m0 + wm * WM + i * 8 + ra;
                        if (m < p.M) col[m] = alpha * acc[i][q >> 1][q & 1] + beta * old[i];
                    }
                }
#pragma unroll
                for (int i = 0; i < TI; ++i) old[i] = nxt[i];
            }
        } else {
#pragma unroll
            for (int j = 0; j < TJ; ++j) {
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const int n = n0 + wn * WN + j * 8 + frag_row<B_KC>(2 * t + c);
                    if (n < p.N) {
                        double* col = Cb + (long long)n * p.ldc;
                        const double* src = Sb + (long long)n * p.ldc;
                        if (beta != 0.0) {
                            // the large accumulator tiles have no registers to spare: half / a quarter of a column at a time
                            constexpr int CH = (TI * TJ >= 32) ? (TI + 3) / 4 : (TI * TJ >= 28) ? (TI + 1) / 2 : TI;
#pragma unroll
                            for (int i0 = 0; i0 < TI; i0 += CH) {
                                double old[CH];
#pragma unroll
                                for (int i = 0; i < CH; ++i) {
                                    const int m = m0 + wm * WM + (i0 + i) * 8 + ra;
                                    old[i] = (i0 + i < TI && m < p.M) ? __ldcg(src + m) : 0.0;
                                }
#pragma unroll
                                for (int i = 0; i < CH; ++i) {
                                    const int m = m0 + wm * WM + (i0 + i) * 8 + ra;
                                    if (i0 + i < TI && m < p.M) col[m] = alpha * acc[(i0 + i) % TI][j][c] + beta * old[i];
                                }
                            }
                        } else {
#pragma unroll
                            for (int i = 0; i < TI; ++i) {
                                const int m = m0 + wm * WM + i * 8 + ra;
                                if (m < p.M) col[m] = alpha * acc[i][j][c];
                            }
                        }
                    }
                }
            }
        }
        // the last stage of the tile: its loads are certainly consumed once the accumulators have been
        // stored (the stores depend on every DMMA, and the arrive cannot move above them: "memory" clobber)
        if (s_prev >= 0) {
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_empty + 8 * s_prev);
        }
    }
}

}  // namespace gemm
}  // namespace jues
