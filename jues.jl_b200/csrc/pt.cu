// Perturbative triples correction E(T) (PerturbativeTriples.jl:35-138) on the device.
//
// The reference loops over occupied triples i >= j >= k and builds, per triple, the v^3 arrays
//   W[a,b,c] = sum over the six simultaneous permutations of (i,j,k)/(a,b,c) of
//              sum_d Vvvvo[b,d,a,i] T2[k,j,c,d] - sum_l Vvooo[c,k,j,l] T2[i,l,a,b]          (:96-101)
//   V[a,b,c] = W[a,b,c] + Vvovo[b,j,c,k] T1[i,a] + Vvovo[a,i,c,k] T1[j,b] + Vvovo[a,i,b,j] T1[k,c]  (:103)
// with 12 small @tensoropt contractions, then a scalar a >= b >= c loop nest (:117-131).
//
// Here (tests/pt_model.py is the numpy statement of the same algorithm):
//   * one "X block" per ordered triple (p,q,r), with the two contractions concatenated along K = v + o:
//         X(p,q,r)[(a,b),c] = sum_kap Acat[(a,b),p,kap] B[c,kap,q,r]
//         Acat = [ <p d|ab> | -T2[p,l,a,b] ],   B = [ T2[r,q,c,d] | <qr|lc> ]
//     = ONE launch of the TMA + DMMA GEMM with NO operand permutation per block: the layouts
//     Acat[a,b,p,kap], Bq[c,q,kap,r], Br[c,r,kap,q] (built once) make every operand a strided matrix;
//   * for a fixed pair (i,j) the six blocks of ALL k <= j are produced by 6 launches: consecutive q
//     (or r) are consecutive rows of Bq (Br), so N = v*nb; consecutive p are consecutive rows of Acat,
//     so M = v^2*nb;
//   * one kernel assembles W and V from the six blocks (index permutations + rank-1 terms), a second
//     one evaluates the energy expression on a >= b >= c and reduces it deterministically
//     (block tree -> one slot per batch; the slots are summed in a fixed order at the end).
// Flops: 6 * 2 v^3 (v + o) per triple, o(o+1)(o+2)/6 triples -- the same count as the reference.
#include "pt.h"
#include "dgemm.h"
#include "dist.h"

#include <algorithm>

namespace jues {


namespace {

template <int THREADS>
__device__ __forceinline__ double block_sum(double v) {
    __shared__ double red[THREADS / 32];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double r = 0.0;
    if (warp == 0) {
        r = (lane < THREADS / 32) ? red[lane] : 0.0;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) r += __shfl_down_sync(0xffffffffu, r, off);
    }
    __syncthreads();
    return r;  // valid in thread 0
}

// W[kk][a,b,c] and V[kk][a,b,c] for the triples (i, j, k0+kk), kk < kb, from the six X blocks
// X[s][kk][.,.,.] (s = 0..5 in the order of the header comment).
__global__ void pt_assemble_kernel(const double* __restrict__ X, const double* __restrict__ Vv,
                                   const double* __restrict__ t1, double* __restrict__ W,
                                   double* __restrict__ V, int o, int v, int i, int j, int k0, int kb) {
    const long long v2 = (long long)v * v, v3 = v2 * v;
    const long long total = v3 * kb;
    const long long bs = v3 * kb;   // stride between the six block families
    for (long long L = (long long)blockIdx.x * blockDim.x + threadIdx.x; L < total;
         L += (long long)gridDim.x * blockDim.x) {
        long long r = L;
        const int a = (int)(r % v); r /= v;
        const int b = (int)(r % v); r /= v;
        const int c = (int)(r % v);
        const int kk = (int)(r / v);
        const int k = k0 + kk;
        const double* x = X + (long long)kk * v3;    // families batched over q or r: [kk][.,.,.]
        const double* xp = X + (long long)kk * v2;   // families batched over p:      [.,.,kk,.]
        const double w = x[a + (long long)v * b + v2 * c]               // X(i,j,k)[a,b,c]
                       + x[bs + a + (long long)v * c + v2 * b]          // X(i,k,j)[a,c,b]
                       + xp[2 * bs + c + (long long)v * a + v2 * kb * b]   // X(k,i,j)[c,a,b], layout [.,.,kk,.]
                       + xp[3 * bs + c + (long long)v * b + v2 * kb * a]   // X(k,j,i)[c,b,a], layout [.,.,kk,.]
                       + x[4 * bs + b + (long long)v * c + v2 * a]      // X(j,k,i)[b,c,a]
                       + x[5 * bs + b + (long long)v * a + v2 * c];     // X(j,i,k)[b,a,c]
        const double vv = w
            + Vv[b + (long long)v * c + v2 * (j + (long long)o * k)] * t1[i + (long long)o * a]
            + Vv[a + (long long)v * c + v2 * (i + (long long)o * k)] * t1[j + (long long)o * b]
            + Vv[a + (long long)v * b + v2 * (i + (long long)o * j)] * t1[k + (long long)o * c];
        W[L] = w;
        V[L] = vv;
    }
}

// sum over kk and a >= b >= c of  E(a,b,c) (2 - d_ij - d_jk) / (Dd (1 + d_ab + d_bc))   (:117-131)
__global__ void pt_energy_kernel(const double* __restrict__ W, const double* __restrict__ V,
                                 const double* __restrict__ eo, const double* __restrict__ ev, int v, int i,
                                 int j, int k0, int kb, double* __restrict__ partial) {
    const long long v2 = (long long)v * v, v3 = v2 * v;
    const long long total = v3 * kb;
    double acc = 0.0;
    for (long long L = (long long)blockIdx.x * blockDim.x + threadIdx.x; L < total;
         L += (long long)gridDim.x * blockDim.x) {
        long long r = L;
        const int a = (int)(r % v); r /= v;
        const int b = (int)(r % v); r /= v;
        const int c = (int)(r % v);
        const int kk = (int)(r / v);
        if (a < b || b < c) continue;
        const int k = k0 + kk;
        const double* w = W + (long long)kk * v3;
        const double* u = V + (long long)kk * v3;
        const long long abc = a + (long long)v * b + v2 * c, acb = a + (long long)v * c + v2 * b;
        const long long bac = b + (long long)v * a + v2 * c, bca = b + (long long)v * c + v2 * a;
        const long long cab = c + (long long)v * a + v2 * b, cba = c + (long long)v * b + v2 * a;
        const double Wabc = w[abc], Wacb = w[acb], Wbac = w[bac], Wbca = w[bca], Wcab = w[cab], Wcba = w[cba];
        const double Vabc = u[abc], Vacb = u[acb], Vbac = u[bac], Vbca = u[bca], Vcab = u[cab], Vcba = u[cba];
        const double Xs = Wabc * Vabc + Wacb * Vacb + Wbac * Vbac + Wbca * Vbca + Wcab * Vcab + Wcba * Vcba;
        const double Y = Vabc + Vbca + Vcab;
        const double Z = Vacb + Vbac + Vcba;
        const double E = (Y - 2.0 * Z) * (Wabc + Wbca + Wcab) + (Z - 2.0 * Y) * (Wacb + Wbac + Wcba) + 3.0 * Xs;
        const double occ = 2.0 - (i == j ? 1.0 : 0.0) - (j == k ? 1.0 : 0.0);
        const double Dd = eo[i] + eo[j] + eo[k] - ev[a] - ev[b] - ev[c];
        const double sym = 1.0 + (a == b ? 1.0 : 0.0) + (b == c ? 1.0 : 0.0);
        acc += E * occ / (Dd * sym);
    }
    const double r = block_sum<256>(acc);
    if (threadIdx.x == 0) partial[blockIdx.x] = r;
}


__global__ void pt_slot_reduce_kernel(const double* __restrict__ partial, int n, double* __restrict__ out) {
    double acc = 0.0;
    for (int q = threadIdx.x; q < n; q += blockDim.x) acc += partial[q];
    const double r = block_sum<256>(acc);
    if (threadIdx.x == 0) out[0] = r;
}

// The X blocks of `nb` consecutive triples: (p,q,r) = (p0 + ps*n, q0 + qs*n, r0 + rs*n), n < nb,
// exactly one of ps, qs, rs being 1.  One GEMM, K = v + o.
//   qs or rs == 1:  out[(a,b),(c,n)] = out[n][(a,b),c]   (N = v*nb: rows (c,q) of Bq resp. (c,r) of Br)
//   ps == 1:        out[(a,b,n),c]                        (M = v^2*nb: rows (a,b,p) of Acat)
void x_blocks(jues_ctx* ctx, const PtInputs& in, int64_t p0, int ps, int64_t q0, int qs, int64_t r0, int rs,
              int64_t nb, double* out) {
    const int64_t o = in.o, v = in.v, v2 = v * v, K = v + o;
    GemmCall g;
    g.K = K;
    g.transA = false; g.A = in.Acat + p0 * v2; g.lda = v2 * o;
    g.transB = true;  g.ldb = v * o;
    // Bq[c,q,kap,r] / Br[c,r,kap,q]: the matrix [c (and the walked index), kap] of a fixed (q,r)
    g.B = rs ? in.Br + v * r0 + v * o * K * q0 : in.Bq + v * q0 + v * o * K * r0;
    g.M = ps ? v2 * nb : v2;
    g.N = ps ? v : v * nb;
    g.C = out; g.ldc = g.M;
    g.alpha = 1.0; g.beta = 0.0;
    dgemm(ctx, g);
}

}  // namespace

void pt_build_operands(jues_ctx* ctx, int64_t o, int64_t v, const double* OAp, const double* T2,
                       const double* ooov, double* acat, double* bq, double* br) {
    const int64_t K = v + o;
    const Ten T(const_cast<double*>(T2), o, o, v, v), O3(const_cast<double*>(ooov), o, o, o, v);
    // Acat[a,b,p,kap]: kap < v is OAp itself (kap is the slowest index: one contiguous block)
    if (OAp)
        JUES_CUDA(cudaMemcpyAsync(acat, OAp, (size_t)(v * v * o * v) * sizeof(double), cudaMemcpyDeviceToDevice,
                                  ctx->stream));
    permute_axpby(ctx, -1.0, T, "plab", 0.0, Ten(acat + v * v * o * v, v, v, o, o), "abpl");
    // Bq[c,q,kap,r] and Br[c,r,kap,q]: permute into dense temporaries, then place the two kap ranges
    const int64_t ddq[4] = {v, o, K, o};
    DTen t1(ctx, v, o, v, o), t2(ctx, v, o, o, o);
    const int64_t e1[4] = {v, o, v, o}, e2[4] = {v, o, o, o};
    permute_axpby(ctx, 1.0, T, "rqck", 0.0, t1, "cqkr");
    block_copy(ctx, t1.p(), e1, bq, ddq, e1);
    permute_axpby(ctx, 1.0, O3, "qrlc", 0.0, t2, "cqlr");
    block_copy(ctx, t2.p(), e2, bq + v * o * v, ddq, e2);
    permute_axpby(ctx, 1.0, T, "rqck", 0.0, t1, "crkq");
    block_copy(ctx, t1.p(), e1, br, ddq, e1);
    permute_axpby(ctx, 1.0, O3, "qrlc", 0.0, t2, "crlq");
    block_copy(ctx, t2.p(), e2, br + v * o * v, ddq, e2);
}

double pt_dev(jues_ctx* ctx, const PtInputs& in, double* scratch, size_t scratch_elems) {
    const int64_t o = in.o, v = in.v, nocc = in.nocc;
    JUES_REQUIRE(o > 0 && v > 0 && nocc > 0 && nocc <= o, "(T): bad extents");
    JUES_REQUIRE((o & 1) == 0 && (v & 1) == 0, "(T): internal extents must be even");
    JUES_REQUIRE(in.Acat && in.Bq && in.Br && in.Vv && in.t1 && in.eo && in.ev, "(T): null input");
    const int64_t v3 = v * v * v;
    // batch as many k as memory allows: 8 v^3-sized arrays per triple (6 X blocks, W, V)
    int64_t kbmax = 0;
    DBuf work;
    double* base = nullptr;
    const int64_t ks = (int64_t)(scratch_elems / (size_t)(8 * v3));
    if (scratch && ks >= std::min<int64_t>(nocc, 4)) {
        kbmax = std::min<int64_t>(ks, nocc);
        base = scratch;
    } else {
        size_t free_b = 0, total_b = 0;
        cudaMemGetInfo(&free_b, &total_b);
        free_b += ctx->big_cached_bytes;
        kbmax = (int64_t)(0.6 * (double)free_b / (8.0 * 8.0 * (double)v3));
        kbmax = std::max<int64_t>(1, std::min<int64_t>(kbmax, nocc));
    }
    if (getenv("JUES_B200_PT_KB")) {                                                       // testing hook
        const int64_t kb_env = std::max(1, atoi(getenv("JUES_B200_PT_KB")));
        kbmax = base ? std::min(kbmax, kb_env) : kb_env;      // never beyond what the scratch block holds
    }
    if (!base) {
        TraceTimer tal(ctx, "pt.alloc");
        work.alloc(ctx, (size_t)(8 * kbmax * v3));
        base = work.p;
    }
    struct View { double* p; };
    const View X{base}, W{base + 6 * kbmax * v3}, V{base + 7 * kbmax * v3};
    const int64_t nchunk_max = (nocc + kbmax - 1) / kbmax;
    const int64_t nslots = nocc * (nocc + 1) / 2 * nchunk_max;
    DBuf slots(ctx, (size_t)nslots + 1);
    slots.zero();
    int64_t slot = 0, pair = 0;
    for (int64_t i = 0; i < nocc; ++i) {
        for (int64_t j = 0; j <= i; ++j, ++pair) {
            if (pair % ctx->nranks != ctx->rank) continue;
            for (int64_t k0 = 0; k0 <= j; k0 += kbmax) {
                const int64_t kb = std::min(kbmax, j + 1 - k0);
                const int64_t fam = kb * v3;   // one family of X blocks
                TraceTimer* tg = new TraceTimer(ctx, "pt.gemm");
                x_blocks(ctx, in, i, 0, j, 0, k0, 1, kb, X.p);             // X(i,j,k)
                x_blocks(ctx, in, i, 0, k0, 1, j, 0, kb, X.p + fam);       // X(i,k,j)
                x_blocks(ctx, in, k0, 1, i, 0, j, 0, kb, X.p + 2 * fam);   // X(k,i,j)
                x_blocks(ctx, in, k0, 1, j, 0, i, 0, kb, X.p + 3 * fam);   // X(k,j,i)
                x_blocks(ctx, in, j, 0, k0, 1, i, 0, kb, X.p + 4 * fam);   // X(j,k,i)
                x_blocks(ctx, in, j, 0, i, 0, k0, 1, kb, X.p + 5 * fam);   // X(j,i,k)
                delete tg;
                TraceTimer ta(ctx, "pt.assemble+energy");
                const int grid = ew_grid(ctx, (size_t)fam, 256);
                pt_assemble_kernel<<<grid, 256, 0, ctx->stream>>>(X.p, in.Vv, in.t1, W.p, V.p, (int)o, (int)v,
                                                                  (int)i, (int)j, (int)k0, (int)kb);
                JUES_CUDA(cudaGetLastError());
                const int rgrid = (int)std::min<long long>(grid, (long long)ctx->red_cap - 4);
                pt_energy_kernel<<<rgrid, 256, 0, ctx->stream>>>(W.p, V.p, in.eo, in.ev, (int)v, (int)i, (int)j,
                                                                 (int)k0, (int)kb, ctx->red_dev);
                JUES_CUDA(cudaGetLastError());
                pt_slot_reduce_kernel<<<1, 256, 0, ctx->stream>>>(ctx->red_dev, rgrid, slots.p + slot);
                JUES_CUDA(cudaGetLastError());
                ctx->stats.aux_launches += 3;
                ++slot;
            }
        }
    }
    pt_slot_reduce_kernel<<<1, 256, 0, ctx->stream>>>(slots.p, (int)nslots, slots.p + nslots);
    JUES_CUDA(cudaGetLastError());
    ctx->stats.aux_launches++;
    double e = 0.0;
    JUES_CUDA(cudaMemcpyAsync(&e, slots.p + nslots, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    JUES_CUDA(cudaStreamSynchronize(ctx->stream));
    return all_reduce_scalar(ctx, e);
}

}  // namespace jues
