// Element-wise, permutation and reduction kernels: the HBM-bound side of the path (index
// permutations, orbital-energy denominators, tau = T2 + t(x)t, energy reductions, split-K
// epilogue, synthetic ERI generator).  All reductions are deterministic (fixed two-pass tree).
#include "tensor_ops.h"
#include "api_util.h"
#include "synth_element.h"
#include <algorithm>

namespace jues {

namespace {

__device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ull;
    unsigned long long z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__global__ void fill_pattern_kernel(double* __restrict__ p, size_t n, unsigned long long seed) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        unsigned long long h = splitmix64(seed * 0x100000001B3ull ^ i);
        p[i] = 2.0 * ((double)(h >> 11) * (1.0 / 9007199254740992.0)) - 1.0;
    }
}

__global__ void fill_kernel(double* __restrict__ p, size_t n, double v) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) p[i] = v;
}

__global__ void axpby_kernel(size_t n, double a, const double* __restrict__ x, double b,
                             double* __restrict__ y) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    if (b == 0.0) {
        for (; i < n; i += stride) y[i] = a * x[i];
    } else {
        for (; i < n; i += stride) y[i] = a * x[i] + b * y[i];
    }
}

__global__ void lincomb2_kernel(size_t n, double a, const double* __restrict__ x1, double b,
                                const double* __restrict__ x2, double* __restrict__ y) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) y[i] = a * x1[i] + b * x2[i];
}

// ---------------------------------------------------------------------------------------------
// permutation
// ---------------------------------------------------------------------------------------------
struct PermArgs {
    long long n[4];     // extents, in OUTPUT axis order
    long long sin[4];   // input stride of each output axis
    long long total;
    double alpha, beta;
};

// output-fastest axis is also input-fastest (sin[0] == 1): rows of n[0] contiguous elements are
// copied as they are; one warp per row (index arithmetic once per row, not per element), lanes
// stride along the row with 4 independent loads in flight each
__global__ void permute_same_fast(const double* __restrict__ in, double* __restrict__ out, PermArgs a) {
    const long long rows = a.n[1] * a.n[2] * a.n[3];
    const int lane = threadIdx.x & 31;
    long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long rstride = (long long)gridDim.x * (blockDim.x >> 5);
    const long long n0 = a.n[0];
    for (; row < rows; row += rstride) {
        long long r = row;
        const long long i1 = r % a.n[1]; r /= a.n[1];
        const long long i2 = r % a.n[2];
        const long long i3 = r / a.n[2];
        const double* __restrict__ src = in + i1 * a.sin[1] + i2 * a.sin[2] + i3 * a.sin[3];
        double* __restrict__ dst = out + row * n0;
        long long i = lane;
        for (; i + 96 < n0; i += 128) {
            const double v0 = src[i], v1 = src[i + 32], v2 = src[i + 64], v3 = src[i + 96];
            if (a.beta == 0.0) {
                dst[i] = a.alpha * v0; dst[i + 32] = a.alpha * v1; dst[i + 64] = a.alpha * v2; dst[i + 96] = a.alpha * v3;
            } else {
                dst[i] = a.alpha * v0 + a.beta * dst[i];
                dst[i + 32] = a.alpha * v1 + a.beta * dst[i + 32];
                dst[i + 64] = a.alpha * v2 + a.beta * dst[i + 64];
                dst[i + 96] = a.alpha * v3 + a.beta * dst[i + 96];
            }
        }
        for (; i < n0; i += 32) {
            const double v = a.alpha * src[i];
            dst[i] = (a.beta == 0.0) ? v : v + a.beta * dst[i];
        }
    }
}

// general case: 32x32 shared-memory tile transpose between the input-fastest axis (output axis
// `qr`) and the output-fastest axis (axis 0); the other two output axes are looped per block.
struct PermT {
    long long nc, nr, nu, nw;          // extents: c = out axis 0, r = input-fastest axis, u, w others
    long long sin_c, sin_u, sin_w;     // input strides (input stride of r is 1)
    long long sout_r, sout_u, sout_w;  // output strides (output stride of c is 1)
    long long tiles_c, tiles_r;
    double alpha, beta;
};

__global__ void permute_tiled(const double* __restrict__ in, double* __restrict__ out, PermT a) {
    __shared__ double tile[32][33];
    long long b = blockIdx.x;
    const long long tc = b % a.tiles_c; b /= a.tiles_c;
    const long long tr = b % a.tiles_r; b /= a.tiles_r;
    const long long u = b % a.nu;
    const long long w = b / a.nu;
    const long long in_base = u * a.sin_u + w * a.sin_w;
    const long long out_base = u * a.sout_u + w * a.sout_w;
    const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
    // load: x runs along r (input-contiguous), y along c
#pragma unroll
    for (int k = 0; k < 32; k += 8) {
        const long long r = tr * 32 + tx, c = tc * 32 + ty + k;
        if (r < a.nr && c < a.nc) tile[ty + k][tx] = in[in_base + r + c * a.sin_c];
    }
    __syncthreads();
    // store: x runs along c (output-contiguous), y along r
#pragma unroll
    for (int k = 0; k < 32; k += 8) {
        const long long c = tc * 32 + tx, r = tr * 32 + ty + k;
        if (r < a.nr && c < a.nc) {
            const long long o = out_base + c + r * a.sout_r;
            const double v = a.alpha * tile[tx][ty + k];
            out[o] = (a.beta == 0.0) ? v : v + a.beta * out[o];
        }
    }
}


__global__ void splitk_reduce_kernel(const double* __restrict__ W, int nsplit, long long M, long long N,
                                     long long batch, double alpha, double beta, double* C,
                                     long long ldc, long long strideC, const double* Cin) {
    const long long mn = M * N;
    const long long total = mn * batch;
    long long L = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; L < total; L += stride) {
        const long long b = L / mn;
        const long long e = L - b * mn;
        const long long n = e / M, m = e - n * M;
        const double* w = W + b * (long long)nsplit * mn + e;
        double s = 0.0;
        for (int z = 0; z < nsplit; ++z) s += w[(long long)z * mn];
        const long long at = b * strideC + n * ldc + m;
        C[at] = (beta == 0.0) ? alpha * s : alpha * s + beta * Cin[at];
    }
}

// ---------------------------------------------------------------------------------------------
// amplitude helpers
// ---------------------------------------------------------------------------------------------
__global__ void tau_kernel(const double* __restrict__ T, const double* __restrict__ t1, double c,
                           double* __restrict__ out, int o, int v) {
    const long long total = (long long)o * o * v * v;
    long long L = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; L < total; L += stride) {
        long long r = L;
        const int i = (int)(r % o); r /= o;
        const int j = (int)(r % o); r /= o;
        const int a = (int)(r % v);
        const int b = (int)(r / v);
        const double base = T ? T[L] : 0.0;
        out[L] = base + c * t1[i + (long long)o * a] * t1[j + (long long)o * b];
    }
}

// The "one o x o block per (a,b)" kernels below work either with one block per (a,b) pair (two block-wide
// barriers per pair) or, when eight o x (o+1) tiles fit in shared memory (small nocc), with one WARP per pair:
// no block-wide barrier, eight independent load streams per block.
struct PairLoop {
    int nw, wid, tid, nt;
    __device__ PairLoop(bool warp_mode)
        : nw(warp_mode ? (int)(blockDim.x >> 5) : 1), wid(warp_mode ? (int)(threadIdx.x >> 5) : 0),
          tid(warp_mode ? (int)(threadIdx.x & 31) : (int)threadIdx.x), nt(warp_mode ? 32 : (int)blockDim.x) {}
};
template <bool WARP>
__device__ __forceinline__ void pair_sync() {
    if (WARP) __syncwarp(); else __syncthreads();
}

// All amplitude-derived operands of a sweep in ONE pass over T2 (one block per (a,b): the o x o block holds
// both T[i,j] and T[j,i]):  Tt = 2T - T(ji),  tau = T + t(x)t,  tauh = T + 1/2 t(x)t,  Tp2 = T + 2 t(x)t.
// Reads 8 B, writes 8 B per output element; the five separate kernels it replaces moved 2.5x as much.
template <bool WARP>
__global__ void amp_combos_kernel(const double* __restrict__ T, const double* __restrict__ t1,
                                  double* __restrict__ Tt, double* __restrict__ tau, double* __restrict__ tauh,
                                  double* __restrict__ Tp2, int o, int v, const AmpExtras ex) {
    extern __shared__ double sh_all[];  // per pair o x (o+1): sh[j*(o+1)+i] = T[i,j,a,b]
    const PairLoop L(WARP);
    double* sh = sh_all + (size_t)L.wid * o * (o + 1);
    const long long oo = (long long)o * o, ov = (long long)o * v;
    const int half = o >> 1;        // o is even: 16-byte accesses along i
    for (long long ab = (long long)blockIdx.x * L.nw + L.wid; ab < (long long)v * v; ab += (long long)gridDim.x * L.nw) {
        const int a = (int)(ab % v), b = (int)(ab / v);
        const long long base = ab * oo;
        for (int e = L.tid; e < half * o; e += L.nt) {
            const int i = 2 * (e % half), j = e / half;
            const double2 x = *reinterpret_cast<const double2*>(T + base + i + (long long)o * j);
            sh[j * (o + 1) + i] = x.x;
            sh[j * (o + 1) + i + 1] = x.y;
        }
        pair_sync<WARP>();
        const int bl = b - ex.b0;
        const bool in_slab = bl >= 0 && bl < ex.vs;
        for (int e = L.tid; e < half * o; e += L.nt) {
            // p runs along the fastest index of every output; (p,q) is element (i=p, j=q) for the outputs in
            // the layout of T and element (i=q, j=p) for the transposed ones
            const int p = 2 * (e % half), q = e / half;
            const long long at = base + p + (long long)o * q;
            const double x0 = sh[q * (o + 1) + p], x1 = sh[q * (o + 1) + p + 1];        // T[p,q], T[p+1,q]
            const double xt0 = sh[p * (o + 1) + q], xt1 = sh[(p + 1) * (o + 1) + q];    // T[q,p], T[q,p+1]
            *reinterpret_cast<double2*>(Tt + at) = make_double2(2.0 * x0 - xt0, 2.0 * x1 - xt1);
            double tt0 = 0.0, tt1 = 0.0, ts0 = 0.0, ts1 = 0.0;
            if (t1) {
                const double2 ta = *reinterpret_cast<const double2*>(t1 + p + (long long)o * a);
                const double tb = t1[q + (long long)o * b];
                tt0 = ta.x * tb; tt1 = ta.y * tb;                                      // t[p,a] t[q,b]
                *reinterpret_cast<double2*>(tau + at) = make_double2(x0 + tt0, x1 + tt1);
                *reinterpret_cast<double2*>(tauh + at) = make_double2(x0 + 0.5 * tt0, x1 + 0.5 * tt1);
                *reinterpret_cast<double2*>(Tp2 + at) = make_double2(x0 + 2.0 * tt0, x1 + 2.0 * tt1);
                if (ex.X_nfjb && in_slab) {
                    const double2 tb2 = *reinterpret_cast<const double2*>(t1 + p + (long long)o * b);
                    const double tqa = t1[q + (long long)o * a];
                    ts0 = tqa * tb2.x; ts1 = tqa * tb2.y;                              // t[q,a] t[p,b]
                }
            }
            // [., b | ., a] position of an (o,v,o,v) operand: run index + o*b + o*v*(other + o*a)
            const long long ra = p + (long long)o * b + ov * (q + (long long)o * a);
            if (ex.T_meia) *reinterpret_cast<double2*>(ex.T_meia + ra) = make_double2(xt0, xt1);
            if (ex.Tt_meia) *reinterpret_cast<double2*>(ex.Tt_meia + ra) = make_double2(2.0 * xt0 - x0, 2.0 * xt1 - x1);
            if (ex.T_meja) *reinterpret_cast<double2*>(ex.T_meja + ra) = make_double2(x0, x1);
            if (in_slab) {
                const long long rs = p + (long long)o * a + ov * (q + (long long)o * bl);   // [., a | ., b in S]
                if (ex.T_nfjb) *reinterpret_cast<double2*>(ex.T_nfjb + rs) = make_double2(x0, x1);
                if (ex.X_nfjb) *reinterpret_cast<double2*>(ex.X_nfjb + rs) = make_double2(xt0 + 2.0 * ts0, xt1 + 2.0 * ts1);
                if (ex.Y_mnfa)
                    *reinterpret_cast<double2*>(ex.Y_mnfa + p + (long long)o * q + oo * (bl + (long long)ex.vs * a)) =
                        make_double2(x0 + 0.5 * tt0, x1 + 0.5 * tt1);
            }
        }
        pair_sync<WARP>();
    }
}

__global__ void divide_Dijab_kernel(const double* __restrict__ R, double* __restrict__ Tn,
                                    const double* __restrict__ eo, const double* __restrict__ ev, int o,
                                    int v) {
    const long long total = (long long)o * o * v * v;
    long long L = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; L < total; L += stride) {
        long long r = L;
        const int i = (int)(r % o); r /= o;
        const int j = (int)(r % o); r /= o;
        const int a = (int)(r % v);
        const int b = (int)(r / v);
        Tn[L] = R[L] / (eo[i] + eo[j] - ev[a] - ev[b]);
    }
}

// Tnew_S = (V_S + L1_S + L2_S + H_S + P(H)_S) / D for the slab b in [b0, b0+vs) of the last index.
// *_S arrays are (o,o,v,vs) slabs; Hfull is the complete (o,o,v,v) half residual (all-gathered).
// One block per (a,b) pair; the o x o block H[.,.,b,a] is transposed through shared memory so that
// every global access is coalesced.
template <bool WARP>
__global__ void residual_finish_kernel(const double* __restrict__ V, const double* __restrict__ L1,
                                       const double* __restrict__ L2, const double* __restrict__ H,
                                       const double* __restrict__ Hfull, double* __restrict__ Tn,
                                       const double* __restrict__ eo, const double* __restrict__ ev, int o,
                                       int v, int b0, int vs) {
    extern __shared__ double sh_all[];  // per pair o x (o+1): sh[y*(o+1)+x] = H[x,y,b,a]
    const PairLoop L(WARP);
    double* sh = sh_all + (size_t)L.wid * o * (o + 1);
    const long long oo = (long long)o * o;
    const int half = o >> 1;        // o is even: 16-byte accesses along i
    for (long long ab = (long long)blockIdx.x * L.nw + L.wid; ab < (long long)v * vs; ab += (long long)gridDim.x * L.nw) {
        const int a = (int)(ab % v), bl = (int)(ab / v);
        const int b = b0 + bl;
        const long long base = ab * oo;                        // slab-local (a, bl)
        const long long baseT = ((long long)b + (long long)v * a) * oo;  // full (b, a)
        for (int e = L.tid; e < half * o; e += L.nt) {
            const int x = 2 * (e % half), y = e / half;
            const double2 h = *reinterpret_cast<const double2*>(Hfull + baseT + x + (long long)o * y);
            sh[y * (o + 1) + x] = h.x;
            sh[y * (o + 1) + x + 1] = h.y;
        }
        pair_sync<WARP>();
        const double dab = -ev[a] - ev[b];
        for (int e = L.tid; e < half * o; e += L.nt) {
            const int i = 2 * (e % half), j = e / half;
            const long long at = base + i + (long long)o * j;
            const double2 vv = *reinterpret_cast<const double2*>(V + at);
            const double2 hh = *reinterpret_cast<const double2*>(H + at);
            double r0 = vv.x + hh.x + sh[i * (o + 1) + j];          // + H[j,i,b,a]
            double r1 = vv.y + hh.y + sh[(i + 1) * (o + 1) + j];
            if (L1) { const double2 l = *reinterpret_cast<const double2*>(L1 + at); r0 += l.x; r1 += l.y; }
            if (L2) { const double2 l = *reinterpret_cast<const double2*>(L2 + at); r0 += l.x; r1 += l.y; }
            *reinterpret_cast<double2*>(Tn + at) = make_double2(r0 / (eo[i] + eo[j] + dab), r1 / (eo[i + 1] + eo[j] + dab));
        }
        pair_sync<WARP>();
    }
}

// H[i,j,a,b] += R1[i,a,j,b] + R2[j,a,i,b] for the slab (b local): the three particle-hole ring products of a
// sweep leave the GEMM in their natural layouts ([ia|jb] and [ja|ib]); one pass adds both to the half residual
// instead of one permute-accumulate pass over H per product.  One block per (a,b): the o x o block of R2 is
// transposed through shared memory, R1 and H move as 16-byte vectors along i (o is even).
template <bool WARP>
__global__ void ring_combine_kernel(const double* __restrict__ R1, const double* __restrict__ R2,
                                    double* __restrict__ H, int o, int v, int vs) {
    extern __shared__ double sh_all[];  // per pair o x (o+1): sh[i*(o+1)+j] = R2[j,a,i,b]
    const PairLoop L(WARP);
    double* sh = sh_all + (size_t)L.wid * o * (o + 1);
    const long long oo = (long long)o * o, ov = (long long)o * v;
    const int half = o >> 1;
    for (long long ab = (long long)blockIdx.x * L.nw + L.wid; ab < (long long)v * vs; ab += (long long)gridDim.x * L.nw) {
        const int a = (int)(ab % v), b = (int)(ab / v);
        const long long rbase = (long long)o * a + ov * o * b;     // [., a, ., b] of an (o,v,o,vs) array
        for (int e = L.tid; e < half * o; e += L.nt) {
            const int x = 2 * (e % half), y = e / half;            // x runs along the contiguous index j of R2
            const double2 r = *reinterpret_cast<const double2*>(R2 + rbase + x + ov * y);
            sh[y * (o + 1) + x] = r.x;
            sh[y * (o + 1) + x + 1] = r.y;
        }
        pair_sync<WARP>();
        const long long hbase = ab * oo;
        for (int e = L.tid; e < half * o; e += L.nt) {
            const int i = 2 * (e % half), j = e / half;
            const double2 r1 = *reinterpret_cast<const double2*>(R1 + rbase + i + ov * j);
            double2 h = *reinterpret_cast<double2*>(H + hbase + i + (long long)o * j);
            h.x += r1.x + sh[i * (o + 1) + j];
            h.y += r1.y + sh[(i + 1) * (o + 1) + j];
            *reinterpret_cast<double2*>(H + hbase + i + (long long)o * j) = h;
        }
        pair_sync<WARP>();
    }
}

__global__ void divide_Dia_kernel(const double* __restrict__ R, double* __restrict__ tn,
                                  const double* __restrict__ eo, const double* __restrict__ ev, int o, int v) {
    const int total = o * v;
    for (int L = blockIdx.x * blockDim.x + threadIdx.x; L < total; L += gridDim.x * blockDim.x) {
        const int i = L % o, a = L / o;
        tn[L] = R[L] / (eo[i] - ev[a]);
    }
}

// ---------------------------------------------------------------------------------------------
// deterministic reductions
// ---------------------------------------------------------------------------------------------
template <int THREADS>
__device__ __forceinline__ double block_reduce(double v) {
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

// E = sum V[ijab] (2 X[ijab] - X[jiab]) = sum Vt[ijab] X[ijab] with the static Vt = 2V - V(ji) (relabel i <-> j in
// the second term) and X = T + t(x)t: a plain dot product of two contiguous arrays -- no transposed partner, no
// shared memory, no barrier.  16-byte accesses along i (o is even), four independent pairs of loads per thread in
// flight; fixed summation order (deterministic).
__global__ void __launch_bounds__(256) cc_energy_kernel(const double* __restrict__ Vt, const double* __restrict__ T,
                                                        const double* __restrict__ t1, int o, int v,
                                                        double* __restrict__ partial) {
    const long long n2 = (long long)o * o * v * v / 2;          // double2 elements
    const int half = o >> 1;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const double2* __restrict__ V2 = reinterpret_cast<const double2*>(Vt);
    const double2* __restrict__ T2 = reinterpret_cast<const double2*>(T);
    double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
    auto term = [&](long long L, const double2 vv, const double2 x) {
        double x0 = x.x, x1 = x.y;
        if (t1) {
            // L -> (i pair, j, a, b)
            long long r = L;
            const int ip = (int)(r % half); r /= half;
            const int j = (int)(r % o); r /= o;
            const int a = (int)(r % v);
            const int b = (int)(r / v);
            const double2 ta = *reinterpret_cast<const double2*>(t1 + 2 * ip + (long long)o * a);
            const double tb = t1[j + (long long)o * b];
            x0 += ta.x * tb; x1 += ta.y * tb;
        }
        return vv.x * x0 + vv.y * x1;
    };
    long long L = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; L + 3 * stride < n2; L += 4 * stride) {
        const double2 v0 = V2[L], v1 = V2[L + stride], v2 = V2[L + 2 * stride], v3 = V2[L + 3 * stride];
        const double2 x0 = T2[L], x1 = T2[L + stride], x2 = T2[L + 2 * stride], x3 = T2[L + 3 * stride];
        acc0 += term(L, v0, x0);
        acc1 += term(L + stride, v1, x1);
        acc2 += term(L + 2 * stride, v2, x2);
        acc3 += term(L + 3 * stride, v3, x3);
    }
    for (; L < n2; L += stride) acc0 += term(L, V2[L], T2[L]);
    const double r = block_reduce<256>((acc0 + acc1) + (acc2 + acc3));
    if (threadIdx.x == 0) partial[blockIdx.x] = r;
}

__global__ void mp2_energy_kernel(const double* __restrict__ V, const double* __restrict__ eo,
                                  const double* __restrict__ ev, int o, int v, int b0, int vs,
                                  double* __restrict__ partial) {
    const long long oo = (long long)o * o;
    double acc = 0.0;
    for (long long ab = blockIdx.x; ab < (long long)v * vs; ab += gridDim.x) {
        const int a = (int)(ab % v), b = b0 + (int)(ab / v);
        const long long base = ab * oo;
        const double dab = -ev[a] - ev[b];
        for (int e = threadIdx.x; e < oo; e += blockDim.x) {
            const int i = e % o, j = e / o;
            const double x = V[base + e];
            // <ij|ba> = <ji|ab>: the partner lives in the same (a,b) block
            acc += x * (2.0 * x - V[base + j + (long long)o * i]) / (eo[i] + eo[j] + dab);
        }
    }
    const double r = block_reduce<256>(acc);
    if (threadIdx.x == 0) partial[blockIdx.x] = r;
}

__global__ void final_reduce_kernel(const double* __restrict__ partial, int n, double* __restrict__ out) {
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) acc += partial[i];
    const double r = block_reduce<256>(acc);
    if (threadIdx.x == 0) out[0] = r;
}

__global__ void dot_axpby_kernel(const double* __restrict__ x, const double* __restrict__ y, long long n,
                                 double alpha, double beta, double* __restrict__ out) {
    double acc = 0.0;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) acc += x[i] * y[i];
    const double r = block_reduce<256>(acc);
    if (threadIdx.x == 0) out[0] = alpha * r + (beta == 0.0 ? 0.0 : beta * out[0]);
}

__global__ void sqdiff_kernel(const double* __restrict__ x, const double* __restrict__ y, long long n,
                              double* __restrict__ partial) {
    double acc = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const double d = x[i] - y[i];
        acc += d * d;
    }
    const double r = block_reduce<256>(acc);
    if (threadIdx.x == 0) partial[blockIdx.x] = r;
}

// P = e(e+1)/2 + f with 0 <= f <= e
__device__ __forceinline__ void pair_decode(long long P, int* e, int* f) {
    long long ee = (long long)((sqrt(8.0 * (double)P + 1.0) - 1.0) * 0.5);
    while (ee * (ee + 1) / 2 > P) --ee;
    while ((ee + 1) * (ee + 2) / 2 <= P) ++ee;
    *e = (int)ee;
    *f = (int)(P - ee * (ee + 1) / 2);
}

// column (zl, t) of the output-pair space: partner w of z = b0 + zl at slot t (or -1 for the pad slot)
__device__ __forceinline__ int sa_partner(int z, int t, int v) {
    const int n_low = z / 2 + 1;                       // partners w <= z: z%2, z%2+2, .., z
    const int w = t < n_low ? (z & 1) + 2 * t : z + 1 + 2 * (t - n_low);
    return w < v ? w : -1;
}

__global__ void pack_vvvv_sa_kernel(const double* __restrict__ W4, int v, int b0, int vs, long long np,
                                    long long ldk, double* __restrict__ Wp, double* __restrict__ Wm) {
    const int hv = v / 2 + 1;
    const long long total = np * vs * hv;
    for (long long L = (long long)blockIdx.x * blockDim.x + threadIdx.x; L < total;
         L += (long long)gridDim.x * blockDim.x) {
        const long long P = L % np, col = L / np;
        const int zl = (int)(col / hv), t = (int)(col % hv);
        const int z = b0 + zl;
        const int w = sa_partner(z, t, v);
        if (w < 0) continue;                           // pad slot stays zero
        int e, f;
        pair_decode(P, &e, &f);
        const long long wz = (long long)v * v * (w + (long long)v * zl);
        const double x = W4[e + (long long)v * f + wz], y = W4[f + (long long)v * e + wz];   // <ef|wz>, <fe|wz>
        Wp[P + ldk * col] = (e == f) ? 2.0 * x : x + y;
        Wm[P + ldk * col] = (w > z) ? x - y : y - x;   // <ef|hi lo> - <fe|hi lo>, (hi,lo) = the ordered pair
    }
}

__global__ void pack_tau_sa_kernel(const double* __restrict__ tau, int oo, int v, long long np,
                                   double* __restrict__ Tp, double* __restrict__ Tm) {
    for (long long P = blockIdx.x; P < np; P += gridDim.x) {
        int e, f;
        pair_decode(P, &e, &f);
        const double* __restrict__ x = tau + (long long)oo * (e + (long long)v * f);
        const double* __restrict__ y = tau + (long long)oo * (f + (long long)v * e);
        for (int ij = threadIdx.x; ij < oo; ij += blockDim.x) {
            const double xv = x[ij], yv = y[ij];
            Tp[ij + (long long)oo * P] = (e == f) ? xv : xv + yv;
            Tm[ij + (long long)oo * P] = xv - yv;
        }
    }
}

__global__ void unpack_ladder_sa_kernel(const double* __restrict__ Lp, const double* __restrict__ Lm, int oo,
                                        int v, int b0, int vs, double* __restrict__ out) {
    const int hv = v / 2 + 1;
    for (long long ab = blockIdx.x; ab < (long long)v * vs; ab += gridDim.x) {
        const int a = (int)(ab % v), b = b0 + (int)(ab / v);
        const int hi = a > b ? a : b, lo = a > b ? b : a;
        const int z = ((hi - lo) & 1) ? lo : hi, w = ((hi - lo) & 1) ? hi : lo;
        const int n_low = z / 2 + 1;
        const int t = w <= z ? (w - (z & 1)) / 2 : n_low + (w - z - 1) / 2;
        const long long Q = (long long)z * hv + t;
        const double s = a > b ? 0.5 : (a < b ? -0.5 : 0.0);
        const double* __restrict__ lp = Lp + (long long)oo * Q;
        const double* __restrict__ lm = Lm + (long long)oo * Q;
        for (int ij = threadIdx.x; ij < oo; ij += blockDim.x) out[ij + (long long)oo * ab] = 0.5 * lp[ij] + s * lm[ij];
    }
}

__global__ void to_float32_kernel(const double* __restrict__ x, const double* __restrict__ y, long long n,
                                  float* __restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x)
        out[i] = (float)(y ? x[i] - y[i] : x[i]);
}

__global__ void dot_float32_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n,
                                   double* __restrict__ partial) {
    double acc = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x)
        acc += (double)a[i] * (double)b[i];
    const double r = block_reduce<256>(acc);
    if (threadIdx.x == 0) partial[blockIdx.x] = r;
}

struct DiisVecs {
    const float* v[8];
    float c[8];
    int n;
};

__global__ void diis_combine_kernel(DiisVecs d, long long n, double* __restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        double acc = 0.0;
        for (int q = 0; q < d.n; ++q) acc += (double)__fmul_rn(d.c[q], d.v[q][i]);   // Float32 product, no FMA
        out[i] = acc;
    }
}

// Element (mu, a1, a2, a3) of the block; one warp per row (a1, a2, a3) so that the pair index of the
// (lam, sig) / (nu, .) part and all divisions are done once per row, not per element.
__global__ void synth_eri_kernel(double* __restrict__ g, long long n, long long np, long long lam_lo,
                                 long long lam_count, long long sig_lo, long long sig_count,
                                 unsigned long long seed, double scale, int phys) {
    const long long rows = np * lam_count * sig_count;
    const int lane = threadIdx.x & 31;
    long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long rstride = (long long)gridDim.x * (blockDim.x >> 5);
    for (; row < rows; row += rstride) {
        long long r = row;
        unsigned long long nu = (unsigned long long)(r % np); r /= np;
        unsigned long long lam = (unsigned long long)(r % lam_count) + lam_lo;
        if (phys) { const unsigned long long tmp = nu; nu = lam; lam = tmp; }  // g'[mu,lam,nu,sig]
        const unsigned long long sig = (unsigned long long)(r / lam_count) + sig_lo;
        double* __restrict__ dst = g + row * np;
        const bool live = nu < (unsigned long long)n && lam < (unsigned long long)n && sig < (unsigned long long)n;
        // n < 65536 (checked by the launcher): pair indices in 32 bits, one wide multiply per element
        // (synth_element.h, the same function a host build compares bit for bit with the numpy generator)
        const uint32_t Q = synth_pair32((uint32_t)lam, (uint32_t)sig);
        const uint32_t nu32 = (uint32_t)nu;
        for (uint32_t mu = lane; mu < (uint32_t)np; mu += 32) {
            double val = 0.0;
            if (live && mu < (uint32_t)n) val = synth_value(synth_pair32(mu, nu32), Q, seed, scale);
            dst[mu] = val;
        }
    }
}

__global__ void block_copy_kernel(const double* __restrict__ src, double* __restrict__ dst, long long e0,
                                  long long e1, long long e2, long long e3, long long s1, long long s2,
                                  long long s3, long long d1, long long d2, long long d3) {
    const long long total = e0 * e1 * e2 * e3;
    long long L = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; L < total; L += stride) {
        long long r = L;
        const long long i0 = r % e0; r /= e0;
        const long long i1 = r % e1; r /= e1;
        const long long i2 = r % e2;
        const long long i3 = r / e2;
        dst[i0 + i1 * d1 + i2 * d2 + i3 * d3] = src[i0 + i1 * s1 + i2 * s2 + i3 * s3];
    }
}

}  // namespace

int ew_grid(jues_ctx* ctx, size_t n, int threads) {
    size_t blocks = (n + threads - 1) / threads;
    size_t cap = (size_t)ctx->sm_count * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

#define AUX_LAUNCHED(ctx)           \
    JUES_CUDA(cudaGetLastError());  \
    (ctx)->stats.aux_launches++

// trace level 2: one event pair around the launch, named "aux <kernel> <algorithmic MB>" (in-situ durations of
// the HBM-bound kernels for the roofline; tools/sweep_gemm_list.py)
struct AuxTimer {
    Timer* t = nullptr;
    AuxTimer(jues_ctx* ctx, const char* kernel, double bytes) {
        if (ctx->trace >= 2) {
            char nm[64];
            snprintf(nm, sizeof nm, "aux %s %.1fMB", kernel, bytes * 1e-6);
            t = new Timer(ctx, nm);
        }
    }
    ~AuxTimer() { delete t; }
};

// launch geometry of the "one o x o block per (a,b)" kernels: warp-per-pair when eight tiles fit in 48 KB
struct PairLaunch { bool warp; unsigned blocks; size_t smem; };
static PairLaunch pair_launch(jues_ctx* ctx, int64_t o, long long pairs, const char* who) {
    const size_t tile = (size_t)o * (o + 1) * sizeof(double);
    JUES_REQUIRE((o & 1) == 0, "padded nocc must be even");
    if (tile > 200 * 1024) throw Error(JUES_B200_EINVAL, std::string("invalid argument: ") + who + ": nocc too large for the shared-memory block (padded nocc <= 158)");
    PairLaunch p;
    p.warp = 8 * tile <= 48 * 1024;
    const long long want = p.warp ? (pairs + 7) / 8 : pairs;
    const long long cap = (long long)ctx->sm_count * 16;
    p.blocks = (unsigned)std::max<long long>(1, std::min(want, cap));
    p.smem = p.warp ? 8 * tile : tile;
    return p;
}
template <class K>
static void raise_smem(jues_ctx* ctx, K kernel) {
    if (ctx->smem_attr_done.insert((const void*)kernel).second)
        JUES_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
}

void fill_pattern(jues_ctx* ctx, double* p, size_t n, unsigned long long seed) {
    fill_pattern_kernel<<<ew_grid(ctx, n, 256), 256, 0, ctx->stream>>>(p, n, seed);
    AUX_LAUNCHED(ctx);
}

void fill(jues_ctx* ctx, double* p, size_t n, double v) {
    if (!n) return;
    fill_kernel<<<ew_grid(ctx, n, 256), 256, 0, ctx->stream>>>(p, n, v);
    AUX_LAUNCHED(ctx);
}

void axpby(jues_ctx* ctx, size_t n, double a, const double* x, double b, double* y) {
    if (!n) return;
    axpby_kernel<<<ew_grid(ctx, n, 256), 256, 0, ctx->stream>>>(n, a, x, b, y);
    AUX_LAUNCHED(ctx);
}

void lincomb2(jues_ctx* ctx, size_t n, double a, const double* x1, double b, const double* x2, double* y) {
    if (!n) return;
    lincomb2_kernel<<<ew_grid(ctx, n, 256), 256, 0, ctx->stream>>>(n, a, x1, b, x2, y);
    AUX_LAUNCHED(ctx);
}

void permute_axpby(jues_ctx* ctx, double alpha, const Ten& in, const char* ii, double beta,
                   const Ten& out, const char* io) {
    int64_t st[4] = {1, 1, 1, 1};
    int64_t s = 1;
    for (int q = 0; q < in.rank; ++q) { st[q] = s; s *= in.d[q]; }
    permute_axpby_strided(ctx, alpha, in, st, ii, beta, out, io);
}

void permute_axpby_strided(jues_ctx* ctx, double alpha, const Ten& in, const int64_t in_strides[4],
                           const char* ii, double beta, const Ten& out, const char* io) {
    const int rank = (int)strlen(ii);
    JUES_REQUIRE(rank == (int)strlen(io) && rank == in.rank && rank == out.rank && rank >= 1 && rank <= 4,
                 "permute: rank mismatch");
    long long sin_in[4];
    for (int q = 0; q < rank; ++q) sin_in[q] = in_strides[q];
    long long n[4] = {1, 1, 1, 1}, sin[4] = {0, 0, 0, 0}, sout[4] = {0, 0, 0, 0};
    long long so = 1;
    for (int q = 0; q < rank; ++q) {
        const char* f = strchr(ii, io[q]);
        JUES_REQUIRE(f != nullptr, "permute: index letters differ");
        const int src = (int)(f - ii);
        JUES_REQUIRE(in.d[src] == out.d[q], "permute: extent mismatch");
        n[q] = out.d[q];
        sin[q] = sin_in[src];
        sout[q] = so;
        so *= out.d[q];
    }
    const long long total = so;
    if (total == 0) return;
    JUES_REQUIRE(in.p != out.p, "permute: in-place not supported");
    Timer* tk = nullptr;
    if (ctx->trace >= 2) {
        char nm[48];
        snprintf(nm, sizeof nm, "perm %s>%s %lldx%lldx%lldx%lld%s", ii, io, (long long)in.d[0], (long long)in.d[1],
                 (long long)in.d[2], (long long)in.d[3], beta != 0.0 ? "+" : "");
        tk = new Timer(ctx, nm);
    }
    struct TkGuard { Timer* t; ~TkGuard() { delete t; } } tkg{tk};
    // merge adjacent output axes that are also adjacent (same order) in the input
    int r = rank;
    for (int q = 0; q + 1 < r;) {
        if (sin[q + 1] == sin[q] * n[q]) {
            n[q] *= n[q + 1];
            for (int k = q + 1; k + 1 < r; ++k) { n[k] = n[k + 1]; sin[k] = sin[k + 1]; }
            --r;
            n[r] = 1; sin[r] = 0;
        } else {
            ++q;
        }
    }
    so = 1;
    for (int q = 0; q < 4; ++q) { sout[q] = so; so *= n[q]; }
    if (sin[0] == 1) {
        PermArgs a;
        for (int q = 0; q < 4; ++q) { a.n[q] = n[q]; a.sin[q] = sin[q]; }
        a.total = total; a.alpha = alpha; a.beta = beta;
        {
            const long long rows = a.n[1] * a.n[2] * a.n[3];
            long long blocks = (rows + 7) / 8;   // 8 warps (rows) per block
            const long long capb = (long long)ctx->sm_count * 32;
            if (blocks > capb) blocks = capb;
            if (blocks < 1) blocks = 1;
            permute_same_fast<<<(unsigned)blocks, 256, 0, ctx->stream>>>(in.p, out.p, a);
        }
        AUX_LAUNCHED(ctx);
        return;
    }
    int qr = -1;
    for (int q = 1; q < r; ++q) if (sin[q] == 1) qr = q;
    JUES_REQUIRE(qr > 0, "permute: internal (no unit-stride axis)");
    int others[2], no = 0;
    for (int q = 1; q < 4; ++q) if (q != qr) others[no++] = q;
    PermT a;
    a.nc = n[0]; a.nr = n[qr]; a.nu = n[others[0]]; a.nw = n[others[1]];
    a.sin_c = sin[0]; a.sin_u = sin[others[0]]; a.sin_w = sin[others[1]];
    a.sout_r = sout[qr]; a.sout_u = sout[others[0]]; a.sout_w = sout[others[1]];
    a.tiles_c = (a.nc + 31) / 32; a.tiles_r = (a.nr + 31) / 32;
    a.alpha = alpha; a.beta = beta;
    const long long blocks = a.tiles_c * a.tiles_r * a.nu * a.nw;
    JUES_REQUIRE(blocks < (1ll << 31), "permute: grid too large");
    permute_tiled<<<(unsigned)blocks, dim3(32, 8), 0, ctx->stream>>>(in.p, out.p, a);
    AUX_LAUNCHED(ctx);
}

void splitk_reduce(jues_ctx* ctx, const double* W, int nsplit, int64_t M, int64_t N, int64_t batch,
                   double alpha, double beta, double* C, int64_t ldc, int64_t strideC, const double* Cin) {
    const size_t total = (size_t)M * N * batch;
    AuxTimer tm(ctx, "splitk_reduce", 8.0 * (double)total * (nsplit + 1.0 + (beta != 0.0 ? 1.0 : 0.0)));
    splitk_reduce_kernel<<<ew_grid(ctx, total, 256), 256, 0, ctx->stream>>>(W, nsplit, M, N, batch, alpha,
                                                                            beta, C, ldc, strideC, Cin ? Cin : C);
    AUX_LAUNCHED(ctx);
}

void tau_build(jues_ctx* ctx, const double* T, const double* t1, double c, double* out, int64_t o, int64_t v) {
    const size_t total = (size_t)o * o * v * v;
    tau_kernel<<<ew_grid(ctx, total, 256), 256, 0, ctx->stream>>>(T, t1, c, out, (int)o, (int)v);
    AUX_LAUNCHED(ctx);
}

void amp_combos(jues_ctx* ctx, const double* T, const double* t1, double* Tt, double* tau, double* tauh,
                double* Tp2, int64_t o, int64_t v, const AmpExtras* extras) {
    if (v == 0) return;
    const AmpExtras ex = extras ? *extras : AmpExtras();
    double outs = t1 ? 4.0 : 1.0;
    for (const double* q : {ex.T_meia, ex.Tt_meia, ex.T_meja}) outs += q ? 1.0 : 0.0;
    for (const double* q : {ex.T_nfjb, ex.X_nfjb, ex.Y_mnfa}) outs += q ? (double)ex.vs / (double)v : 0.0;
    AuxTimer tm(ctx, "amp_combos", 8.0 * (double)(o * o * v * v) * (1.0 + outs));
    const PairLaunch g = pair_launch(ctx, o, (long long)v * v, "amp_combos");
    if (g.warp) {
        amp_combos_kernel<true><<<g.blocks, 256, g.smem, ctx->stream>>>(T, t1, Tt, tau, tauh, Tp2, (int)o, (int)v, ex);
    } else {
        raise_smem(ctx, amp_combos_kernel<false>);
        amp_combos_kernel<false><<<g.blocks, 256, g.smem, ctx->stream>>>(T, t1, Tt, tau, tauh, Tp2, (int)o, (int)v, ex);
    }
    AUX_LAUNCHED(ctx);
}

void divide_Dijab(jues_ctx* ctx, const double* R, double* Tnew, const double* eo, const double* ev,
                  int64_t o, int64_t v) {
    const size_t total = (size_t)o * o * v * v;
    divide_Dijab_kernel<<<ew_grid(ctx, total, 256), 256, 0, ctx->stream>>>(R, Tnew, eo, ev, (int)o, (int)v);
    AUX_LAUNCHED(ctx);
}

void residual_finish(jues_ctx* ctx, const double* V, const double* L1, const double* L2, const double* H,
                     const double* Hfull, double* Tnew, const double* eo, const double* ev, int64_t o,
                     int64_t v, int64_t b0, int64_t vs) {
    if (v * vs == 0) return;
    const PairLaunch g = pair_launch(ctx, o, (long long)v * vs, "residual_finish");
    AuxTimer tm(ctx, "residual_finish", 8.0 * (double)(o * o * v * vs) * (4.0 + (L1 ? 1.0 : 0.0) + (L2 ? 1.0 : 0.0)));
    if (g.warp) {
        residual_finish_kernel<true><<<g.blocks, 256, g.smem, ctx->stream>>>(V, L1, L2, H, Hfull, Tnew, eo, ev, (int)o,
                                                                             (int)v, (int)b0, (int)vs);
    } else {
        raise_smem(ctx, residual_finish_kernel<false>);
        residual_finish_kernel<false><<<g.blocks, 256, g.smem, ctx->stream>>>(V, L1, L2, H, Hfull, Tnew, eo, ev, (int)o,
                                                                              (int)v, (int)b0, (int)vs);
    }
    AUX_LAUNCHED(ctx);
}

void ring_combine(jues_ctx* ctx, const double* R1, const double* R2, double* H, int64_t o, int64_t v, int64_t vs) {
    if (v * vs == 0) return;
    const PairLaunch g = pair_launch(ctx, o, (long long)v * vs, "ring_combine");
    AuxTimer tm(ctx, "ring_combine", 8.0 * (double)(o * o * v * vs) * 4.0);
    if (g.warp) {
        ring_combine_kernel<true><<<g.blocks, 256, g.smem, ctx->stream>>>(R1, R2, H, (int)o, (int)v, (int)vs);
    } else {
        raise_smem(ctx, ring_combine_kernel<false>);
        ring_combine_kernel<false><<<g.blocks, 256, g.smem, ctx->stream>>>(R1, R2, H, (int)o, (int)v, (int)vs);
    }
    AUX_LAUNCHED(ctx);
}

void divide_Dia(jues_ctx* ctx, const double* R1, double* tnew, const double* eo, const double* ev,
                int64_t o, int64_t v) {
    divide_Dia_kernel<<<ew_grid(ctx, (size_t)o * v, 256), 256, 0, ctx->stream>>>(R1, tnew, eo, ev, (int)o, (int)v);
    AUX_LAUNCHED(ctx);
}

static double finish_reduction(jues_ctx* ctx, int nblocks) {
    final_reduce_kernel<<<1, 256, 0, ctx->stream>>>(ctx->red_dev, nblocks, ctx->red_dev + ctx->red_cap - 1);
    AUX_LAUNCHED(ctx);
    JUES_CUDA(cudaMemcpyAsync(ctx->red_host, ctx->red_dev + ctx->red_cap - 1, sizeof(double),
                              cudaMemcpyDeviceToHost, ctx->stream));
    JUES_CUDA(cudaStreamSynchronize(ctx->stream));
    return ctx->red_host[0];
}

static int cc_energy_launch(jues_ctx* ctx, const double* Vt, const double* T, const double* t1, int64_t o, int64_t v) {
    JUES_REQUIRE((o & 1) == 0, "cc_energy: padded nocc must be even");
    AuxTimer tm(ctx, "cc_energy", 8.0 * (double)(o * o * v * v) * 2.0);
    const size_t n2 = (size_t)(o * o * v * v) / 2;
    int blocks = ew_grid(ctx, (n2 + 3) / 4, 256);                        // four pairs of loads per thread
    blocks = (int)std::min<long long>(blocks, (long long)ctx->red_cap - 2);
    cc_energy_kernel<<<blocks, 256, 0, ctx->stream>>>(Vt, T, t1, (int)o, (int)v, ctx->red_dev);
    AUX_LAUNCHED(ctx);
    return blocks;
}

void cc_energy_async(jues_ctx* ctx, const double* Vt, const double* T, const double* t1, int64_t o, int64_t v,
                     double* dev_out) {
    const int blocks = cc_energy_launch(ctx, Vt, T, t1, o, v);
    final_reduce_kernel<<<1, 256, 0, ctx->stream>>>(ctx->red_dev, blocks, dev_out);
    AUX_LAUNCHED(ctx);
}

double cc_energy(jues_ctx* ctx, const double* Vt, const double* T, const double* t1, int64_t o, int64_t v) {
    const int blocks = cc_energy_launch(ctx, Vt, T, t1, o, v);
    return finish_reduction(ctx, blocks);
}

double mp2_energy(jues_ctx* ctx, const double* V, const double* eo, const double* ev, int64_t o, int64_t v,
                  int64_t b0, int64_t vs) {
    long long blocks = (long long)v * vs;
    const long long cap = std::min<long long>((long long)ctx->sm_count * 16, (long long)ctx->red_cap - 4);
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    mp2_energy_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(V, eo, ev, (int)o, (int)v, (int)b0, (int)vs,
                                                                  ctx->red_dev);
    AUX_LAUNCHED(ctx);
    return finish_reduction(ctx, (int)blocks);
}

void dot_axpby(jues_ctx* ctx, size_t n, double alpha, const double* x, const double* y, double beta,
               double* dev_out) {
    dot_axpby_kernel<<<1, 256, 0, ctx->stream>>>(x, y, (long long)n, alpha, beta, dev_out);
    AUX_LAUNCHED(ctx);
}

void sqdiff_async(jues_ctx* ctx, size_t n, const double* x, const double* y, double* dev_out) {
    int blocks = ew_grid(ctx, n, 256);
    if (blocks > (int)ctx->red_cap - 4) blocks = (int)ctx->red_cap - 4;
    sqdiff_kernel<<<blocks, 256, 0, ctx->stream>>>(x, y, (long long)n, ctx->red_dev);
    AUX_LAUNCHED(ctx);
    final_reduce_kernel<<<1, 256, 0, ctx->stream>>>(ctx->red_dev, blocks, dev_out);
    AUX_LAUNCHED(ctx);
}

void pack_vvvv_sa(jues_ctx* ctx, const double* W4, int64_t v, int64_t b0, int64_t vs, int64_t ldk, double* Wpm) {
    const int64_t np = sa_pairs(v), nq = vs * sa_slots(v);
    pack_vvvv_sa_kernel<<<ew_grid(ctx, (size_t)(np * nq), 256), 256, 0, ctx->stream>>>(
        W4, (int)v, (int)b0, (int)vs, np, ldk, Wpm, Wpm + ldk * nq);
    AUX_LAUNCHED(ctx);
}

void pack_tau_sa(jues_ctx* ctx, const double* tau, int64_t oo, int64_t v, int64_t ldk, double* Tpm) {
    const int64_t np = sa_pairs(v);
    const int threads = oo >= 256 ? 256 : (oo >= 128 ? 128 : 64);
    AuxTimer tm(ctx, "pack_tau_sa", 8.0 * (double)(oo * v * v) * 2.0);
    long long blocks = std::min<long long>(np, (long long)ctx->sm_count * 16);
    pack_tau_sa_kernel<<<(unsigned)blocks, threads, 0, ctx->stream>>>(tau, (int)oo, (int)v, np, Tpm, Tpm + oo * ldk);
    AUX_LAUNCHED(ctx);
}

void unpack_ladder_sa(jues_ctx* ctx, const double* Lpm, int64_t oo, int64_t v, int64_t b0, int64_t vs, double* out) {
    AuxTimer tm(ctx, "unpack_ladder_sa", 8.0 * (double)(oo * v * vs) * 3.0);   // every [L+|L-] element serves (a,b) and (b,a)
    const int threads = oo >= 256 ? 256 : (oo >= 128 ? 128 : 64);
    long long blocks = std::min<long long>(v * vs, (long long)ctx->sm_count * 16);
    if (blocks < 1) return;
    unpack_ladder_sa_kernel<<<(unsigned)blocks, threads, 0, ctx->stream>>>(
        Lpm, Lpm + oo * v * sa_slots(v), (int)oo, (int)v, (int)b0, (int)vs, out);
    AUX_LAUNCHED(ctx);
}

void to_float32(jues_ctx* ctx, size_t n, const double* x, const double* y, float* out32) {
    if (!n) return;
    to_float32_kernel<<<ew_grid(ctx, n, 256), 256, 0, ctx->stream>>>(x, y, (long long)n, out32);
    AUX_LAUNCHED(ctx);
}

void dot_float32_async(jues_ctx* ctx, size_t n, const float* a, const float* b, double* dev_out) {
    int blocks = ew_grid(ctx, n, 256);
    if (blocks > (int)ctx->red_cap - 4) blocks = (int)ctx->red_cap - 4;
    dot_float32_kernel<<<blocks, 256, 0, ctx->stream>>>(a, b, (long long)n, ctx->red_dev);
    AUX_LAUNCHED(ctx);
    final_reduce_kernel<<<1, 256, 0, ctx->stream>>>(ctx->red_dev, blocks, dev_out);
    AUX_LAUNCHED(ctx);
}

void diis_combine(jues_ctx* ctx, size_t n, int nvec, const float* const* vecs, const float* c, double* out) {
    JUES_REQUIRE(nvec >= 1 && nvec <= 8, "diis_combine: 1..8 vectors");
    DiisVecs d;
    d.n = nvec;
    for (int q = 0; q < 8; ++q) { d.v[q] = q < nvec ? vecs[q] : nullptr; d.c[q] = q < nvec ? c[q] : 0.0f; }
    diis_combine_kernel<<<ew_grid(ctx, n, 256), 256, 0, ctx->stream>>>(d, (long long)n, out);
    AUX_LAUNCHED(ctx);
}

void block_copy(jues_ctx* ctx, const double* src, const int64_t sd[4], double* dst, const int64_t dd[4],
                const int64_t ext[4]) {
    const size_t total = (size_t)(ext[0] * ext[1] * ext[2] * ext[3]);
    if (!total) return;
    block_copy_kernel<<<ew_grid(ctx, total, 256), 256, 0, ctx->stream>>>(
        src, dst, ext[0], ext[1], ext[2], ext[3], sd[0], sd[0] * sd[1], sd[0] * sd[1] * sd[2], dd[0],
        dd[0] * dd[1], dd[0] * dd[1] * dd[2]);
    AUX_LAUNCHED(ctx);
}

void synth_eri_fill(jues_ctx* ctx, double* g, int64_t n_logical, int64_t n_padded, int64_t sig_lo,
                    int64_t sig_count, unsigned long long seed, double scale, bool phys) {
    synth_eri_block(ctx, g, n_logical, n_padded, 0, n_padded, sig_lo, sig_count, seed, scale, phys);
}

void synth_eri_block(jues_ctx* ctx, double* g, int64_t n_logical, int64_t n_padded, int64_t lam_lo,
                     int64_t lam_cnt, int64_t sig_lo, int64_t sig_cnt, unsigned long long seed, double scale,
                     bool phys) {
    const size_t total = (size_t)n_padded * n_padded * lam_cnt * sig_cnt;
    if (!total) return;
    JUES_REQUIRE(n_padded < 65536, "synthetic ERIs: more than 65535 basis functions");
    long long blocks = (long long)((total / (size_t)n_padded + 7) / 8);    // 8 rows (warps) per block
    const long long capb = (long long)ctx->sm_count * 32;
    if (blocks > capb) blocks = capb;
    if (blocks < 1) blocks = 1;
    synth_eri_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(g, n_logical, n_padded, lam_lo, lam_cnt,
                                                                sig_lo, sig_cnt, seed, scale, phys ? 1 : 0);
    AUX_LAUNCHED(ctx);
}

}  // namespace jues
