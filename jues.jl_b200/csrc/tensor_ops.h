// Element-wise / permutation / reduction kernels and the dense device tensor view.  Internal.
#pragma once
#include "jues_common.h"

namespace jues {

// Dense column-major device tensor view, rank <= 4 (first index fastest).
struct Ten {
    double* p = nullptr;
    int rank = 0;
    int64_t d[4] = {1, 1, 1, 1};
    Ten() {}
    Ten(double* p_, int64_t d0) : p(p_), rank(1) { d[0] = d0; }
    Ten(double* p_, int64_t d0, int64_t d1) : p(p_), rank(2) { d[0] = d0; d[1] = d1; }
    Ten(double* p_, int64_t d0, int64_t d1, int64_t d2) : p(p_), rank(3) { d[0] = d0; d[1] = d1; d[2] = d2; }
    Ten(double* p_, int64_t d0, int64_t d1, int64_t d2, int64_t d3) : p(p_), rank(4) {
        d[0] = d0; d[1] = d1; d[2] = d2; d[3] = d3;
    }
    int64_t size() const { return d[0] * d[1] * d[2] * d[3]; }
};

// Owning tensor.
struct DTen {
    DBuf buf;
    Ten t;
    DTen() {}
    DTen(jues_ctx* ctx, int64_t d0, int64_t d1 = -1, int64_t d2 = -1, int64_t d3 = -1) { alloc(ctx, d0, d1, d2, d3); }
    void alloc(jues_ctx* ctx, int64_t d0, int64_t d1 = -1, int64_t d2 = -1, int64_t d3 = -1) {
        int64_t n = d0 * (d1 > 0 ? d1 : 1) * (d2 > 0 ? d2 : 1) * (d3 > 0 ? d3 : 1);
        buf.alloc(ctx, (size_t)n);
        t = Ten();
        t.p = buf.p;
        t.rank = d3 > 0 ? 4 : d2 > 0 ? 3 : d1 > 0 ? 2 : 1;
        t.d[0] = d0;
        if (d1 > 0) t.d[1] = d1;
        if (d2 > 0) t.d[2] = d2;
        if (d3 > 0) t.d[3] = d3;
    }
    operator Ten() const { return t; }
    double* p() const { return buf.p; }
    void release() { buf.release(); t = Ten(); }
};

int ew_grid(jues_ctx* ctx, size_t n, int threads);

// out[<io order>] = alpha * in[<ii order>] + beta * out   (io is a permutation of ii's letters)
void permute_axpby(jues_ctx* ctx, double alpha, const Ten& in, const char* ii, double beta,
                   const Ten& out, const char* io);

// the same with an input VIEW: in.d are the extents of the view, in_strides its element strides (a block of
// a larger dense tensor); the output is dense
void permute_axpby_strided(jues_ctx* ctx, double alpha, const Ten& in, const int64_t in_strides[4],
                           const char* ii, double beta, const Ten& out, const char* io);

// y = a*x + b*y  (same layout, n elements)
void axpby(jues_ctx* ctx, size_t n, double a, const double* x, double b, double* y);
// y = a*x1 + b*x2
void lincomb2(jues_ctx* ctx, size_t n, double a, const double* x1, double b, const double* x2, double* y);
void fill(jues_ctx* ctx, double* p, size_t n, double v);

// split-K epilogue: C[m,n] = alpha * sum_z W[z][m,n] + beta * Cin[m,n]   (W slices dense M x N; Cin has
// the layout of C, nullptr = C itself)
void splitk_reduce(jues_ctx* ctx, const double* W, int nsplit, int64_t M, int64_t N, int64_t batch,
                   double alpha, double beta, double* C, int64_t ldc, int64_t strideC,
                   const double* Cin = nullptr);

// out[i,j,a,b] = T[i,j,a,b] + c * t[i,a] * t[j,b]        (o,o,v,v), t is (o,v); T may be null (=0)
void tau_build(jues_ctx* ctx, const double* T, const double* t1, double c, double* out, int64_t o, int64_t v);

// One pass over T2 (o,o,v,v): Tt = 2T - T(ji); with t1 (o,v) also tau = T + t(x)t, tauh = T + 1/2 t(x)t,
// Tp2 = T + 2 t(x)t  (t1 == nullptr: only Tt is written, the other outputs may be null)
// Optional extra outputs: operand layouts of the sweep's ring / Fae products, written while the o x o block
// is in shared memory (each replaces a separate permutation pass: 16 B read + written per element saved).
// S = the slab [b0, b0+vs) of the last index; X = Tp2 with singles, T without; Y = tauh with singles, T without.
struct AmpExtras {
    double* T_meia = nullptr;    // [m,e,i,a] = T[i,m,a,e]              (o,v,o,v)
    double* Tt_meia = nullptr;   // [m,e,i,a] = Tt[i,m,a,e]
    double* T_meja = nullptr;    // [m,e,j,a] = T[m,j,a,e]
    double* T_nfjb = nullptr;    // [n,f,j,b] = T[n,j,f,b], b in S      (o,v,o,vs)
    double* X_nfjb = nullptr;    // [n,f,j,b] = X[j,n,f,b], b in S
    double* Y_mnfa = nullptr;    // [m,n,f,a] = Y[m,n,a,f], f in S      (o,o,vs,v)
    int b0 = 0, vs = 0;
};
void amp_combos(jues_ctx* ctx, const double* T, const double* t1, double* Tt, double* tau, double* tauh,
                double* Tp2, int64_t o, int64_t v, const AmpExtras* extras = nullptr);

// Tnew[i,j,a,b] = R[i,j,a,b] / (eo[i] + eo[j] - ev[a] - ev[b])       (may be in place)
void divide_Dijab(jues_ctx* ctx, const double* R, double* Tnew, const double* eo, const double* ev,
                  int64_t o, int64_t v);
// Tnew_S = (V_S + L1_S + L2_S + H_S + P(Hfull)_S) / D for the last-index slab [b0, b0+vs);
// *_S are (o,o,v,vs) slabs, Hfull the complete (o,o,v,v) half residual; P(H)[i,j,a,b] = H[j,i,b,a].
void residual_finish(jues_ctx* ctx, const double* V, const double* L1, const double* L2, const double* H,
                     const double* Hfull, double* Tnew, const double* eo, const double* ev, int64_t o,
                     int64_t v, int64_t b0, int64_t vs);
// H[i,j,a,b] += R1[i,a,j,b] + R2[j,a,i,b]: H an (o,o,v,vs) slab, R1 and R2 (o,v,o,vs) -- the ring products in
// the layouts the GEMM leaves them in, added to the half residual in one pass
void ring_combine(jues_ctx* ctx, const double* R1, const double* R2, double* H, int64_t o, int64_t v, int64_t vs);
// tnew[i,a] = R1[i,a] / (eo[i] - ev[a])
void divide_Dia(jues_ctx* ctx, const double* R1, double* tnew, const double* eo, const double* ev,
                int64_t o, int64_t v);

// deterministic reductions (fixed-shape two-pass tree); result returned on the host
// E = sum_{ijab} V[ijab] * (2*X[ijab] - X[jiab]) = sum_{ijab} Vt[ijab] * X[ijab],  X = T + t(x)t (t nullable),
// Vt = 2V - V(ji) (the static combination the sweep already holds): a plain dot product
double cc_energy(jues_ctx* ctx, const double* Vt, const double* T, const double* t1, int64_t o, int64_t v);
// same reduction, result left in dev_out[0] (no host synchronisation)
void cc_energy_async(jues_ctx* ctx, const double* Vt, const double* T, const double* t1, int64_t o, int64_t v,
                     double* dev_out);
// E = sum_{ijab} v[ijab] (2 v[ijab] - v[ijba]) / (eo[i]+eo[j]-ev[a]-ev[b])
// v is the last-index slab (o,o,vv,vs) of <ij|ab>, b in [b0, b0+vs); v_ijba is taken as v_jiab
double mp2_energy(jues_ctx* ctx, const double* v, const double* eo, const double* ev, int64_t o, int64_t vv,
                  int64_t b0, int64_t vs);

// dev_out[0] = alpha * sum_k x[k] y[k] + beta * dev_out[0]   (one block, fixed order; for o*v-sized vectors)
void dot_axpby(jues_ctx* ctx, size_t n, double alpha, const double* x, const double* y, double beta,
               double* dev_out);
// dev_out[0] = sum_k (x[k] - y[k])^2   (deterministic two-pass tree, no host synchronisation)
void sqdiff_async(jues_ctx* ctx, size_t n, const double* x, const double* y, double* dev_out);

// ---- symmetric / antisymmetric particle-particle ladder ---------------------------------------------
// sum_ef tau[ij,ef] <ef|ab> = 1/2 sum_{e>=f} tau+[ij,(ef)] W+[(ef),(ab)] + 1/2 sum_{e>f} tau-[ij,(ef)] W-[(ef),(ab)]
// with x+- = x[ef] +- x[fe] (the e == f member of the + part counted once in tau+, twice in W+), because
// <ef|ab> = <fe|ba>: W+ is symmetric and W- antisymmetric in (a,b), so each unordered pair {a,b} is
// computed once -- half the flops of the plain (o^2 x v^2)(v^2 x v^2) product.
//   * summed pairs (K space): P(e,f) = e(e+1)/2 + f, e >= f, np = v(v+1)/2 (the diagonal of the - part is
//     zero), leading dimension ldk = np rounded up to even;
//   * output pairs (N space), laid out so that they shard with the virtual slabs of the CC driver: the
//     pair {x,y} belongs to the column block of z = max(x,y) when x - y is even, of z = min(x,y) when it
//     is odd, at position t among z's partners (w = z%2, z%2+2, .., z, then z+1, z+3, ..), i.e. column
//     Q = z*hv + t with hv = v/2 + 1 slots per z (one is a zero pad when z is odd).  Every z owns ~v/2
//     pairs, so equal slabs of z are equal shares of the work, and W+-[.,(z,w)] needs only <..|w z>
//     with z in the rank's slab of the LAST index -- exactly the <vv|vv> slab the rank already holds.
inline int64_t sa_pairs(int64_t v) { return v * (v + 1) / 2; }
inline int64_t sa_slots(int64_t v) { return v / 2 + 1; }
// Wpm[0] = W+ (ldk x vs*hv), Wpm[1] = W- from the slab W4[e,f,w,z-b0] = <ef|wz>, z in [b0, b0+vs).
// Wpm must be zero-initialised (pad columns).  Once per calculation.
void pack_vvvv_sa(jues_ctx* ctx, const double* W4, int64_t v, int64_t b0, int64_t vs, int64_t ldk, double* Wpm);
// Tpm[0] = tau+ (oo x ldk), Tpm[1] = tau- from tau[i,j,e,f] (oo = o*o rows).  Once per sweep.
void pack_tau_sa(jues_ctx* ctx, const double* tau, int64_t oo, int64_t v, int64_t ldk, double* Tpm);
// out[ij,a,b-b0] = 1/2 (L+[ij,Q(a,b)] + sign(a-b) L-[ij,Q(a,b)]) for b in [b0, b0+vs), all a;
// Lpm = [L+ | L-], each oo x v*hv (ALL column blocks: all-gathered when there are several ranks).
void unpack_ladder_sa(jues_ctx* ctx, const double* Lpm, int64_t oo, int64_t v, int64_t b0, int64_t vs, double* out);

// ---- mRCCD's DIIS keeps its vectors in Float32 (mRCCD.jl:64-65,171,175) --------------------------
// out32[k] = float(x[k] - y[k])   (y nullable: plain conversion)
void to_float32(jues_ctx* ctx, size_t n, const double* x, const double* y, float* out32);
// dev_out[0] = sum_k a[k] b[k] of two Float32 vectors, accumulated in FP64 (deterministic tree)
void dot_float32_async(jues_ctx* ctx, size_t n, const float* a, const float* b, double* dev_out);
// out[k] = sum_{q<nvec} double(c[q] * vecs[q][k])   -- Float32 products summed in FP64 in the order q = 0,1,...
// exactly as `tiJaB_d .+= convert(Float32,ci[num])*diis_vals_t2[num+1]` does (mRCCD.jl:200-202)
void diis_combine(jues_ctx* ctx, size_t n, int nvec, const float* const* vecs, const float* c, double* out);

// counter-based synthetic ERIs (same function as jues.jl_b200.synth.counter_eri_element)
void synth_eri_fill(jues_ctx* ctx, double* g, int64_t n_logical, int64_t n_padded, int64_t sig_lo,
                    int64_t sig_count, unsigned long long seed, double scale, bool phys = false);

// general block g[:, :, l0:l0+lc, s0:s0+sc] of the same generator, dense (np, np, lc, sc)
void synth_eri_block(jues_ctx* ctx, double* g, int64_t n_logical, int64_t n_padded, int64_t lam_lo,
                     int64_t lam_cnt, int64_t sig_lo, int64_t sig_cnt, unsigned long long seed, double scale,
                     bool phys);

}  // namespace jues

namespace jues {
// dst[i0,i1,i2,i3] = src[i0,i1,i2,i3] for i_q < ext[q]; both dense column-major with their own
// extents (sd, dd >= ext): pads / unpads / slices a rank-4 block on the device.
void block_copy(jues_ctx* ctx, const double* src, const int64_t sd[4], double* dst, const int64_t dd[4],
                const int64_t ext[4]);
}  // namespace jues
