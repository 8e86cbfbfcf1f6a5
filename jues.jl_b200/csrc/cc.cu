// RMP2, RCCD and RCCSD on the device.
//
// The amplitude equations are the reference's (RCCD.jl:119-303, RCCSD.jl:150-289) re-factorised
// for the GPU (tests/factorized_model.py is the numpy statement of exactly this algorithm and is
// checked against the literal oracle):
//   * six symmetry-unique MO integral classes instead of 15 transformed-and-permuted copies
//     (make_rccsd_integrals, RCCSD.jl:117-142), produced directly in physicists' order by
//     transforming the AO tensor re-ordered once to [mu,lam,nu,sig];
//   * Wabef (v^4, rebuilt every sweep by RCCSD.jl:220-228) is never formed:
//       tau.Wabef = tau.vvvv  -  (1+P) t.(tau.ovvv)  +  X.tau,   X = 1/2 tau.oovv
//     and X is shared with Wmnij, so the hole-hole ladder uses  oooo + t.ooov + t.oovo + 2X;
//   * every term that comes with its (i<->j, a<->b) image is evaluated once:
//       R2 = oovv + Lpp + Lhh + (1 + P)(H),  finished by one fused kernel that also divides by
//       the orbital-energy denominator (Dijab is never materialised).
// Every contraction below is one launch of the TMA + DMMA GEMM (contract.cu).
#include "cc.h"
#include "dist.h"
#include "pt.h"
#include "dgemm.h"

#include <algorithm>
#include <cmath>

namespace jues {

void setup_problem(jues_ctx* ctx, Problem& P, int64_t nao, const double* Cao, int64_t nocc,
                   const double* Cav, int64_t nvir, const double* eps) {
    JUES_REQUIRE(nao > 0 && nocc > 0 && nvir > 0, "nao, nocc and nvir must be positive");
    JUES_REQUIRE(Cao && Cav && eps, "null orbital data");
    P.nao = nao; P.nocc = nocc; P.nvir = nvir;
    // v is padded to a multiple of 2*nranks so that the virtual index splits into equal even slabs
    P.np = round_up(nao, 2); P.o = round_up(nocc, 2); P.v = round_up(nvir, 2 * (int64_t)ctx->nranks);
    upload_padded_matrix(ctx, P.Co, Cao, nao, nocc, P.np, P.o);
    upload_padded_matrix(ctx, P.Cv, Cav, nao, nvir, P.np, P.v);
    double emin = eps[0], emax = eps[0];
    for (int64_t k = 0; k < nocc + nvir; ++k) { emin = std::min(emin, eps[k]); emax = std::max(emax, eps[k]); }
    std::vector<double> eo(P.o, emin - 1.0e3), ev(P.v, emax + 1.0e3);
    for (int64_t i = 0; i < nocc; ++i) eo[i] = eps[i];
    for (int64_t a = 0; a < nvir; ++a) ev[a] = eps[nocc + a];
    P.eo.alloc(ctx, P.o); P.ev.alloc(ctx, P.v);
    JUES_CUDA(cudaMemcpyAsync(P.eo.p, eo.data(), P.o * 8, cudaMemcpyHostToDevice, ctx->stream));
    JUES_CUDA(cudaMemcpyAsync(P.ev.p, ev.data(), P.v * 8, cudaMemcpyHostToDevice, ctx->stream));
    JUES_CUDA(cudaStreamSynchronize(ctx->stream));  // eo/ev are stack vectors
}

// ---------------------------------------------------------------------------------------------
// RMP2
// ---------------------------------------------------------------------------------------------
double rmp2_dev(jues_ctx* ctx, Problem& P, GaoSource& gao) {
    const int64_t o = P.o, v = P.v;
    int64_t b0, vs;
    slab_of(ctx, v, &b0, &vs);   // this rank's slab of the virtual index b (everything when nranks == 1)
    // E = sum_{ij a b} v_ijab (2 v_ijab - v_ijba) / D with v_ijba = v_jiab: every (a, b in slab) block is
    // self-contained, so the energy needs no exchange but the final scalar.  <ij|a b_S> = (ia|j b_S)
    // (IntegralTransformation.jl:96-98) comes out of the sharded one-pass transform already in the layout
    // the energy kernel reads: each rank touches 1/P of the AO tensor, occupied indices are contracted
    // first (2 N^4 o / P flops), and the exchanged block is only o^2 (N/P) v.
    DBuf ijab;
    {
        Timer t(ctx, "mp2.transform");
        const std::vector<int64_t> counts((size_t)ctx->nranks, vs);
        tei_transform_sharded(ctx, gao, P.Co.p, o, P.Co.p, o, P.Cv.p, v, P.Cv.p, counts, ijab);
    }
    Timer t(ctx, "mp2.energy");
    return all_reduce_scalar(ctx, mp2_energy(ctx, ijab.p, P.eo.p, P.ev.p, o, v, b0, vs));
}

// ---------------------------------------------------------------------------------------------
// get_fock (IntegralTransformation.jl:119-141)
// ---------------------------------------------------------------------------------------------
void fock_dev(jues_ctx* ctx, GaoSource& gao, const double* hao, const double* C, int64_t nmo,
              const double* Co, int64_t nocc, double* f_out) {
    JUES_REQUIRE(hao && C && Co && f_out, "get_fock: null argument");
    JUES_REQUIRE(nmo > 0 && nocc > 0, "get_fock: nmo and nocc must be positive");
    const int64_t n = gao.n, np = gao.np;
    const int64_t mp = round_up(nmo, 2), op = round_up(nocc, 2);
    DBuf Cd, Cod, hd;
    upload_padded_matrix(ctx, Cd, C, n, nmo, np, mp);
    upload_padded_matrix(ctx, Cod, Co, n, nocc, np, op);
    upload_padded_matrix(ctx, hd, hao, n, n, np, np);
    std::vector<double> eye_h((size_t)(op * op), 0.0);
    for (int64_t k = 0; k < op; ++k) eye_h[k + op * k] = 1.0;
    DTen f(ctx, mp, mp), tmp(ctx, np, mp), eye(ctx, op, op);
    JUES_CUDA(cudaMemcpyAsync(eye.p(), eye_h.data(), eye_h.size() * sizeof(double), cudaMemcpyHostToDevice,
                              ctx->stream));
    JUES_CUDA(cudaStreamSynchronize(ctx->stream));   // eye_h is a local vector
    const Ten Ct(Cd.p, np, mp), ht(hd.p, np, np);
    // f[p,q] = C[mu,p] C[nu,q] hao[mu,nu]                                               (:136)
    contract(ctx, 1.0, ht, "mn", Ct, "nq", 0.0, tmp, "mq");
    contract(ctx, 1.0, Ct, "mp", tmp, "mq", 0.0, f, "pq");
    {   // + 2 sum_k (pq|kk): the (C,C,Co,Co) transform traced over its occupied pair          (:137)
        DTen A(ctx, mp, mp, op, op);
        const double* Cm[4] = {Cd.p, Cd.p, Cod.p, Cod.p};
        const int64_t dp[4] = {mp, mp, op, op};
        tei_transform_dev(ctx, gao, Cm, dp, A.p());
        contract(ctx, 2.0, A, "pqkl", eye, "kl", 1.0, f, "pq");
    }
    {   // - sum_k (pk|qk)                                                                     (:138)
        DTen B(ctx, mp, op, mp, op);
        const double* Cm[4] = {Cd.p, Cod.p, Cd.p, Cod.p};
        const int64_t dp[4] = {mp, op, mp, op};
        tei_transform_dev(ctx, gao, Cm, dp, B.p());
        contract(ctx, -1.0, B, "pkql", eye, "kl", 1.0, f, "pq");
    }
    JUES_CUDA(cudaMemcpy2DAsync(f_out, nmo * 8, f.p(), mp * 8, nmo * 8, nmo, cudaMemcpyDeviceToHost,
                                ctx->stream));
    JUES_CUDA(cudaStreamSynchronize(ctx->stream));
}

// ---------------------------------------------------------------------------------------------
// coupled cluster
// ---------------------------------------------------------------------------------------------
namespace {

// view of the last-index slab [b0, b0+n) of a dense tensor
Ten last_slab(const Ten& X, int64_t b0, int64_t n) {
    Ten r = X;
    int64_t lead = 1;
    for (int q = 0; q + 1 < X.rank; ++q) lead *= X.d[q];
    r.p = X.p + b0 * lead;
    r.d[X.rank - 1] = n;
    return r;
}

// One rank's share of the RCCD / RCCSD problem.  The output virtual index b of T2new[i,j,a,b] is
// split into equal slabs; this rank owns b in [b0, b0+vs) (everything when nranks == 1).
// tests/sharded_model.py is the numpy statement of exactly this algorithm (checked against the
// oracle with 2 and 3 gloo ranks on the CPU).
struct CC {
    jues_ctx* ctx;
    Problem& P;
    bool singles;
    int64_t o, v, b0, vs;
    // replicated integral classes (physicists' order, names as in the reference)
    DTen V, J, ooov, oooo;
    // last-index slabs of the large classes:  W4[e,f,a,b] = <ef|ab>,  OA[e,f,m,b] = <ef|mb>
    // (= ovvv[m,b,e,f]),  OB[a,j,e,b] = <aj|eb> (= (ae|jb)),  b in the slab
    DTen W4, OA, OB;
    // <vv|vv> packed into its (ef)-symmetric and -antisymmetric parts [W+ | W-] for this rank's column
    // block of the output-pair space (tensor_ops.h: half the ladder flops); W4 itself is released once
    // they are built
    DTen Wsa;
    int64_t sa_ld = 0, sa_nq = 0;
    bool sa_ladder = false;
    // static combinations
    DTen Vt, oovo, ooov_t;
    // OC[a,m,e,f] = 2 <am|ef> - <ma|ef> for f in the slab (one operand for the two singles terms that need both);
    // ooov_p[i,j,m,b] = <mj|ib>, b in the slab (start value of the o^3 v intermediate contracted with t[m,a])
    DTen OC, ooov_p;
    // Vx[m,e,j,b] = <mj|eb>, b in the slab: start value of the ring intermediate WJ
    DTen Vx;
    // The sweep with layouts chosen so that no GEMM output needs a permutation pass (batched products over the
    // slab index, ring products added to H by one kernel, the Fmi term through its (ij)(ab) image).
    // JUES_B200_PLAIN_SWEEP=1 keeps the one-contraction-per-term form of round 1 (A/B measurements).
    bool relaid = getenv("JUES_B200_PLAIN_SWEEP") == nullptr;
    // DF-RCCD.jl:256 contracts <mn|ef> where RCCD.jl:402 contracts <nm|ef> in the third term of WmBeJ, i.e. its
    // ring intermediate is <mb|ej> + 1/2 <mn|ef> (T[njfb] - T[jnfb]): the density-fitted driver follows its file
    bool df_wmbej = false;
    static bool no_amp_extras() {
        static const bool off = getenv("JUES_B200_NO_AMP_EXTRAS") != nullptr;     // A/B switch for measurements
        return off;
    }
    // off-diagonal Fock blocks of a non-canonical reference (AutoRCCSD.jl:218-231), zero diagonals,
    // zero padding:  foT[m,i] = f[i,m] (o,o),  fov[m,e] (o,v),  fvv[e,a] (v,v).  fock == false: canonical.
    DTen foT, fov, fvv;
    bool fock = false;
    // amplitudes (replicated)
    DTen T1, T2, T1n, T2n;
    PermCache pcache;         // permuted operand copies (static: once; amplitude-derived: once per sweep)

    CC(jues_ctx* c, Problem& p, bool s) : ctx(c), P(p), singles(s), o(p.o), v(p.v) { slab_of(c, v, &b0, &vs); }

    // All integral classes from ONE pass over the AO tensor, shared by the ranks (tei_transform_sharded):
    // the slab M[p,q,r,s] = <pq|rs> over all MO p, q, r and this rank's columns s = [its share of the
    // occupied orbitals | its virtual slab]; the classes are sub-blocks of it (RCCSD.jl:117-132 runs 15
    // separate transforms, mRCCSD.jl:102-110 shows 6 classes suffice).  The o^2 v^2-sized classes every rank
    // needs in full are all-gathered from their slabs.
    void extract(const DBuf& M, int64_t nmo, int64_t nsl, int64_t p0, int64_t q0, int64_t r0, int64_t s0,
                 double* dst, int64_t e0, int64_t e1, int64_t e2, int64_t e3) {
        const int64_t sd[4] = {nmo, nmo, nmo, nsl}, dd[4] = {e0, e1, e2, e3};
        block_copy(ctx, M.p + p0 + nmo * (q0 + nmo * (r0 + nmo * s0)), sd, dst, dd, dd);
    }

    void build_integrals(GaoSource& gao) {
        {
            Timer t(ctx, "cc.transform");
            const int P_ = ctx->nranks;
            const int64_t np = gao.np, nmo = o + v;
            const int64_t oP = round_up(o, P_), os = oP / P_, nsl = os + vs;
            DBuf Call(ctx, (size_t)(np * nmo)), Cs(ctx, (size_t)(np * nsl * P_));
            JUES_CUDA(cudaMemcpyAsync(Call.p, P.Co.p, (size_t)(np * o) * 8, cudaMemcpyDeviceToDevice, ctx->stream));
            JUES_CUDA(cudaMemcpyAsync(Call.p + np * o, P.Cv.p, (size_t)(np * v) * 8, cudaMemcpyDeviceToDevice,
                                      ctx->stream));
            Cs.zero();
            for (int d = 0; d < P_; ++d) {
                const int64_t i0 = std::min<int64_t>(o, d * os), i1 = std::min<int64_t>(o, (d + 1) * os);
                if (i1 > i0)
                    JUES_CUDA(cudaMemcpyAsync(Cs.p + np * nsl * d, P.Co.p + np * i0, (size_t)(np * (i1 - i0)) * 8,
                                              cudaMemcpyDeviceToDevice, ctx->stream));
                JUES_CUDA(cudaMemcpyAsync(Cs.p + np * (nsl * d + os), P.Cv.p + np * vs * d, (size_t)(np * vs) * 8,
                                          cudaMemcpyDeviceToDevice, ctx->stream));
            }
            DBuf M;
            tei_transform_sharded(ctx, gao, Call.p, nmo, Call.p, nmo, Call.p, nmo, Cs.p,
                                  std::vector<int64_t>((size_t)P_, nsl), M);
            TraceTimer tc(ctx, "cc.classes");
            // slabs of the replicated classes, gathered in place
            V.alloc(ctx, o, o, v, v);
            extract(M, nmo, nsl, 0, 0, o, os, V.p() + b0 * o * o * v, o, o, v, vs);          // <ij|ab>
            all_gather_inplace(ctx, V.p(), (size_t)(o * o * v * vs));
            J.alloc(ctx, o, v, o, v);
            extract(M, nmo, nsl, 0, o, 0, os, J.p() + b0 * o * v * o, o, v, o, vs);          // <mb|je>
            all_gather_inplace(ctx, J.p(), (size_t)(o * v * o * vs));
            oooo.alloc(ctx, o, o, o, oP);                                                      // <mn|ij>, j shared out
            extract(M, nmo, nsl, 0, 0, 0, 0, oooo.p() + (int64_t)ctx->rank * os * o * o * o, o, o, o, os);
            all_gather_inplace(ctx, oooo.p(), (size_t)(o * o * o * os));
            oooo.t.d[3] = o;                                                                   // the padding shares are zero
            W4.alloc(ctx, v, v, v, vs);
            extract(M, nmo, nsl, o, o, o, os, W4.p(), v, v, v, vs);                            // <ef|ab>
            if (singles) {
                ooov.alloc(ctx, o, o, o, v);
                extract(M, nmo, nsl, 0, 0, 0, os, ooov.p() + b0 * o * o * o, o, o, o, vs);   // <mn|ie>
                all_gather_inplace(ctx, ooov.p(), (size_t)(o * o * o * vs));
                OA.alloc(ctx, v, v, o, vs);
                extract(M, nmo, nsl, o, o, 0, os, OA.p(), v, v, o, vs);                        // <ef|mb>
                OB.alloc(ctx, v, o, v, vs);
                extract(M, nmo, nsl, o, 0, o, os, OB.p(), v, o, v, vs);                        // <aj|eb>
            }
        }
        build_static();
    }

    // packed ladder operand and the static combinations of the classes (shared by the transform and the
    // density-fitted builds)
    void build_static() {
        Timer t(ctx, "cc.static");
        if (getenv("JUES_B200_PLAIN_LADDER") == nullptr) {
            sa_ld = round_up(sa_pairs(v), 2);
            sa_nq = vs * sa_slots(v);
            Wsa.alloc(ctx, sa_ld, sa_nq, 2);
            Wsa.buf.zero();
            pack_vvvv_sa(ctx, W4.p(), v, b0, vs, sa_ld, Wsa.p());
            W4.release();
            sa_ladder = true;
        }
        const size_t n2 = (size_t)(o * o * v * v);
        Vt.alloc(ctx, o, o, v, v);
        axpby(ctx, n2, 2.0, V.p(), 0.0, Vt.p());
        permute_axpby(ctx, -1.0, V, "ijab", 1.0, Vt, "jiab");          // Vt = 2V - V(ji)
        if (relaid) {
            Vx.alloc(ctx, o, v, o, vs);
            permute_axpby(ctx, 1.0, last_slab(V, b0, vs), "mjeb", 0.0, Vx, "mejb");
        }
        if (singles) {
            oovo.alloc(ctx, o, o, v, o);
            permute_axpby(ctx, 1.0, ooov, "nmje", 0.0, oovo, "mnej");   // <mn|ej> = <nm|je>
            ooov_t.alloc(ctx, o, o, o, v);
            axpby(ctx, (size_t)(o * o * o * v), 2.0, ooov.p(), 0.0, ooov_t.p());
            permute_axpby(ctx, -1.0, ooov, "mnie", 1.0, ooov_t, "nmie");
            if (relaid) {
                OC.alloc(ctx, v, o, v, vs);
                permute_axpby(ctx, -1.0, OA, "eamf", 0.0, OC, "amef");
                axpby(ctx, (size_t)(v * o * v * vs), 2.0, OB.p(), 1.0, OC.p());
                ooov_p.alloc(ctx, o, o, o, vs);
                permute_axpby(ctx, 1.0, last_slab(ooov, b0, vs), "mjib", 0.0, ooov_p, "ijmb");
            }
        }
    }

    // The integral classes of RCCD from three-index tensors (DF-RCCD.jl:77-89 and the products its sweep
    // forms from them): bov[i,a,Q], boo[i,j,Q], bvv[a,b,Q] on the device, Q padded to nx.
    //   <ij|ab> = bov[i,a,Q] bov[j,b,Q]      <me|jb> = boo[m,j,Q] bvv[e,b,Q]
    //   <mn|ij> = boo[m,i,Q] boo[n,j,Q]      <ef|ab> = bvv[e,a,Q] bvv[f,b,Q]   (b in this rank's slab)
    // The o^2 v^2-sized classes are formed on every rank (o^2 v^2 naux flops), <vv|vv> only for the slab.
    void build_integrals_df(const Ten& bov, const Ten& boo, const Ten& bvv) {
        JUES_REQUIRE(!singles, "the density-fitted classes are those of RCCD");
        {
            Timer t(ctx, "cc.transform");
            const int64_t nx = bov.d[2];
            V.alloc(ctx, o, o, v, v);
            contract(ctx, 1.0, bov, "iaQ", bov, "jbQ", 0.0, V, "ijab");
            J.alloc(ctx, o, v, o, v);
            contract(ctx, 1.0, boo, "mjQ", bvv, "ebQ", 0.0, J, "mejb");
            oooo.alloc(ctx, o, o, o, o);
            contract(ctx, 1.0, boo, "miQ", boo, "njQ", 0.0, oooo, "mnij");
            DTen bvvS(ctx, v, vs, nx);
            const int64_t sd[4] = {v, v, nx, 1}, dd[4] = {v, vs, nx, 1}, ext[4] = {v, vs, nx, 1};
            block_copy(ctx, bvv.p + v * b0, sd, bvvS.p(), dd, ext);
            W4.alloc(ctx, v, v, v, vs);
            contract(ctx, 1.0, bvv, "eaQ", bvvS, "fbQ", 0.0, W4, "efab");
        }
        build_static();
    }

    void register_static() {
        ctx->perm_cache = &pcache;
        for (DTen* t : {&V, &J, &oooo, &ooov, &Vt, &oovo, &ooov_t, &OA, &OB, &OC})
            if (t->p()) pcache.add(t->t, false);
    }
    ~CC() {
        bool had_graph = false;
        for (auto& g : graphs)
            if (g.exec) { cudaGraphExecDestroy(g.exec); had_graph = true; }
        if (had_graph) cudaDeviceGraphMemTrim(ctx->device);   // give the graphs' memory back to the device
        ctx->perm_cache = nullptr;
        pcache.clear();
        if (ctx->arena == &arena) ctx->arena = nullptr;        // members that still hold arena blocks just drop them
        if (ctx->arena_main == &arena) ctx->arena_main = nullptr;
        if (ctx->arena_side == &arena2) ctx->arena_side = nullptr;
        if (ev_fork) { cudaEventDestroy(ev_fork); cudaEventDestroy(ev_join); cudaEventDestroy(ev_fork2); cudaEventDestroy(ev_join2); }
    }

    double energy() { return cc_energy(ctx, Vt.p(), T2.p(), singles ? T1.p() : nullptr, o, v); }
    void energy_async(double* dev_out) { cc_energy_async(ctx, Vt.p(), T2.p(), singles ? T1.p() : nullptr, o, v, dev_out); }

    void guess(int guess_mode) {
        T2.alloc(ctx, o, o, v, v); T2n.alloc(ctx, o, o, v, v);
        T1.alloc(ctx, o, v); T1n.alloc(ctx, o, v);
        T1.buf.zero(); T1n.buf.zero();
        if (!singles && guess_mode == 0) {
            // RCCD.jl:45,145-160: T2[i,j,a,b] = ovov[i,a,j,b] / D = (ij|ab) / D
            DTen tmp(ctx, o, o, v, v);
            permute_axpby(ctx, 1.0, J, "iajb", 0.0, tmp, "ijab");
            divide_Dijab(ctx, tmp.p(), T2.p(), P.eo.p, P.ev.p, o, v);
        } else {
            divide_Dijab(ctx, V.p(), T2.p(), P.eo.p, P.ev.p, o, v);      // RCCSD.jl:80
        }
    }

    // one Jacobi sweep: (T1,T2) -> (T1n,T2n), then swap
    void iterate() {
        const Ten t = T1, T = T2;
        const Ten tS = last_slab(t, b0, vs), T_S = last_slab(T, b0, vs);
        const Ten V_S = last_slab(V, b0, vs), Vt_S = last_slab(Vt, b0, vs), J_S = last_slab(J, b0, vs);
        // every amplitude-derived operand of the sweep in one pass over T2:
        // Tt = 2T - T(ji), tau = T + tt, tauh = T + tt/2, Tp2 = T + 2tt
        DTen tau, tauh, Tt, Tp2;
        Tt.alloc(ctx, o, o, v, v);
        Ten tauv = T, tauhv = T;
        // ... and, in the same pass, the operand layouts the ring / Fae products of this sweep want (each one
        // a permutation pass less): handed to contract() through the PermCache under the keys it will look up
        AmpExtras ex;
        DBuf xb[6];
        if (relaid && !no_amp_extras()) {
            const size_t nfull = (size_t)(o * o * v * v), nslab = (size_t)(o * o * v * vs);
            for (int q = 0; q < 6; ++q) xb[q].alloc(ctx, q < 3 ? nfull : nslab);
            ex.T_meia = xb[0].p; ex.Tt_meia = xb[1].p; ex.T_meja = xb[2].p;
            ex.T_nfjb = xb[3].p; ex.X_nfjb = xb[4].p; ex.Y_mnfa = xb[5].p;
            ex.b0 = (int)b0; ex.vs = (int)vs;
        }
        if (singles) {
            tau.alloc(ctx, o, o, v, v); tauh.alloc(ctx, o, o, v, v); Tp2.alloc(ctx, o, o, v, v);
            amp_combos(ctx, T.p, t.p, Tt.p(), tau.p(), tauh.p(), Tp2.p(), o, v, &ex);
            tauv = tau; tauhv = tauh;
        } else {
            amp_combos(ctx, T.p, nullptr, Tt.p(), nullptr, nullptr, nullptr, o, v, &ex);
        }
        const Ten tau_S = last_slab(tauv, b0, vs), tauh_S = last_slab(tauhv, b0, vs);
        pcache.add(T, true); pcache.add(Tt, true);
        if (singles) { pcache.add(tau, true); pcache.add(tauh, true); }
        if (ex.T_meia) {
            // the index strings are those of the contract() calls below
            pcache.provide(perm_key(T, "imae", "meia"), std::move(xb[0]));
            pcache.provide(perm_key(Tt, "imae", "meia"), std::move(xb[1]));
            pcache.provide(perm_key(T, "mjae", "meja"), std::move(xb[2]));
            pcache.provide(perm_key(T_S, "njfb", "nfjb"), std::move(xb[3]));
            pcache.provide(perm_key(singles ? last_slab(Tp2, b0, vs) : T_S, "jnfb", "nfjb"), std::move(xb[4]));
            pcache.provide(perm_key(tauh_S, "mnaf", "mnfa"), std::move(xb[5]));
        }

        // ---- small intermediates + T1 update: independent of the ring intermediates below, so this branch
        //      (~0.7 ms of low-occupancy kernels at config 3) is issued on the second stream and runs under the
        //      ring / ladder GEMMs; joined before the hole-hole ladder, its first consumer ------------------
        // (only for short sweeps: at nbf=300 on 2 GPUs the branch's own large GEMMs competing with the ring
        // products cost 2 %, 0.892 against 0.875 s)
        const bool overlap = relaid && short_sweeps() && ctx->arena == &arena && ctx->arena_side == &arena2 &&
                             ctx->trace < 2 && !no_overlap();
        SideScope side(*this, overlap);
        // ---- small intermediates: partial sums over f in the slab, one all-reduce -------------------
        TraceTimer* tr_small = new TraceTimer(ctx, "cc.part.small");
        const int64_t nFae = v * v, nFmi = o * o, nW = o * o * o * o, nR1 = o * v;
        DBuf small(ctx, (size_t)(nFae + nFmi + nW + nR1));
        Ten FaeT(small.p, v, v), Fmi(small.p + nFae, o, o), Wpp(small.p + nFae + nFmi, o, o, o, o),
            R1(small.p + nFae + nFmi + nW, o, v);
        contract(ctx, -1.0, tauh_S, "mnaf", Vt_S, "mnef", 0.0, FaeT, "ea");          // stored [e,a]
        contract(ctx, 1.0, Vt_S, "mnef", tauh_S, "inef", 0.0, Fmi, "mi");
        contract(ctx, 1.0, V_S, "mnef", tau_S, "ijef", 0.0, Wpp, "mnij");            // 2X
        if (singles) {
            contract(ctx, -1.0, T_S, "mnae", last_slab(ooov_t, b0, vs), "mnie", 0.0, R1, "ia");
            if (relaid) {
                contract(ctx, 1.0, OC, "amef", tS, "mf", 1.0, FaeT, "ea");
                contract(ctx, 1.0, T_S, "imef", OC, "amef", 1.0, R1, "ia");
            } else {
                contract(ctx, 2.0, OB, "amef", tS, "mf", 1.0, FaeT, "ea");
                contract(ctx, -1.0, OA, "eamf", tS, "mf", 1.0, FaeT, "ea");
                contract(ctx, 2.0, T_S, "imef", OB, "amef", 1.0, R1, "ia");
                contract(ctx, -1.0, T_S, "imef", OA, "eamf", 1.0, R1, "ia");
            }
        } else {
            fill(ctx, R1.p, (size_t)nR1, 0.0);
        }
        delete tr_small;
        { TraceTimer tt(ctx, "cc.comm.allreduce"); all_reduce_sum(ctx, small.p, small.n); }
        if (fock) {
            // one-body terms of a non-canonical reference (AutoRCCSD.jl:78-83,131-136 in the factorised
            // form of tests/factorized_model.py):  Fae[a,e] += f[e,a] - 1/2 f[m,e] t[m,a],
            // Fmi[m,i] += f[i,m] + 1/2 f[m,e] t[i,e],  Fme += f[m,e],  R1 += f[i,a]
            axpby(ctx, (size_t)nFae, 1.0, fvv.p(), 1.0, FaeT.p);
            contract(ctx, -0.5, fov, "me", t, "ma", 1.0, FaeT, "ea");
            axpby(ctx, (size_t)nFmi, 1.0, foT.p(), 1.0, Fmi.p);
            contract(ctx, 0.5, fov, "me", t, "ie", 1.0, Fmi, "mi");
            axpby(ctx, (size_t)nR1, 1.0, fov.p(), 1.0, R1.p);
        }
        axpby(ctx, (size_t)nW, 1.0, oooo.p(), 1.0, Wpp.p);
        DTen Fme, FaeT_t, Fmi_t;
        Ten FaeTt = FaeT, FmiT = Fmi;
        if (singles) {
            Fme.alloc(ctx, o, v);
            contract(ctx, 1.0, Vt, "mnef", t, "nf", 0.0, Fme, "me");
            if (fock) axpby(ctx, (size_t)nR1, 1.0, fov.p(), 1.0, Fme.p());
            contract(ctx, 1.0, ooov_t, "mnie", t, "ne", 1.0, Fmi, "mi");
            contract(ctx, 1.0, ooov, "mnie", t, "je", 1.0, Wpp, "mnij");
            contract(ctx, 1.0, oovo, "mnej", t, "ie", 1.0, Wpp, "mnij");
            // ---- T1 (RCCSD.jl:248-259) --------------------------------------------------------------
            contract(ctx, 1.0, t, "ie", FaeT, "ea", 1.0, R1, "ia");
            contract(ctx, -1.0, Fmi, "mi", t, "ma", 1.0, R1, "ia");
            contract(ctx, 1.0, Tt, "imae", Fme, "me", 1.0, R1, "ia");
            contract(ctx, 2.0, V, "imae", t, "me", 1.0, R1, "ia");
            contract(ctx, -1.0, J, "maie", t, "me", 1.0, R1, "ia");
            divide_Dia(ctx, R1.p, T1n.p(), P.eo.p, P.ev.p, o, v);
            FaeT_t.alloc(ctx, v, v); Fmi_t.alloc(ctx, o, o);
            axpby(ctx, (size_t)nFae, 1.0, FaeT.p, 0.0, FaeT_t.p());
            contract(ctx, -0.5, Fme, "me", t, "mb", 1.0, FaeT_t, "eb");
            axpby(ctx, (size_t)nFmi, 1.0, Fmi.p, 0.0, Fmi_t.p());
            contract(ctx, 0.5, Fme, "me", t, "je", 1.0, Fmi_t, "mj");
            FaeTt = FaeT_t; FmiT = Fmi_t;
        }
        side.close();
        // ---- ring intermediates for the slab, layout [m,e,j,b] ------------------------------------
        TraceTimer* tr_ring = new TraceTimer(ctx, "cc.part.ringW");
        const size_t ns = (size_t)(o * o * v * vs);
        DTen WJ(ctx, o, v, o, vs), WE(ctx, o, v, o, vs);
        // WJ starts from <mb|ej> = <mj|eb>, WE from -<mb|je> = -(mj|eb): read by the epilogue of the first
        // product into each (Cin), not copied first
        if (relaid) {
            contract(ctx, 0.5, df_wmbej ? V : Vt, "mnef", T_S, "njfb", 1.0, WJ, "mejb", false, Vx.p());
        } else {
            permute_axpby(ctx, 1.0, V_S, "mjeb", 0.0, WJ, "mejb");
            axpby(ctx, ns, -1.0, J_S.p, 0.0, WE.p());
            contract(ctx, 0.5, df_wmbej ? V : Vt, "mnef", T_S, "njfb", 1.0, WJ, "mejb");
        }
        if (singles) {
            // T/2 + tt = (T + 2 tt) / 2: one operand serves both rings
            pcache.add(Tp2, true);
            contract(ctx, -0.5, V, "mnef", last_slab(Tp2, b0, vs), "jnfb", 1.0, WJ, "mejb");
            if (relaid) contract(ctx, 0.5, V, "nmef", last_slab(Tp2, b0, vs), "jnfb", -1.0, WE, "mejb", false, J_S.p);
            else contract(ctx, 0.5, V, "nmef", last_slab(Tp2, b0, vs), "jnfb", 1.0, WE, "mejb");
            // for every b of the slab a plain matrix product [(m,e) x f][f x j]: batched, no output permutation
            contract(ctx, 1.0, OA, "efmb", t, "jf", 1.0, WJ, "mejb", relaid);
            contract(ctx, -1.0, oovo, "mnej", tS, "nb", 1.0, WJ, "mejb");
            contract(ctx, -1.0, OA, "femb", t, "jf", 1.0, WE, "mejb", relaid);
            contract(ctx, 1.0, oovo, "nmej", tS, "nb", 1.0, WE, "mejb");
        } else {
            contract(ctx, -0.5, V, "mnef", T_S, "jnfb", 1.0, WJ, "mejb");
            if (relaid) contract(ctx, 0.5, V, "nmef", T_S, "jnfb", -1.0, WE, "mejb", false, J_S.p);
            else contract(ctx, 0.5, V, "nmef", T_S, "jnfb", 1.0, WE, "mejb");
        }
        delete tr_ring;
        // ---- ladders ---------------------------------------------------------------------------------
        TraceTimer* tr_lad = new TraceTimer(ctx, "cc.part.ladders+H");
        DTen Lpp(ctx, o, o, v, vs), Lhh(ctx, o, o, v, vs), Hfull(ctx, o, o, v, v);
        const Ten H = last_slab(Hfull, b0, vs);
        DBuf Tpm, Lpm;
        if (sa_ladder) {
            // tau.vvvv through the packed symmetric / antisymmetric parts: one batched GEMM of two
            // (o^2 x np)(np x nq) products, np = v(v+1)/2 summed pairs, nq = this rank's block of the
            // ~v^2/2 output pairs; the blocks of the other ranks are all-gathered (o^2 v^2 doubles in all)
            // on the second stream, under the hole-hole ladder and the half residual below
            const int64_t oo = o * o, np = sa_pairs(v), nq_all = v * sa_slots(v);
            Tpm.alloc(ctx, (size_t)(2 * oo * sa_ld));
            Lpm.alloc(ctx, (size_t)(2 * oo * nq_all));
            pack_tau_sa(ctx, tauv.p, oo, v, sa_ld, Tpm.p);
            GemmCall g;
            g.M = oo; g.N = sa_nq; g.K = np; g.batch = 2;
            g.A = Tpm.p; g.lda = oo; g.strideA = oo * sa_ld;
            g.B = Wsa.p(); g.ldb = sa_ld; g.strideB = sa_ld * sa_nq;
            g.C = Lpm.p + oo * sa_nq * ctx->rank; g.ldc = oo; g.strideC = oo * nq_all;
            dgemm(ctx, g);
            if (ctx->nranks > 1) {
                ensure_events();
                JUES_CUDA(cudaEventRecord(ev_fork, ctx->stream));
                StreamScope sc(ctx, ctx->comm_stream);
                JUES_CUDA(cudaStreamWaitEvent(ctx->stream, ev_fork, 0));
                TraceTimer tt(ctx, "cc.comm.gatherL");
                all_gather_inplace(ctx, Lpm.p, (size_t)(oo * sa_nq));
                all_gather_inplace(ctx, Lpm.p + oo * nq_all, (size_t)(oo * sa_nq));
                JUES_CUDA(cudaEventRecord(ev_join, ctx->stream));
            }
        } else {
            contract(ctx, 1.0, tauv, "ijef", W4, "efab", 0.0, Lpp, "ijab");
        }
        if (overlap) JUES_CUDA(cudaStreamWaitEvent(ctx->stream, ev_join2, 0));        // the side branch has finished
        contract(ctx, 1.0, Wpp, "mnij", tau_S, "mnab", 0.0, Lhh, "ijab");
        // ---- half residual H for the slab (its (ij)(ab) image is added by residual_finish) ----------
        contract(ctx, 1.0, T, "ijae", last_slab(FaeTt, b0, vs), "eb", 0.0, H, "ijab");
        if (relaid) {
            // - T[i,m,a,b] Fmi[m,j] enters through its (ij)(ab) image - Fmi[m,i] T[m,j,a,b] (residual_finish adds
            // the image of everything in H): [i x m][m x (j,a,b)], no permutation on either side
            contract(ctx, -1.0, FmiT, "mi", T_S, "mjab", 1.0, H, "ijab");
            // the three ring products stay in the layouts the GEMM writes ([ia|jb], [ja|ib]); one kernel adds
            // them to H (three permute-accumulate passes over H before)
            DTen Ra(ctx, o, v, o, vs), Rb(ctx, o, v, o, vs);
            contract(ctx, 1.0, Tt, "imae", WJ, "mejb", 0.0, Ra, "iajb");
            contract(ctx, 1.0, T, "imae", WE, "mejb", 1.0, Ra, "iajb");
            contract(ctx, 1.0, T, "mjae", WE, "meib", 0.0, Rb, "jaib");   // image of T[mibe] WmBEj[maej]
            ring_combine(ctx, Ra.p(), Rb.p(), H.p, o, v, vs);
        } else {
            contract(ctx, -1.0, T_S, "imab", FmiT, "mj", 1.0, H, "ijab");
            contract(ctx, 1.0, Tt, "imae", WJ, "mejb", 1.0, H, "ijab");
            contract(ctx, 1.0, T, "imae", WE, "mejb", 1.0, H, "ijab");
            contract(ctx, 1.0, T, "mjae", WE, "meib", 1.0, H, "ijab");    // image of T[mibe] WmBEj[maej]
        }
        if (singles) {
            DTen Yp(ctx, o, o, o, vs), Q2(ctx, o, o, v, o);
            if (relaid) {
                // everything that is contracted with t[m,a] over m is summed first:
                // Yp[i,j,m,b] = <mj|ib> + tau[ijef] <ef|mb> + t[ie] <mj|eb>, then ONE product batched over b
                axpby(ctx, (size_t)(o * o * o * vs), 1.0, ooov_p.p(), 0.0, Yp.p());
                contract(ctx, 1.0, tauv, "ijef", OA, "efmb", 1.0, Yp, "ijmb");
                contract(ctx, 1.0, t, "ie", V_S, "mjeb", 1.0, Yp, "ijmb");
                contract(ctx, -1.0, Yp, "ijmb", t, "ma", 1.0, H, "ijab", true);
            } else {
                contract(ctx, 1.0, tauv, "ijef", OA, "efmb", 0.0, Yp, "ijmb");
                contract(ctx, -1.0, Yp, "ijmb", t, "ma", 1.0, H, "ijab");
                // rank-1 ring corrections  - t[ie] t[ma] <mb|ej>  - t[ie] t[mb] <am|ej>: contract t[ie] into the
                // integral first (o^3 v intermediates) instead of building v^3 o ones
                DTen Q1(ctx, o, o, o, vs);
                contract(ctx, 1.0, t, "ie", V_S, "mjeb", 0.0, Q1, "imjb");
                contract(ctx, -1.0, Q1, "imjb", t, "ma", 1.0, H, "ijab");
                contract(ctx, -1.0, t, "ma", last_slab(ooov, b0, vs), "mjib", 1.0, H, "ijab");
            }
            contract(ctx, 1.0, t, "ie", J, "maje", 0.0, Q2, "imaj");
            contract(ctx, -1.0, Q2, "imaj", tS, "mb", 1.0, H, "ijab");
            contract(ctx, 1.0, t, "ie", OB, "ajeb", 1.0, H, "ijab");     // t . <ab|ej>
        }
        if (sa_ladder) {
            if (ctx->nranks > 1) JUES_CUDA(cudaStreamWaitEvent(ctx->stream, ev_join, 0));   // ladder blocks have arrived
            unpack_ladder_sa(ctx, Lpm.p, o * o, v, b0, vs, Lpp.p());
        }
        delete tr_lad;
        { TraceTimer tt(ctx, "cc.comm.gatherH"); all_gather_inplace(ctx, Hfull.p(), ns); }
        const Ten Tn_S = last_slab(T2n, b0, vs);
        residual_finish(ctx, V_S.p, Lpp.p(), Lhh.p(), H.p, Hfull.p(), Tn_S.p, P.eo.p, P.ev.p, o, v, b0, vs);
        { TraceTimer tt(ctx, "cc.comm.gatherT2"); all_gather_inplace(ctx, T2n.p(), ns); }
        pcache.end_sweep();
        swap_amplitudes();
    }

    void swap_amplitudes() {
        std::swap(T2.buf, T2n.buf); std::swap(T2.t, T2n.t);
        if (singles) { std::swap(T1.buf, T1n.buf); std::swap(T1.t, T1n.t); }
    }

    // ---- one sweep, replayed from a CUDA graph once the launch sequence is warm ---------------------
    // A sweep is ~100 kernel launches and ~300 stream-ordered allocations; issued one by one the GPU
    // idles whenever the launching thread is descheduled for longer than the queue it has built up
    // (measured on shared hosts: single sweeps of 9 ms stretched to 15-240 ms).  After one eager sweep
    // (every lazily built operand copy, kernel attribute and pool block exists) the sweep is captured
    // once per amplitude-buffer parity (T2 -> T2n and back) and replayed with one cudaGraphLaunch, so
    // the host can run arbitrarily far ahead.  Capture failures fall back to eager launches.
    struct SweepGraph { cudaGraphExec_t exec = nullptr; Stats delta; };
    SweepGraph graphs[2];
    const double* t2_even = nullptr;
    int eager_sweeps = 0;
    bool graph_ok = true;
    // Capturing + instantiating the two graphs costs ~80 ms of host time (measured, nbf = 120), so the
    // driver turns replay on only for long fixed-length runs of launch-bound sweeps (see allow_graphs).
    bool use_graphs = false;
    // sweeps of less than ~50 ms on this rank: launch- and latency-bound kernels are a visible share
    bool short_sweeps() const {
        const double fl = 2.0 * (double)(o * o) * (double)(v * v) * (double)(v * v) +
                          22.0 * (double)(o * o * o) * (double)(v * v * v);
        return fl / (double)ctx->nranks < 1.5e12;
    }
    void allow_graphs(int sweeps) { use_graphs = sweeps >= 12 && short_sweeps(); }

    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;     // second-stream fork / join inside a sweep (ladder gather)
    cudaEvent_t ev_fork2 = nullptr, ev_join2 = nullptr;   // ... (side branch: small intermediates + T1 update)
    void ensure_events() {
        if (ev_fork) return;
        JUES_CUDA(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
        JUES_CUDA(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
        JUES_CUDA(cudaEventCreateWithFlags(&ev_fork2, cudaEventDisableTiming));
        JUES_CUDA(cudaEventCreateWithFlags(&ev_join2, cudaEventDisableTiming));
    }

    // Temporaries of a sweep (~300 stream-ordered allocations) come from one arena sized by the first sweep:
    // no driver call per temporary, and a captured sweep contains kernels only.
    Arena arena;
    DBuf arena_block;
    void install_arena(size_t need) {
        if (ctx->arena || need == 0 || need > (size_t(24) << 30)) return;
        // first-fit over blocks of very different sizes fragments: half as much again as the first sweep held
        // (twice while that is cheap: blocks of 64 MB and more next to small ones fragment a first-fit list)
        const size_t slack = need < (size_t(6) << 30) ? need : need / 2;
        const size_t bytes = ((need + slack + (size_t(16) << 20)) + 255) & ~size_t(255);
        try {
            arena_block.alloc(ctx, bytes / 8);
        } catch (const Error&) {
            cudaGetLastError();
            return;                               // no room: keep the stream-ordered pool
        }
        arena.reset(arena_block.p, bytes);
        ctx->arena = &arena;
        ctx->arena_main = &arena;
        // the side branch of the sweep (small intermediates + T1 update, issued on the second stream under the
        // ring-intermediate GEMMs) takes its temporaries from an arena of its own
        if (side_peak == 0 || no_overlap()) return;
        const size_t b2 = ((side_peak + side_peak / 2 + (size_t(16) << 20)) + 255) & ~size_t(255);
        try {
            ArenaPause ap(ctx);
            arena2_block.alloc(ctx, b2 / 8);
        } catch (const Error&) {
            cudaGetLastError();
            return;
        }
        arena2.reset(arena2_block.p, b2);
        ctx->arena_side = &arena2;
    }
    Arena arena2;
    DBuf arena2_block;
    size_t side_peak = 0;          // what the side branch held at once in the measured first sweep
    static bool no_overlap() {
        static const bool off = getenv("JUES_B200_NO_OVERLAP") != nullptr;       // A/B switch for measurements
        return off;
    }
    // The side branch runs on ctx->comm_stream from its own arena; misses go to the stream-ordered pool.
    struct SideScope {
        CC& cc;
        bool on;
        cudaStream_t saved_stream = nullptr;
        Arena* saved_arena = nullptr;
        bool saved_nbc = false;
        size_t live0 = 0, peak0 = 0;
        bool measuring = false;
        SideScope(CC& c, bool enable) : cc(c), on(enable) {
            jues_ctx* ctx = cc.ctx;
            measuring = ctx->measure;
            if (measuring) { live0 = ctx->temp_live; peak0 = ctx->temp_peak; ctx->temp_peak = live0; }
            if (!on) return;
            cc.ensure_events();
            JUES_CUDA(cudaEventRecord(cc.ev_fork2, ctx->stream));
            saved_stream = ctx->stream; saved_arena = ctx->arena; saved_nbc = ctx->no_big_cache;
            ctx->stream = ctx->comm_stream;
            ctx->arena = &cc.arena2;
            ctx->no_big_cache = true;
            JUES_CUDA(cudaStreamWaitEvent(ctx->stream, cc.ev_fork2, 0));
        }
        bool closed = false;
        void close() {
            if (closed) return;
            closed = true;
            jues_ctx* ctx = cc.ctx;
            if (measuring) {
                cc.side_peak = std::max(cc.side_peak, ctx->temp_peak - live0);
                ctx->temp_peak = std::max(peak0, ctx->temp_peak);
            }
            if (!on) return;
            cudaEventRecord(cc.ev_join2, ctx->stream);
            ctx->stream = saved_stream; ctx->arena = saved_arena;
            ctx->no_big_cache = saved_nbc;
            on = false;
        }
        ~SideScope() { close(); }
    };

    void sweep() {
        static const bool no_graph = getenv("JUES_B200_NO_GRAPH") != nullptr;
        static const bool no_arena = getenv("JUES_B200_NO_ARENA") != nullptr;
        const bool off = no_graph || ctx->trace > 0;
        if (eager_sweeps < 1) {
            // the first sweep runs eagerly from the pool: it builds every lazily created operand copy and
            // kernel attribute, and measures what the temporaries of a sweep need
            ctx->temp_live = ctx->temp_peak = 0;
            ctx->measure = true;
            try {
                iterate();
            } catch (...) {
                ctx->measure = false;
                throw;
            }
            ctx->measure = false;
            ++eager_sweeps;
            if (!no_arena) install_arena(ctx->temp_peak);
            return;
        }
        if (off || !use_graphs || !graph_ok) {
            iterate();
            ++eager_sweeps;
            return;
        }
        if (!t2_even) t2_even = T2.p();
        SweepGraph& g = graphs[T2.p() == t2_even ? 0 : 1];
        if (g.exec) {
            swap_amplitudes();              // what iterate() does on the host side
        } else {
            const Stats before = ctx->stats;
            const long long big0 = ctx->big_allocs;
            cudaGraph_t graph = nullptr;
            bool captured = false;
            if (cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeRelaxed) == cudaSuccess) {
                bool threw = false;
                // a temporary that does not fit the arena while capturing (first-fit fragmentation with blocks of
                // 64 MB and more: seen at the 8-GPU bench shape) comes from the stream-ordered pool, i.e. becomes
                // an allocation node owned by the graph -- never a cached cudaMalloc block, whose address the
                // cache may hand to somebody else between replays
                const bool nbc = ctx->no_big_cache;
                ctx->no_big_cache = true;
                try {
                    iterate();
                } catch (const Error&) {
                    threw = true;
                }
                ctx->no_big_cache = nbc;
                const cudaError_t e = cudaStreamEndCapture(ctx->stream, &graph);
                // a sweep that took blocks from the big-block cache is not replayed: the graph would keep
                // using addresses that the cache may hand to somebody else
                if (!threw && e == cudaSuccess && graph && ctx->big_allocs == big0 &&
                    cudaGraphInstantiate(&g.exec, graph, 0) == cudaSuccess)
                    captured = true;
                if (graph) cudaGraphDestroy(graph);
                if (!captured) {
                    cudaGetLastError();
                    g.exec = nullptr;
                    if (!threw) swap_amplitudes();   // undo the host-side swap of the sweep that never ran
                    pcache.end_sweep();
                }
            } else {
                cudaGetLastError();
            }
            if (!captured) {
                graph_ok = false;
                ctx->stats = before;
                iterate();
                return;
            }
            g.delta.gemm_flops = ctx->stats.gemm_flops - before.gemm_flops;
            g.delta.gemm_launches = ctx->stats.gemm_launches - before.gemm_launches;
            g.delta.aux_launches = ctx->stats.aux_launches - before.aux_launches;
            g.delta.collectives = ctx->stats.collectives - before.collectives;
            g.delta.collective_bytes = ctx->stats.collective_bytes - before.collective_bytes;
            ctx->stats = before;
        }
        JUES_CUDA(cudaGraphLaunch(g.exec, ctx->stream));
        ctx->stats.gemm_flops += g.delta.gemm_flops;
        ctx->stats.gemm_launches += g.delta.gemm_launches;
        ctx->stats.aux_launches += g.delta.aux_launches;
        ctx->stats.collectives += g.delta.collectives;
        ctx->stats.collective_bytes += g.delta.collective_bytes;
        ctx->stats.graph_launches += 1;
    }

    // Off-diagonal Fock blocks (host, unpadded, diagonals already removed by the caller):
    // foo (nocc,nocc) indexed [i,k], fov (nocc,nvir) [k,c], fvv (nvir,nvir) [c,a] as AutoRCCSD.jl uses them.
    void set_fock(const double* foo, const double* fov_h, const double* fvv_h) {
        JUES_REQUIRE(singles, "off-diagonal Fock terms need the singles equations");
        JUES_REQUIRE(foo && fov_h && fvv_h, "null Fock block");
        const int64_t no = P.nocc, nv = P.nvir;
        std::vector<double> fT((size_t)(no * no));
        for (int64_t i = 0; i < no; ++i)
            for (int64_t m = 0; m < no; ++m) fT[m + no * i] = foo[i + no * m];   // foT[m,i] = f[i,m]
        DBuf a, b, c;
        upload_padded_matrix(ctx, a, fT.data(), no, no, o, o);
        upload_padded_matrix(ctx, b, fov_h, no, nv, o, v);
        upload_padded_matrix(ctx, c, fvv_h, nv, nv, v, v);
        JUES_CUDA(cudaStreamSynchronize(ctx->stream));   // fT is a stack-owned vector
        foT.buf = std::move(a); foT.t = Ten(foT.buf.p, o, o);
        fov.buf = std::move(b); fov.t = Ten(fov.buf.p, o, v);
        fvv.buf = std::move(c); fvv.t = Ten(fvv.buf.p, v, v);
        fock = true;
    }

    // E(T) from the current amplitudes (PerturbativeTriples.jl:35-138 through pt.cu).  Destroys the
    // integral classes the triples do not read; call after the last sweep.
    double triples() {
        JUES_REQUIRE(singles, "(T) needs the RCCSD integral classes");
        pcache.clear();
        OB.release(); J.release(); oooo.release(); Vt.release(); oovo.release(); ooov_t.release();
        T2n.release();
        DBuf scratch = std::move(sa_ladder ? Wsa.buf : W4.buf);     // <vv|vv> becomes the triples' work space
        const int64_t K = v + o;
        TraceTimer* tp = new TraceTimer(ctx, "pt.prepare");
        DTen Acat(ctx, v, v, o, K), Bq(ctx, v, o, K, o), Br(ctx, v, o, K, o), Vv(ctx, v, v, o, o);
        // OA[e,f,m,b] = <ef|mb> = <mb|ef> is the kap < v block of Acat[a,b,p,kap]; with several ranks each
        // holds the slab b in S_r of it, which is all-gathered in place
        const size_t cnt = (size_t)(v * v * o * vs);
        JUES_CUDA(cudaMemcpyAsync(Acat.p() + (size_t)ctx->rank * cnt, OA.p(), cnt * sizeof(double),
                                  cudaMemcpyDeviceToDevice, ctx->stream));
        all_gather_inplace(ctx, Acat.p(), cnt);
        OA.release();
        pt_build_operands(ctx, o, v, nullptr, T2.p(), ooov.p(), Acat.p(), Bq.p(), Br.p());
        permute_axpby(ctx, 1.0, V, "ijab", 0.0, Vv, "abij");
        delete tp;
        PtInputs in;
        in.o = o; in.v = v; in.nocc = P.nocc;
        in.Acat = Acat.p(); in.Bq = Bq.p(); in.Br = Br.p(); in.Vv = Vv.p(); in.t1 = T1.p();
        in.eo = P.eo.p; in.ev = P.ev.p;
        return pt_dev(ctx, in, scratch.p, scratch.n);
    }

    // host copies in the caller's (unpadded) layout
    void download(double* T1_host, double* T2_host) {
        if (T2_host) {
            const int64_t sd[4] = {o, o, v, v}, dd[4] = {P.nocc, P.nocc, P.nvir, P.nvir};
            DBuf tmp(ctx, (size_t)(dd[0] * dd[1] * dd[2] * dd[3]));
            block_copy(ctx, T2.p(), sd, tmp.p, dd, dd);
            JUES_CUDA(cudaMemcpyAsync(T2_host, tmp.p, tmp.n * 8, cudaMemcpyDeviceToHost, ctx->stream));
            JUES_CUDA(cudaStreamSynchronize(ctx->stream));
        }
        if (T1_host) {
            const int64_t sd[4] = {o, v, 1, 1}, dd[4] = {P.nocc, P.nvir, 1, 1};
            DBuf tmp(ctx, (size_t)(dd[0] * dd[1]));
            block_copy(ctx, T1.p(), sd, tmp.p, dd, dd);
            JUES_CUDA(cudaMemcpyAsync(T1_host, tmp.p, tmp.n * 8, cudaMemcpyDeviceToHost, ctx->stream));
            JUES_CUDA(cudaStreamSynchronize(ctx->stream));
        }
    }
};

}  // namespace

static CCResult cc_run(jues_ctx* ctx, Problem& P, CC& cc, bool singles, int maxit, int guess_mode, double* T1_out,
                       double* T2_out, jues_b200_amp_cb cb, void* cb_user);

CCResult cc_dev(jues_ctx* ctx, Problem& P, GaoSource& gao, bool singles, int maxit, int guess_mode,
                double* T1_out, double* T2_out, jues_b200_amp_cb cb, void* cb_user) {
    JUES_REQUIRE(maxit >= 0, "maxit must be non-negative");
    CC cc(ctx, P, singles);
    cc.build_integrals(gao);
    return cc_run(ctx, P, cc, singles, maxit, guess_mode, T1_out, T2_out, cb, cb_user);
}

// ---------------------------------------------------------------------------------------------
// Density fitting (DF.jl:52-58, DF-RMP2.jl:1-46, DF-RCCD.jl:11-89)
// ---------------------------------------------------------------------------------------------
namespace {

// b[p,q,Q] = C1[mu,p] C2[nu,q] pqP[mu,nu,P] Jpqh[P,Q] for the orbital blocks the caller asks for.  The
// reference multiplies by Jpqh first (N^2 naux^2); here the orbital coefficients go first (the same sums in
// a cheaper order: o N^2 naux + ... + o v naux^2).
struct DFTensors {
    DTen bov, boo, bvv;
    int64_t nx = 0;
};

void df_tensors(jues_ctx* ctx, Problem& P, const double* pqP, int64_t naux, const double* Jpqh, bool all_blocks,
                DFTensors& B) {
    JUES_REQUIRE(pqP && Jpqh, "density fitting: null pqP / Jpqh");
    JUES_REQUIRE(naux > 0, "density fitting: naux must be positive");
    const int64_t n = P.nao, np = P.np, o = P.o, v = P.v, nx = round_up(naux, 2);
    B.nx = nx;
    DTen pq(ctx, np, np, nx), Jh(ctx, nx, nx);
    pq.buf.zero(); Jh.buf.zero();
    for (int64_t q = 0; q < naux; ++q)       // one (nao x nao) plane per auxiliary function into the padded block
        JUES_CUDA(cudaMemcpy2DAsync(pq.p() + np * np * q, (size_t)np * 8, pqP + n * n * q, (size_t)n * 8, (size_t)n * 8,
                                    (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    JUES_CUDA(cudaMemcpy2DAsync(Jh.p(), (size_t)nx * 8, Jpqh, (size_t)naux * 8, (size_t)naux * 8, (size_t)naux,
                                cudaMemcpyHostToDevice, ctx->stream));
    JUES_CUDA(cudaStreamSynchronize(ctx->stream));     // the caller's arrays are not retained
    const Ten Co(P.Co.p, np, o), Cv(P.Cv.p, np, v);
    DTen X1(ctx, o, np, nx);
    contract(ctx, 1.0, Co, "mi", pq, "mnP", 0.0, X1, "inP");
    {
        DTen Y(ctx, o, v, nx);
        contract(ctx, 1.0, X1, "inP", Cv, "na", 0.0, Y, "iaP", true);
        B.bov.alloc(ctx, o, v, nx);
        contract(ctx, 1.0, Y, "iaP", Jh, "PQ", 0.0, B.bov, "iaQ");
    }
    if (!all_blocks) return;
    {
        DTen Y(ctx, o, o, nx);
        contract(ctx, 1.0, X1, "inP", Co, "nj", 0.0, Y, "ijP", true);
        B.boo.alloc(ctx, o, o, nx);
        contract(ctx, 1.0, Y, "ijP", Jh, "PQ", 0.0, B.boo, "ijQ");
    }
    X1.release();
    DTen X2(ctx, v, np, nx), Y(ctx, v, v, nx);
    contract(ctx, 1.0, Cv, "ma", pq, "mnP", 0.0, X2, "anP");
    contract(ctx, 1.0, X2, "anP", Cv, "nb", 0.0, Y, "abP", true);
    B.bvv.alloc(ctx, v, v, nx);
    contract(ctx, 1.0, Y, "abP", Jh, "PQ", 0.0, B.bvv, "abQ");
}

}  // namespace

double df_rmp2_dev(jues_ctx* ctx, Problem& P, const double* pqP, int64_t naux, const double* Jpqh) {
    const int64_t o = P.o, v = P.v;
    DFTensors B;
    DTen ijab(ctx, o, o, v, v);
    {
        Timer t(ctx, "mp2.transform");
        df_tensors(ctx, P, pqP, naux, Jpqh, false, B);
        contract(ctx, 1.0, B.bov, "iaQ", B.bov, "jbQ", 0.0, ijab, "ijab");     // (ia|jb), DF-RMP2.jl:33-35
    }
    // every rank of a multi-GPU context forms the whole (small) sum: no exchange
    Timer t(ctx, "mp2.energy");
    return mp2_energy(ctx, ijab.p(), P.eo.p, P.ev.p, o, v, 0, v);
}

CCResult df_rccd_dev(jues_ctx* ctx, Problem& P, const double* pqP, int64_t naux, const double* Jpqh, int maxit,
                     double* T2_out, jues_b200_amp_cb cb, void* cb_user) {
    JUES_REQUIRE(maxit >= 0, "maxit must be non-negative");
    CC cc(ctx, P, false);
    cc.df_wmbej = true;
    {
        DFTensors B;
        df_tensors(ctx, P, pqP, naux, Jpqh, true, B);
        cc.build_integrals_df(B.bov, B.boo, B.bvv);
    }
    return cc_run(ctx, P, cc, false, maxit, 1, nullptr, T2_out, cb, cb_user);   // T2_init!: the MP2 guess (:111-135)
}

static CCResult cc_run(jues_ctx* ctx, Problem& P, CC& cc, bool singles, int maxit, int guess_mode, double* T1_out,
                       double* T2_out, jues_b200_amp_cb cb, void* cb_user) {
    cc.register_static();
    cc.guess(guess_mode);
    if (!cb) cc.allow_graphs(maxit);
    CCResult res;
    res.e_hist.resize(maxit + 1);
    std::vector<double> h1, h2;
    auto report = [&](int it, double e) {
        if (!cb) return;
        h2.resize((size_t)(P.nocc * P.nocc * P.nvir * P.nvir));
        if (singles) h1.resize((size_t)(P.nocc * P.nvir));
        cc.download(singles ? h1.data() : nullptr, h2.data());
        cb(cb_user, it, e, singles ? h1.data() : nullptr, h2.data());
    };
    // Energies are reduced on the device into e_dev[it] and read back once at the end, so the host
    // never waits for a sweep (unless a per-sweep amplitude callback wants the values): the launch
    // queue stays full across sweeps.
    DBuf e_dev(ctx, (size_t)maxit + 1);
    cc.energy_async(e_dev.p);
    const auto host_t0 = std::chrono::steady_clock::now();
    if (cb) {
        res.e_hist[0] = cc.energy();
        report(0, res.e_hist[0]);
    }
    for (int it = 1; it <= maxit; ++it) {
        {
            // one timed "step" = one sweep + the energy the reference evaluates every sweep
            // (RCCSD.jl:104)
            const double f0 = ctx->stats.gemm_flops;
            Timer t(ctx, "cc.iteration");
            cc.sweep();
            cc.energy_async(e_dev.p + it);
            t.stop();
            // FP64 flops this rank's GEMM launches executed in the sweep (reported as a pseudo-phase)
            ctx->timings.emplace_back("cc.iteration.gflop", (float)((ctx->stats.gemm_flops - f0) * 1e-9));
            // host clock when this sweep had been handed to the driver (diagnostic: how far the launching
            // thread runs ahead of the GPU)
            ctx->timings.emplace_back("cc.iteration.host_ms",
                (float)(std::chrono::duration<double>(std::chrono::steady_clock::now() - host_t0).count() * 1e3));
        }
        if (cb) {
            res.e_hist[it] = cc.energy();
            report(it, res.e_hist[it]);
        }
    }
    JUES_CUDA(cudaMemcpyAsync(res.e_hist.data(), e_dev.p, ((size_t)maxit + 1) * sizeof(double),
                              cudaMemcpyDeviceToHost, ctx->stream));
    JUES_CUDA(cudaStreamSynchronize(ctx->stream));
    res.energy = res.e_hist[maxit];
    ctx->timings.emplace_back("cc.graph_launches", (float)ctx->stats.graph_launches);
    ctx->timings.emplace_back("cc.arena_mb", (float)(cc.arena.bytes / 1048576.0));
    ctx->timings.emplace_back("cc.arena_misses", (float)cc.arena.misses);
    cc.download(singles ? T1_out : nullptr, T2_out);
    return res;
}

// AutoRCCSD.do_rccsd (AutoRCCSD.jl:193-301): RCCSD for a possibly non-canonical reference with real
// convergence control.  P.eo / P.ev hold the DIAGONAL of the Fock matrix (the resolvents d, D of
// :242-244); foo/fov/fvv its off-diagonal blocks.  The sweep is the same factorised sweep as
// do_rccsd plus the one-body terms; per sweep the host reads back three doubles (energy and the two
// squared amplitude changes) to take the reference's stop decision (:270-285).
AutoResult auto_rccsd_dev(jues_ctx* ctx, Problem& P, GaoSource& gao, const double* foo, const double* fov,
                          const double* fvv, const AutoOptions& opt, double* T1_out, double* T2_out,
                          jues_b200_amp_cb cb, void* cb_user) {
    JUES_REQUIRE(opt.max_iter >= 0, "cc_max_iter must be non-negative");
    CC cc(ctx, P, true);
    cc.set_fock(foo, fov, fvv);
    cc.build_integrals(gao);
    cc.register_static();
    cc.guess(1);                                                              // T2 = <ij|ab>/D (:247)
    divide_Dia(ctx, cc.fov.p(), cc.T1.p(), P.eo.p, P.ev.p, cc.o, cc.v);       // T1 = f_ov/d    (:246)
    const size_t n1 = (size_t)(cc.o * cc.v), n2 = (size_t)(cc.o * cc.o * cc.v * cc.v);
    const double len1 = (double)(P.nocc * P.nvir), len2 = (double)(P.nocc * P.nocc) * (double)(P.nvir * P.nvir);
    DBuf scal(ctx, 4);
    double h[3] = {0, 0, 0};
    auto energy = [&]() {
        // update_energy (:40-52):  2 f_kc t_kc + sum <kl|cd> (2 tau_klcd - tau_lkcd)
        cc.energy_async(scal.p);
        dot_axpby(ctx, n1, 2.0, cc.fov.p(), cc.T1.p(), 1.0, scal.p);
    };
    auto read = [&](int n) {
        JUES_CUDA(cudaMemcpyAsync(h, scal.p, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        JUES_CUDA(cudaStreamSynchronize(ctx->stream));
    };
    std::vector<double> h1, h2;
    auto report = [&](int it, double e) {
        if (!cb) return;
        h2.resize((size_t)(P.nocc * P.nocc * P.nvir * P.nvir));
        h1.resize((size_t)(P.nocc * P.nvir));
        cc.download(h1.data(), h2.data());
        cb(cb_user, it, e, h1.data(), h2.data());
    };
    AutoResult res;
    energy();
    read(1);
    double Ecc = h[0];
    res.e_hist.push_back(Ecc);
    res.rms_hist.push_back(1.0);
    report(0, Ecc);
    double dE = 1.0, rms = 1.0;
    int ite = 1;
    // One sweep + its three scalars [E, |dT1|^2, |dT2|^2] into scalars[3*(k%2)..], copied to pinned host
    // memory; `done[k%2]` fires when they have landed.
    DBuf sc2(ctx, 8);
    double* hp = ctx->red_host + 8;                    // pinned landing zone: 2 x 3 doubles
    struct Events {                                    // destroyed on every exit path
        cudaEvent_t e[2] = {nullptr, nullptr};
        ~Events() { for (auto x : e) if (x) cudaEventDestroy(x); }
        cudaEvent_t& operator[](int k) { return e[k]; }
    } done;
    JUES_CUDA(cudaEventCreateWithFlags(&done[0], cudaEventDisableTiming));
    JUES_CUDA(cudaEventCreateWithFlags(&done[1], cudaEventDisableTiming));
    auto enqueue = [&](int k) {
        const double f0 = ctx->stats.gemm_flops;
        Timer t(ctx, "cc.iteration");
        cc.sweep();                         // afterwards T1/T2 are the new, T1n/T2n the old amplitudes
        double* s3 = sc2.p + 3 * (k & 1);
        sqdiff_async(ctx, n1, cc.T1.p(), cc.T1n.p(), s3 + 1);                 // :174-175
        sqdiff_async(ctx, n2, cc.T2.p(), cc.T2n.p(), s3 + 2);
        cc.energy_async(s3);
        dot_axpby(ctx, n1, 2.0, cc.fov.p(), cc.T1.p(), 1.0, s3);
        t.stop();
        ctx->timings.emplace_back("cc.iteration.gflop", (float)((ctx->stats.gemm_flops - f0) * 1e-9));
        JUES_CUDA(cudaMemcpyAsync(hp + 3 * (k & 1), s3, 3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        JUES_CUDA(cudaEventRecord(done[k & 1], ctx->stream));
    };
    // The reference's loop (:270-285) decides after every sweep.  Without a per-sweep amplitude
    // callback the NEXT sweep is enqueued before the host waits for the scalars of the current one,
    // so the launch queue never drains; when the current sweep turns out to be the last, the
    // speculative one is discarded (the amplitudes it started from are still in T1n/T2n).
    const bool speculate = (cb == nullptr);
    bool in_flight = false;                 // sweep `ite` already enqueued
    while (std::fabs(dE) > opt.e_conv || rms > opt.max_rms) {                 // :270
        if (ite > opt.max_iter) break;                                        // :271-274
        if (!in_flight) enqueue(ite);
        in_flight = false;
        if (speculate && ite + 1 <= opt.max_iter) {
            enqueue(ite + 1);
            in_flight = true;
        }
        JUES_CUDA(cudaEventSynchronize(done[ite & 1]));
        const double* hk = hp + 3 * (ite & 1);
        rms = std::max(std::sqrt(hk[1]) / len1, std::sqrt(hk[2]) / len2);     // :278
        const double oldE = Ecc;
        Ecc = hk[0];
        dE = Ecc - oldE;
        res.e_hist.push_back(Ecc);
        res.rms_hist.push_back(rms);
        report(ite, Ecc);
        ++ite;
    }
    if (in_flight) {
        // discard the speculative sweep: swap the previous amplitudes back
        std::swap(cc.T2.buf, cc.T2n.buf); std::swap(cc.T2.t, cc.T2n.t);
        std::swap(cc.T1.buf, cc.T1n.buf); std::swap(cc.T1.t, cc.T1n.t);
        ctx->timings.emplace_back("cc.speculative_sweeps", 1.0f);
    }
    JUES_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->timings.emplace_back("cc.graph_launches", (float)ctx->stats.graph_launches);
    res.iterations = ite - 1;
    res.converged = std::fabs(dE) < opt.e_conv && rms < opt.max_rms;          // :288
    res.ecc = Ecc;
    cc.download(T1_out, T2_out);
    if (opt.do_pT) {                                                          // :294-300
        Timer t(ctx, "cc.triples");
        res.ept = cc.triples();
        res.has_pt = true;
    }
    return res;
}

// mRCCD.do_rccd (mRCCD.jl:37-120, cciter with DIIS :143-207): RCCD from ZERO amplitudes (T2_init!,
// :239-246) with Pulay extrapolation over at most five (amplitude, error) pairs that the reference
// stores in Float32 (:64-65,171,175).  B[n1,n2] = <e_n1|e_n2> normalised by max|B| (:188-194), the
// coefficients solve B c = (0,...,0,-1) in Float32 (:195-198), the new amplitudes are
// sum_k Float32(c_k) * Float32(T_k) accumulated in Float64 (:199-202).  Stops when the 2-norm of the
// un-extrapolated change drops below 1e-7 (:106,205-206) or after maxit sweeps.
// The sweep itself is the factorised RCCD sweep above (equal to mRCCD's GEMM chain).
MrccdResult mrccd_dev(jues_ctx* ctx, Problem& P, GaoSource& gao, int maxit, double* T2_out,
                      jues_b200_amp_cb cb, void* cb_user) {
    JUES_REQUIRE(maxit >= 0, "maxit must be non-negative");
    CC cc(ctx, P, false);
    cc.build_integrals(gao);
    cc.register_static();
    cc.guess(1);
    cc.T2.buf.zero();                                   // zeros ./ Dijab
    const size_t n2 = (size_t)(cc.o * cc.o * cc.v * cc.v);   // even: every extent is even
    constexpr int kMaxPairs = 5;                        // max_diis = 6 vectors, the oldest never used (:179-184)
    struct Pair { DBuf val, err; };
    std::vector<Pair> pairs;
    std::vector<std::vector<float>> dots;               // dots[a][b] = Float32(<e_a|e_b>)
    DBuf scal(ctx, 8);
    std::vector<double> h2;
    MrccdResult res;
    for (int it = 0; it < maxit; ++it) {
        double hbuf[8];
        {
            Timer t(ctx, "cc.iteration");
            cc.sweep();                                 // T2 = un-extrapolated new amplitudes, T2n = old ones
            Pair p;
            {
                ArenaPause keep(ctx);                   // DIIS vectors live across sweeps
                p.val.alloc(ctx, n2 / 2 + 1);
                p.err.alloc(ctx, n2 / 2 + 1);
            }
            to_float32(ctx, n2, cc.T2.p(), nullptr, reinterpret_cast<float*>(p.val.p));
            to_float32(ctx, n2, cc.T2.p(), cc.T2n.p(), reinterpret_cast<float*>(p.err.p));
            sqdiff_async(ctx, n2, cc.T2.p(), cc.T2n.p(), scal.p);          // ||T2new - T2old||^2 in FP64 (:205)
            pairs.push_back(std::move(p));
            if ((int)pairs.size() > kMaxPairs) {
                pairs.erase(pairs.begin());
                dots.erase(dots.begin());
                for (auto& row : dots) row.erase(row.begin());
            }
            const int n = (int)pairs.size();
            for (int a = 0; a < n; ++a)
                dot_float32_async(ctx, n2, reinterpret_cast<const float*>(pairs[a].err.p),
                                  reinterpret_cast<const float*>(pairs[n - 1].err.p), scal.p + 1 + a);
            JUES_CUDA(cudaMemcpyAsync(hbuf, scal.p, (size_t)(n + 1) * sizeof(double), cudaMemcpyDeviceToHost,
                                      ctx->stream));
            JUES_CUDA(cudaStreamSynchronize(ctx->stream));
            for (auto& row : dots) row.push_back(0.0f);
            dots.emplace_back((size_t)n, 0.0f);
            for (int a = 0; a < n; ++a) dots[a][n - 1] = dots[n - 1][a] = (float)hbuf[1 + a];
            // B (n+1 x n+1, Float32), normalised, solved by Gaussian elimination with partial pivoting
            const int m = n + 1;
            std::vector<float> B((size_t)m * m, -1.0f), rhs((size_t)m, 0.0f);
            B[(size_t)m * m - 1] = 0.0f;
            float bmax = 0.0f;
            for (int a = 0; a < n; ++a)
                for (int b = 0; b < n; ++b) bmax = std::max(bmax, std::fabs(dots[a][b]));
            for (int a = 0; a < n; ++a)
                for (int b = 0; b < n; ++b) B[a + (size_t)m * b] = bmax > 0.0f ? dots[a][b] / bmax : dots[a][b];
            rhs[m - 1] = -1.0f;
            for (int c = 0; c < m; ++c) {
                int piv = c;
                for (int r = c + 1; r < m; ++r)
                    if (std::fabs(B[r + (size_t)m * c]) > std::fabs(B[piv + (size_t)m * c])) piv = r;
                if (piv != c) {
                    for (int k = 0; k < m; ++k) std::swap(B[c + (size_t)m * k], B[piv + (size_t)m * k]);
                    std::swap(rhs[c], rhs[piv]);
                }
                const float d = B[c + (size_t)m * c];
                JUES_REQUIRE(d != 0.0f, "mRCCD: singular DIIS matrix");
                for (int r = c + 1; r < m; ++r) {
                    const float f = B[r + (size_t)m * c] / d;
                    for (int k = c; k < m; ++k) B[r + (size_t)m * k] -= f * B[c + (size_t)m * k];
                    rhs[r] -= f * rhs[c];
                }
            }
            std::vector<float> ci((size_t)m, 0.0f);
            for (int r = m - 1; r >= 0; --r) {
                float acc = rhs[r];
                for (int k = r + 1; k < m; ++k) acc -= B[r + (size_t)m * k] * ci[k];
                ci[r] = acc / B[r + (size_t)m * r];
            }
            const float* vecs[8];
            for (int a = 0; a < n; ++a) vecs[a] = reinterpret_cast<const float*>(pairs[a].val.p);
            diis_combine(ctx, n2, n, vecs, ci.data(), cc.T2.p());
        }
        const double r2 = std::sqrt(hbuf[0]);
        res.rms_hist.push_back(r2);
        res.e_hist.push_back(cc.energy());
        res.iterations = it + 1;
        if (cb) {
            h2.resize((size_t)(P.nocc * P.nocc * P.nvir * P.nvir));
            cc.download(nullptr, h2.data());
            cb(cb_user, it + 1, res.e_hist.back(), nullptr, h2.data());
        }
        if (r2 < 1e-7) break;                                                  // :106
    }
    res.energy = cc.energy();                                                  // :116-118
    cc.download(nullptr, T2_out);
    return res;
}

}  // namespace jues
