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
    // self-contained, so the energy needs no exchange but the final scalar.
    DTen ijab;
    if (gao.resident()) {
        // resident AO tensor: transform only this rank's slab, (ia|j b_S)
        DTen t4(ctx, o, v, o, vs);
        {
            Timer t(ctx, "mp2.transform");
            const double* Cm[4] = {P.Co.p, P.Cv.p, P.Co.p, P.Cv.p + b0 * gao.np};
            const int64_t dp[4] = {o, v, o, vs};
            tei_transform_dev(ctx, gao, Cm, dp, t4.p());
        }
        Timer t(ctx, "mp2.energy");
        ijab.alloc(ctx, o, o, v, vs);
        permute_axpby(ctx, 1.0, t4, "iajb", 0.0, ijab, "ijab");   // <ij|ab> (IntegralTransformation.jl:96-98)
        return all_reduce_scalar(ctx, mp2_energy(ctx, ijab.p(), P.eo.p, P.ev.p, o, v, b0, vs));
    }
    // streamed AO tensor: contracted over sigma first with the occupied index in the last slot
    // ((ia|jb) = (ia|bj): an N^3 x o accumulator instead of N^3 x v).  The transform is linear in
    // gao, so each rank streams only its share of the sigma range -- generation and the dominant
    // first quarter are divided by the number of ranks -- and the partial (ia|bj) tensors are summed.
    DTen t4(ctx, o, v, v, o);
    {
        Timer t(ctx, "mp2.transform");
        if (ctx->nranks > 1) {
            const int64_t per = round_up((gao.np + ctx->nranks - 1) / ctx->nranks, 2);
            gao.sig_lo = std::min<int64_t>(gao.np, per * ctx->rank);
            gao.sig_hi = std::min<int64_t>(gao.np, per * (ctx->rank + 1));
        }
        const double* Cm[4] = {P.Co.p, P.Cv.p, P.Cv.p, P.Co.p};
        const int64_t dp[4] = {o, v, v, o};
        tei_transform_dev(ctx, gao, Cm, dp, t4.p());
        gao.sig_lo = 0; gao.sig_hi = -1;
        all_reduce_sum(ctx, t4.p(), (size_t)t4.t.size());
    }
    Timer t(ctx, "mp2.energy");
    ijab.alloc(ctx, o, o, v, v);
    permute_axpby(ctx, 1.0, t4, "iabj", 0.0, ijab, "ijab");
    const double e = mp2_energy(ctx, ijab.p() + b0 * o * o * v, P.eo.p, P.ev.p, o, v, b0, vs);
    return all_reduce_scalar(ctx, e);
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
    // static combinations
    DTen Vt, oovo, ooov_t;
    // amplitudes (replicated)
    DTen T1, T2, T1n, T2n;
    TransformWorkspace tws;   // shared by the class transforms, released before the sweeps
    PermCache pcache;         // permuted operand copies (static: once; amplitude-derived: once per sweep)

    CC(jues_ctx* c, Problem& p, bool s) : ctx(c), P(p), singles(s), o(p.o), v(p.v) { slab_of(c, v, &b0, &vs); }

    void klass(GaoSource& gphys, DTen& out, const char* slots, bool slab4) {
        const double* Cm[4];
        int64_t dp[4];
        for (int q = 0; q < 4; ++q) {
            Cm[q] = slots[q] == 'o' ? P.Co.p : P.Cv.p;
            dp[q] = slots[q] == 'o' ? o : v;
        }
        if (slab4) {
            Cm[3] = P.Cv.p + b0 * gphys.np;   // columns [b0, b0+vs) of Cav
            dp[3] = vs;
        }
        out.alloc(ctx, dp[0], dp[1], dp[2], dp[3]);
        if (gphys.resident()) {
            tei_transform_dev(ctx, gphys, Cm, dp, out.p(), false, &tws);
            return;
        }
        // A streamed AO tensor must be contracted over its last index first, and the first quarter's
        // output is N^3 x d4: use <pq|rs> = <rq|ps> = <ps|rq> = <qp|sr> (8 slot orders in all) to put
        // the SMALLEST extent in the last slot, transform, and permute back.
        static const int sym[8][4] = {{0, 1, 2, 3}, {0, 3, 2, 1}, {1, 0, 3, 2}, {1, 2, 3, 0},
                                      {2, 1, 0, 3}, {2, 3, 0, 1}, {3, 0, 1, 2}, {3, 2, 1, 0}};
        int best = 0;
        for (int g = 1; g < 8; ++g)
            if (dp[sym[g][3]] < dp[sym[best][3]]) best = g;
        if (best == 0) {
            tei_transform_dev(ctx, gphys, Cm, dp, out.p(), false, &tws);
            return;
        }
        const double* Cm2[4];
        int64_t dp2[4];
        char ly[5] = {0, 0, 0, 0, 0};
        const char lx[5] = "pqrs";
        for (int k = 0; k < 4; ++k) { Cm2[k] = Cm[sym[best][k]]; dp2[k] = dp[sym[best][k]]; ly[k] = lx[sym[best][k]]; }
        DTen y(ctx, dp2[0], dp2[1], dp2[2], dp2[3]);
        tei_transform_dev(ctx, gphys, Cm2, dp2, y.p(), false, &tws);
        permute_axpby(ctx, 1.0, y, ly, 0.0, out, lx);
    }

    // A class every rank needs in full: with a streamed AO tensor each rank transforms only its
    // share of the sigma planes (the transform is linear) and the partial tensors are summed.
    void replicated_klass(GaoSource& gphys, DTen& out, const char* slots) {
        const bool split = !gphys.resident() && ctx->nranks > 1;
        if (split) {
            const int64_t per = round_up((gphys.np + ctx->nranks - 1) / ctx->nranks, 2);
            gphys.sig_lo = std::min<int64_t>(gphys.np, per * ctx->rank);
            gphys.sig_hi = std::min<int64_t>(gphys.np, per * (ctx->rank + 1));
        }
        klass(gphys, out, slots, false);
        if (split) {
            gphys.sig_lo = 0; gphys.sig_hi = -1;
            all_reduce_sum(ctx, out.p(), (size_t)out.t.size());
        }
    }

    void all_classes(GaoSource& gphys) {
        replicated_klass(gphys, V, "oovv");
        replicated_klass(gphys, J, "ovov");
        replicated_klass(gphys, oooo, "oooo");
        klass(gphys, W4, "vvvv", true);
        if (singles) {
            replicated_klass(gphys, ooov, "ooov");
            klass(gphys, OA, "vvov", true);
            klass(gphys, OB, "vovv", true);
        }
    }

    void build_integrals(GaoSource& gao) {
        {
            Timer t(ctx, "cc.transform");
            // transforming g'[mu,lam,nu,sig] = g[mu,nu,lam,sig] slot by slot yields <pq|rs> directly
            if (gao.resident()) {
                const int64_t np = gao.np;
                DTen gp(ctx, np, np, np, np);
                Ten g(const_cast<double*>(gao.base()), np, np, np, np);
                permute_axpby(ctx, 1.0, g, "mnls", 0.0, gp, "mlns");
                DeviceGao gphys(gp.p(), gao.n, np);
                all_classes(gphys);
            } else {
                gao.phys = true;     // slabs are produced in [mu,lam,nu,sig] order
                all_classes(gao);
                gao.phys = false;
            }
        }
        tws.buf[0].release(); tws.buf[1].release();
        Timer t(ctx, "cc.static");
        const size_t n2 = (size_t)(o * o * v * v);
        Vt.alloc(ctx, o, o, v, v);
        axpby(ctx, n2, 2.0, V.p(), 0.0, Vt.p());
        permute_axpby(ctx, -1.0, V, "ijab", 1.0, Vt, "jiab");          // Vt = 2V - V(ji)
        if (singles) {
            oovo.alloc(ctx, o, o, v, o);
            permute_axpby(ctx, 1.0, ooov, "nmje", 0.0, oovo, "mnej");   // <mn|ej> = <nm|je>
            ooov_t.alloc(ctx, o, o, o, v);
            axpby(ctx, (size_t)(o * o * o * v), 2.0, ooov.p(), 0.0, ooov_t.p());
            permute_axpby(ctx, -1.0, ooov, "mnie", 1.0, ooov_t, "nmie");
        }
    }

    void register_static() {
        ctx->perm_cache = &pcache;
        for (DTen* t : {&V, &J, &oooo, &ooov, &Vt, &oovo, &ooov_t, &OA, &OB})
            if (t->p()) pcache.add(t->t, false);
    }
    ~CC() { ctx->perm_cache = nullptr; pcache.clear(); }

    double energy() { return cc_energy(ctx, V.p(), T2.p(), singles ? T1.p() : nullptr, o, v); }
    void energy_async(double* dev_out) { cc_energy_async(ctx, V.p(), T2.p(), singles ? T1.p() : nullptr, o, v, dev_out); }

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
        const size_t n2 = (size_t)(o * o * v * v);
        const Ten t = T1, T = T2;
        const Ten tS = last_slab(t, b0, vs), T_S = last_slab(T, b0, vs);
        const Ten V_S = last_slab(V, b0, vs), Vt_S = last_slab(Vt, b0, vs), J_S = last_slab(J, b0, vs);
        DTen tau, tauh, Tt;
        Tt.alloc(ctx, o, o, v, v);
        axpby(ctx, n2, 2.0, T.p, 0.0, Tt.p());
        permute_axpby(ctx, -1.0, T, "ijab", 1.0, Tt, "jiab");           // Tt = 2T - T(ji)
        Ten tauv = T, tauhv = T;
        if (singles) {
            tau.alloc(ctx, o, o, v, v); tauh.alloc(ctx, o, o, v, v);
            tau_build(ctx, T.p, t.p, 1.0, tau.p(), o, v);
            tau_build(ctx, T.p, t.p, 0.5, tauh.p(), o, v);
            tauv = tau; tauhv = tauh;
        }
        const Ten tau_S = last_slab(tauv, b0, vs), tauh_S = last_slab(tauhv, b0, vs);
        pcache.add(T, true); pcache.add(Tt, true);
        if (singles) { pcache.add(tau, true); pcache.add(tauh, true); }

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
            contract(ctx, 2.0, OB, "amef", tS, "mf", 1.0, FaeT, "ea");
            contract(ctx, -1.0, OA, "eamf", tS, "mf", 1.0, FaeT, "ea");
            contract(ctx, -1.0, T_S, "mnae", last_slab(ooov_t, b0, vs), "mnie", 0.0, R1, "ia");
            contract(ctx, 2.0, T_S, "imef", OB, "amef", 1.0, R1, "ia");
            contract(ctx, -1.0, T_S, "imef", OA, "eamf", 1.0, R1, "ia");
        } else {
            fill(ctx, R1.p, (size_t)nR1, 0.0);
        }
        delete tr_small;
        { TraceTimer tt(ctx, "cc.comm.allreduce"); all_reduce_sum(ctx, small.p, small.n); }
        axpby(ctx, (size_t)nW, 1.0, oooo.p(), 1.0, Wpp.p);
        DTen Fme, FaeT_t, Fmi_t;
        Ten FaeTt = FaeT, FmiT = Fmi;
        if (singles) {
            Fme.alloc(ctx, o, v);
            contract(ctx, 1.0, Vt, "mnef", t, "nf", 0.0, Fme, "me");
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
        // ---- ring intermediates for the slab, layout [m,e,j,b] ------------------------------------
        TraceTimer* tr_ring = new TraceTimer(ctx, "cc.part.ringW");
        const size_t ns = (size_t)(o * o * v * vs);
        DTen WJ(ctx, o, v, o, vs), WE(ctx, o, v, o, vs);
        permute_axpby(ctx, 1.0, V_S, "mjeb", 0.0, WJ, "mejb");                       // <mb|ej> = <mj|eb>
        axpby(ctx, ns, -1.0, J_S.p, 0.0, WE.p());                                      // -<mb|je> = -(mj|eb)
        contract(ctx, 0.5, Vt, "mnef", T_S, "njfb", 1.0, WJ, "mejb");
        if (singles) {
            DTen Tp2(ctx, o, o, v, v), Tph(ctx, o, o, v, v);
            tau_build(ctx, T.p, t.p, 2.0, Tp2.p(), o, v);                             // T + 2 tt
            axpby(ctx, n2, 0.5, T.p, 0.0, Tph.p());
            tau_build(ctx, Tph.p(), t.p, 1.0, Tph.p(), o, v);                         // T/2 + tt
            contract(ctx, -0.5, V, "mnef", last_slab(Tp2, b0, vs), "jnfb", 1.0, WJ, "mejb");
            contract(ctx, 1.0, V, "nmef", last_slab(Tph, b0, vs), "jnfb", 1.0, WE, "mejb");
            contract(ctx, 1.0, OA, "efmb", t, "jf", 1.0, WJ, "mejb");
            contract(ctx, -1.0, oovo, "mnej", tS, "nb", 1.0, WJ, "mejb");
            contract(ctx, -1.0, OA, "femb", t, "jf", 1.0, WE, "mejb");
            contract(ctx, 1.0, oovo, "nmej", tS, "nb", 1.0, WE, "mejb");
        } else {
            contract(ctx, -0.5, V, "mnef", T_S, "jnfb", 1.0, WJ, "mejb");
            contract(ctx, 0.5, V, "nmef", T_S, "jnfb", 1.0, WE, "mejb");
        }
        delete tr_ring;
        // ---- ladders ---------------------------------------------------------------------------------
        TraceTimer* tr_lad = new TraceTimer(ctx, "cc.part.ladders+H");
        DTen Lpp(ctx, o, o, v, vs), Lhh(ctx, o, o, v, vs), Hfull(ctx, o, o, v, v);
        const Ten H = last_slab(Hfull, b0, vs);
        contract(ctx, 1.0, tauv, "ijef", W4, "efab", 0.0, Lpp, "ijab");
        contract(ctx, 1.0, Wpp, "mnij", tau_S, "mnab", 0.0, Lhh, "ijab");
        // ---- half residual H for the slab (its (ij)(ab) image is added by residual_finish) ----------
        contract(ctx, 1.0, T, "ijae", last_slab(FaeTt, b0, vs), "eb", 0.0, H, "ijab");
        contract(ctx, -1.0, T_S, "imab", FmiT, "mj", 1.0, H, "ijab");
        contract(ctx, 1.0, Tt, "imae", WJ, "mejb", 1.0, H, "ijab");
        contract(ctx, 1.0, T, "imae", WE, "mejb", 1.0, H, "ijab");
        contract(ctx, 1.0, T, "mjae", WE, "meib", 1.0, H, "ijab");    // image of T[mibe] WmBEj[maej]
        if (singles) {
            DTen Yp(ctx, o, o, o, vs);
            contract(ctx, 1.0, tauv, "ijef", OA, "efmb", 0.0, Yp, "ijmb");
            contract(ctx, -1.0, Yp, "ijmb", t, "ma", 1.0, H, "ijab");
            // rank-1 ring corrections  - t[ie] t[ma] <mb|ej>  - t[ie] t[mb] <am|ej>: contract t[ie] into the
            // integral first (o^3 v intermediates) instead of building v^3 o ones
            DTen Q1(ctx, o, o, o, vs), Q2(ctx, o, o, v, o);
            contract(ctx, 1.0, t, "ie", V_S, "mjeb", 0.0, Q1, "imjb");
            contract(ctx, -1.0, Q1, "imjb", t, "ma", 1.0, H, "ijab");
            contract(ctx, 1.0, t, "ie", J, "maje", 0.0, Q2, "imaj");
            contract(ctx, -1.0, Q2, "imaj", tS, "mb", 1.0, H, "ijab");
            contract(ctx, 1.0, t, "ie", OB, "ajeb", 1.0, H, "ijab");     // t . <ab|ej>
            contract(ctx, -1.0, t, "ma", last_slab(ooov, b0, vs), "mjib", 1.0, H, "ijab");
        }
        delete tr_lad;
        { TraceTimer tt(ctx, "cc.comm.gatherH"); all_gather_inplace(ctx, Hfull.p(), ns); }
        const Ten Tn_S = last_slab(T2n, b0, vs);
        residual_finish(ctx, V_S.p, Lpp.p(), Lhh.p(), H.p, Hfull.p(), Tn_S.p, P.eo.p, P.ev.p, o, v, b0, vs);
        { TraceTimer tt(ctx, "cc.comm.gatherT2"); all_gather_inplace(ctx, T2n.p(), ns); }
        pcache.end_sweep();
        std::swap(T2.buf, T2n.buf); std::swap(T2.t, T2n.t);
        if (singles) { std::swap(T1.buf, T1n.buf); std::swap(T1.t, T1n.t); }
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

CCResult cc_dev(jues_ctx* ctx, Problem& P, GaoSource& gao, bool singles, int maxit, int guess_mode,
                double* T1_out, double* T2_out, jues_b200_amp_cb cb, void* cb_user) {
    JUES_REQUIRE(maxit >= 0, "maxit must be non-negative");
    CC cc(ctx, P, singles);
    cc.build_integrals(gao);
    cc.register_static();
    cc.guess(guess_mode);
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
            cc.iterate();
            cc.energy_async(e_dev.p + it);
            t.stop();
            // FP64 flops this rank's GEMM launches executed in the sweep (reported as a pseudo-phase)
            ctx->timings.emplace_back("cc.iteration.gflop", (float)((ctx->stats.gemm_flops - f0) * 1e-9));
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
    cc.download(singles ? T1_out : nullptr, T2_out);
    return res;
}

}  // namespace jues
