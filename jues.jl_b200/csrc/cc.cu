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

#include <algorithm>
#include <cmath>

namespace jues {

void setup_problem(jues_ctx* ctx, Problem& P, int64_t nao, const double* Cao, int64_t nocc,
                   const double* Cav, int64_t nvir, const double* eps) {
    JUES_REQUIRE(nao > 0 && nocc > 0 && nvir > 0, "nao, nocc and nvir must be positive");
    JUES_REQUIRE(Cao && Cav && eps, "null orbital data");
    P.nao = nao; P.nocc = nocc; P.nvir = nvir;
    P.np = round_up(nao, 2); P.o = round_up(nocc, 2); P.v = round_up(nvir, 2);
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
    DTen iajb(ctx, o, v, o, v), ijab(ctx, o, o, v, v);
    {
        Timer t(ctx, "mp2.transform");
        const double* Cm[4] = {P.Co.p, P.Cv.p, P.Co.p, P.Cv.p};
        const int64_t dp[4] = {o, v, o, v};
        tei_transform_dev(ctx, gao, Cm, dp, iajb.p());   // (ia|jb), chemists' order
    }
    Timer t(ctx, "mp2.energy");
    permute_axpby(ctx, 1.0, iajb, "iajb", 0.0, ijab, "ijab");  // <ij|ab> (IntegralTransformation.jl:96-98)
    return mp2_energy(ctx, ijab.p(), P.eo.p, P.ev.p, o, v);
}

// ---------------------------------------------------------------------------------------------
// coupled cluster
// ---------------------------------------------------------------------------------------------
namespace {

struct CC {
    jues_ctx* ctx;
    Problem& P;
    bool singles;
    int64_t o, v;
    // unique integral classes (physicists' order, names as in the reference)
    DTen V, J, ooov, ovvv, oooo, vvvv;
    // static combinations
    DTen Vt, ovvo, oovo, ooov_t, Ot;
    // amplitudes
    DTen T1, T2, T1n, T2n;

    CC(jues_ctx* c, Problem& p, bool s) : ctx(c), P(p), singles(s), o(p.o), v(p.v) {}

    void klass(GaoSource& gphys, DTen& out, const char* slots) {
        const double* Cm[4];
        int64_t dp[4];
        for (int q = 0; q < 4; ++q) {
            Cm[q] = slots[q] == 'o' ? P.Co.p : P.Cv.p;
            dp[q] = slots[q] == 'o' ? o : v;
        }
        out.alloc(ctx, dp[0], dp[1], dp[2], dp[3]);
        tei_transform_dev(ctx, gphys, Cm, dp, out.p());
    }

    void build_integrals(GaoSource& gao) {
        JUES_REQUIRE(gao.resident(), "coupled cluster needs the AO integrals resident on the device");
        const int64_t np = gao.np;
        {
            Timer t(ctx, "cc.transform");
            // g'[mu,lam,nu,sig] = g[mu,nu,lam,sig]: transforming g' slot by slot yields <pq|rs> directly
            DTen gp(ctx, np, np, np, np);
            Ten g(const_cast<double*>(gao.base()), np, np, np, np);
            permute_axpby(ctx, 1.0, g, "mnls", 0.0, gp, "mlns");
            DeviceGao gphys(gp.p(), gao.n, np);
            klass(gphys, V, "oovv");
            klass(gphys, J, "ovov");
            klass(gphys, oooo, "oooo");
            klass(gphys, vvvv, "vvvv");
            if (singles) {
                klass(gphys, ooov, "ooov");
                klass(gphys, ovvv, "ovvv");
            }
        }
        Timer t(ctx, "cc.static");
        const size_t n2 = (size_t)(o * o * v * v);
        Vt.alloc(ctx, o, o, v, v);
        axpby(ctx, n2, 2.0, V.p(), 0.0, Vt.p());
        permute_axpby(ctx, -1.0, V, "ijab", 1.0, Vt, "jiab");          // Vt = 2V - V(ji)
        ovvo.alloc(ctx, o, v, v, o);
        permute_axpby(ctx, 1.0, V, "mjeb", 0.0, ovvo, "mbej");          // <mb|ej> = <mj|eb>
        if (singles) {
            oovo.alloc(ctx, o, o, v, o);
            permute_axpby(ctx, 1.0, ooov, "nmje", 0.0, oovo, "mnej");   // <mn|ej> = <nm|je>
            ooov_t.alloc(ctx, o, o, o, v);
            axpby(ctx, (size_t)(o * o * o * v), 2.0, ooov.p(), 0.0, ooov_t.p());
            permute_axpby(ctx, -1.0, ooov, "mnie", 1.0, ooov_t, "nmie");
            Ot.alloc(ctx, o, v, v, v);
            axpby(ctx, (size_t)(o * v * v * v), -1.0, ovvv.p(), 0.0, Ot.p());
            permute_axpby(ctx, 2.0, ovvv, "mafe", 1.0, Ot, "maef");     // 2 <am|ef> - <ma|ef>
        }
    }

    double energy() { return cc_energy(ctx, V.p(), T2.p(), singles ? T1.p() : nullptr, o, v); }

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
        // ---- one- and two-index intermediates --------------------------------------------------
        DTen Fme(ctx, o, v), Fae(ctx, v, v), Fmi(ctx, o, o), Wpp(ctx, o, o, o, o);
        contract(ctx, -1.0, tauhv, "mnaf", Vt, "mnef", 0.0, Fae, "ae");
        contract(ctx, 1.0, Vt, "mnef", tauhv, "inef", 0.0, Fmi, "mi");
        axpby(ctx, (size_t)(o * o * o * o), 1.0, oooo.p(), 0.0, Wpp.p());
        contract(ctx, 1.0, V, "mnef", tauv, "ijef", 1.0, Wpp, "mnij");   // + 2X
        DTen Fae_t, Fmi_t;
        if (singles) {
            contract(ctx, 1.0, Vt, "mnef", t, "nf", 0.0, Fme, "me");
            contract(ctx, 1.0, Ot, "maef", t, "mf", 1.0, Fae, "ae");
            contract(ctx, 1.0, ooov_t, "mnie", t, "ne", 1.0, Fmi, "mi");
            Fae_t.alloc(ctx, v, v); Fmi_t.alloc(ctx, o, o);
            axpby(ctx, (size_t)(v * v), 1.0, Fae.p(), 0.0, Fae_t.p());
            contract(ctx, -0.5, t, "mb", Fme, "me", 1.0, Fae_t, "be");
            axpby(ctx, (size_t)(o * o), 1.0, Fmi.p(), 0.0, Fmi_t.p());
            contract(ctx, 0.5, Fme, "me", t, "je", 1.0, Fmi_t, "mj");
            contract(ctx, 1.0, ooov, "mnie", t, "je", 1.0, Wpp, "mnij");
            contract(ctx, 1.0, oovo, "mnej", t, "ie", 1.0, Wpp, "mnij");
        }
        const Ten FaeT = singles ? (Ten)Fae_t : (Ten)Fae;
        const Ten FmiT = singles ? (Ten)Fmi_t : (Ten)Fmi;
        // ---- ring intermediates ------------------------------------------------------------------
        DTen WmBeJ(ctx, o, v, v, o), WmBEj(ctx, o, v, v, o);
        axpby(ctx, n2, 1.0, ovvo.p(), 0.0, WmBeJ.p());
        permute_axpby(ctx, -1.0, J, "mbje", 0.0, WmBEj, "mbej");
        contract(ctx, 0.5, Vt, "mnef", T, "njfb", 1.0, WmBeJ, "mbej");
        if (singles) {
            DTen Tp2(ctx, o, o, v, v), Tph(ctx, o, o, v, v);
            tau_build(ctx, T.p, t.p, 2.0, Tp2.p(), o, v);                // T + 2 tt
            axpby(ctx, n2, 0.5, T.p, 0.0, Tph.p());
            tau_build(ctx, Tph.p(), t.p, 1.0, Tph.p(), o, v);            // T/2 + tt
            contract(ctx, -0.5, V, "mnef", Tp2, "jnfb", 1.0, WmBeJ, "mbej");
            contract(ctx, 1.0, V, "nmef", Tph, "jnfb", 1.0, WmBEj, "mbej");
            contract(ctx, 1.0, ovvv, "mbef", t, "jf", 1.0, WmBeJ, "mbej");
            contract(ctx, -1.0, oovo, "mnej", t, "nb", 1.0, WmBeJ, "mbej");
            contract(ctx, -1.0, ovvv, "mbfe", t, "jf", 1.0, WmBEj, "mbej");
            contract(ctx, 1.0, oovo, "nmej", t, "nb", 1.0, WmBEj, "mbej");
        } else {
            contract(ctx, -0.5, V, "mnef", T, "jnfb", 1.0, WmBeJ, "mbej");
            contract(ctx, 0.5, V, "nmef", T, "jnfb", 1.0, WmBEj, "mbej");
        }
        // ---- T1 (RCCSD.jl:248-259) -----------------------------------------------------------------
        if (singles) {
            DTen R1(ctx, o, v);
            contract(ctx, 1.0, t, "ie", Fae, "ae", 0.0, R1, "ia");
            contract(ctx, -1.0, Fmi, "mi", t, "ma", 1.0, R1, "ia");
            contract(ctx, 1.0, Tt, "imae", Fme, "me", 1.0, R1, "ia");
            contract(ctx, 2.0, V, "imae", t, "me", 1.0, R1, "ia");
            contract(ctx, -1.0, J, "maie", t, "me", 1.0, R1, "ia");
            contract(ctx, -1.0, ooov_t, "mnie", T, "mnae", 1.0, R1, "ia");
            contract(ctx, 1.0, T, "imef", Ot, "maef", 1.0, R1, "ia");
            divide_Dia(ctx, R1.p(), T1n.p(), P.eo.p, P.ev.p, o, v);
        }
        // ---- T2: ladders --------------------------------------------------------------------------------
        DTen Lpp(ctx, o, o, v, v), Lhh(ctx, o, o, v, v), H(ctx, o, o, v, v);
        contract(ctx, 1.0, tauv, "ijef", vvvv, "abef", 0.0, Lpp, "ijab");
        contract(ctx, 1.0, Wpp, "mnij", tauv, "mnab", 0.0, Lhh, "ijab");
        // ---- T2: half residual H (its (ij)(ab) image is added by residual_finish) -------------------
        contract(ctx, 1.0, T, "ijae", FaeT, "be", 0.0, H, "ijab");
        contract(ctx, -1.0, T, "imab", FmiT, "mj", 1.0, H, "ijab");
        contract(ctx, 1.0, Tt, "imae", WmBeJ, "mbej", 1.0, H, "ijab");
        contract(ctx, 1.0, T, "imae", WmBEj, "mbej", 1.0, H, "ijab");
        contract(ctx, 1.0, T, "mibe", WmBEj, "maej", 1.0, H, "ijab");
        if (singles) {
            DTen Yp(ctx, o, o, o, v);
            contract(ctx, 1.0, tauv, "ijef", ovvv, "mbef", 0.0, Yp, "ijmb");
            contract(ctx, -1.0, Yp, "ijmb", t, "ma", 1.0, H, "ijab");
            DTen Z1(ctx, v, v, v, o), Z2(ctx, v, v, o, v);
            contract(ctx, 1.0, t, "ma", ovvo, "mbej", 0.0, Z1, "abej");
            contract(ctx, -1.0, t, "ie", Z1, "abej", 1.0, H, "ijab");
            contract(ctx, 1.0, t, "mb", J, "maje", 0.0, Z2, "baje");
            contract(ctx, -1.0, t, "ie", Z2, "baje", 1.0, H, "ijab");
            contract(ctx, 1.0, t, "ie", ovvv, "jabe", 1.0, H, "ijab");
            contract(ctx, -1.0, t, "ma", ooov, "mjib", 1.0, H, "ijab");
        }
        residual_finish(ctx, V.p(), Lpp.p(), Lhh.p(), H.p(), T2n.p(), P.eo.p, P.ev.p, o, v);
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
    res.e_hist[0] = cc.energy();
    report(0, res.e_hist[0]);
    for (int it = 1; it <= maxit; ++it) {
        {
            // one timed "step" = one sweep + the energy the reference evaluates every sweep
            // (RCCSD.jl:104); the energy read-back is the step's device->host result
            Timer t(ctx, "cc.iteration");
            cc.iterate();
            res.e_hist[it] = cc.energy();
        }
        report(it, res.e_hist[it]);
    }
    res.energy = res.e_hist[maxit];
    cc.download(singles ? T1_out : nullptr, T2_out);
    return res;
}

}  // namespace jues
