"""The density-fitted oracle (oracle/jues_oracle_df.py, DF-RMP2.jl / DF-RCCD.jl restated) pinned on the
conventional oracle: for an exact factorisation (mu nu|lam sig) = sum_Q b[mu,nu,Q] b[lam,sig,Q] the DF
expressions are the conventional ones."""
import numpy as np
import pytest

import factorized_model as fm
from oracle import jues_oracle as orc
from oracle import jues_oracle_df as odf


def df_inputs(nbf, nocc, naux, seed):
    """pqP, Jpqh as DF.setup_df returns them (DF.jl:30-51), synthetic: pqP symmetric in (p,q), J SPD."""
    rng = np.random.default_rng(seed)
    pqP = rng.standard_normal((nbf, nbf, naux)) * (0.6 / np.sqrt(naux))
    pqP = 0.5 * (pqP + pqP.transpose(1, 0, 2))
    X = rng.standard_normal((naux, naux))
    J = X @ X.T / naux + np.eye(naux)
    w, U = np.linalg.eigh(J)
    Jpqh = (U * w ** -0.5) @ U.T
    C = np.linalg.qr(rng.standard_normal((nbf, nbf)))[0]
    eps = np.concatenate([-2.0 - rng.random(nocc)[::-1].cumsum()[::-1] * 0.1, 1.0 + rng.random(nbf - nocc).cumsum() * 0.1])
    return pqP, Jpqh, C, eps


@pytest.mark.parametrize("nbf,nocc,naux", [(7, 3, 11), (10, 4, 23)])
def test_df_equals_conventional_for_an_exact_factorisation(nbf, nocc, naux):
    """DF-RMP2 == RMP2.  DF-RCCD == RCCD (MP2 guess) term by term EXCEPT the third term of WmBeJ, where
    DF-RCCD.jl:256 contracts bov[m,e]*bov[n,f] = <mn|ef> and RCCD.jl:402 contracts oovv[n,m,e,f] = <nm|ef>:
    the density-fitted reference is a different iteration (its test is commented out, TestCoupledCluster.jl:48).
    The restatement follows the DF file literally; the sweep below is the conventional one with exactly that
    one intermediate replaced."""
    es = lambda *a: np.einsum(*a, optimize=True)       # noqa: E731
    pqP, Jpqh, C, eps = df_inputs(nbf, nocc, naux, 5 + nbf)
    nvir = nbf - nocc
    b = odf.make_b(pqP, Jpqh)
    gao = np.einsum("mnQ,lsQ->mnls", b, b)                  # the AO tensor this factorisation represents
    w = orc.Wfn(nocc, nvir, eps, C[:, :nocc].copy(), C[:, nocc:].copy(), gao)
    assert abs(odf.do_df_rmp2(pqP, Jpqh, C, nocc, nvir, eps) - orc.do_rmp2(w)) < 1e-13
    ints = orc.make_rccd_integrals(gao, w.Cao, w.Cav)
    I6 = fm.unique_integrals(gao, w.Cao, w.Cav)
    oovv, ovov, ovvo, oooo, vvvv = ints
    D = orc.form_Dijab(nocc, nvir, eps)
    T = oovv / D
    dfh = []
    odf.do_df_rccd(pqP, Jpqh, w.Cao, w.Cav, eps, maxit=6, callback=lambda it, e, X: dfh.append((e, X.copy())))
    assert len(dfh) == 7
    for it in range(7):
        e_df, T_df = dfh[it]
        assert abs(orc.rccd_energy(T, oovv) - e_df) < 1e-13
        assert np.abs(T - T_df).max() < 1e-13
        Fae, Fmi, Wabef, Wmnij, WmBeJ, WmBEj = orc.rccd_intermediates(T, *ints)
        WmBeJ_df = ovvo + es("mnef,njfb->mbej", oovv, T - T.transpose(1, 0, 2, 3)) / 2
        assert np.abs(WmBeJ_df - WmBeJ).max() > 1e-6         # the two references do differ here
        T_next = orc.rccd_residual(T, Fae, Fmi, WmBeJ_df, WmBEj, Wabef, Wmnij, oovv) / D
        # ... and the numpy statement of what the device runs for it (re-laid sweep, DF ring intermediate)
        assert np.abs(fm.rccd_iteration(I6, T, D, relaid=True, df_wmbej=True) - T_next).max() < 1e-13
        T = T_next


def test_df_rccd_honours_maxit_and_returns_T2():
    pqP, Jpqh, C, eps = df_inputs(6, 2, 9, 1)
    e3, T3 = odf.do_df_rccd(pqP, Jpqh, C[:, :2].copy(), C[:, 2:].copy(), eps, maxit=3, return_T2=True)
    e0 = odf.do_df_rccd(pqP, Jpqh, C[:, :2].copy(), C[:, 2:].copy(), eps, maxit=0)
    assert T3.shape == (2, 2, 4, 4) and e3 != e0


def test_mirror_setup_df_needs_the_tensors():
    """Host side of the mirror: DF.setup_df hands back Wfn.df, and says so when it is missing (no device call)."""
    import jues.jl_b200 as jb
    pqP, Jpqh, C, eps = df_inputs(6, 2, 9, 3)
    w = jb.Wfn(2, 4, eps, C[:, :2].copy(), C[:, 2:].copy(), None, Ca=C, df=(pqP, Jpqh))
    a, b = jb.DF.setup_df(w, dfbname="cc-pvdz-ri")
    assert a is pqP and b is Jpqh
    with pytest.raises(jb.JuesError):
        jb.DF.setup_df(jb.Wfn(2, 4, eps, C[:, :2].copy(), C[:, 2:].copy(), None, Ca=C))
