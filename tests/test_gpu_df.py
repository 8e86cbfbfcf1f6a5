"""Density-fitted entry points on the GPU (through the C ABI) against the literal restatement of
DF-RMP2.jl / DF-RCCD.jl (oracle/jues_oracle_df.py): energy of every sweep and final amplitudes."""
import numpy as np
import pytest

import jues.jl_b200 as jb
from oracle import jues_oracle_df as odf
from test_oracle_df import df_inputs

pytestmark = pytest.mark.gpu

E_TOL, T_TOL = 1e-10, 1e-9      # BASELINE.json north_star


def make_wfn(nbf, nocc, naux, seed):
    pqP, Jpqh, C, eps = df_inputs(nbf, nocc, naux, seed)
    w = jb.Wfn(nocc, nbf - nocc, eps, np.asfortranarray(C[:, :nocc]), np.asfortranarray(C[:, nocc:]), None,
               Ca=np.asfortranarray(C), df=(pqP, Jpqh))
    return w, pqP, Jpqh, C, eps


@pytest.mark.parametrize("nbf,nocc,naux", [(7, 3, 11), (12, 4, 30), (24, 5, 61), (40, 8, 112)])
def test_df_rmp2(ctx, nbf, nocc, naux):
    w, pqP, Jpqh, C, eps = make_wfn(nbf, nocc, naux, 100 + nbf)
    ref = odf.do_df_rmp2(pqP, Jpqh, C, nocc, nbf - nocc, eps)
    got = jb.do_df_rmp2(w, ctx=ctx)
    assert abs(got - ref) <= E_TOL


@pytest.mark.parametrize("nbf,nocc,naux,maxit", [(7, 3, 11, 40), (12, 4, 30, 40), (24, 5, 61, 12), (31, 6, 77, 5)])
def test_df_rccd_every_sweep(ctx, nbf, nocc, naux, maxit):
    w, pqP, Jpqh, C, eps = make_wfn(nbf, nocc, naux, 200 + nbf)
    ref_e = []
    e_ref, T_ref = odf.do_df_rccd(pqP, Jpqh, w.Cao, w.Cav, eps, maxit=maxit, return_T2=True,
                                  callback=lambda it, e, T: ref_e.append(e))
    hist = []
    e, T2 = jb.DFRCCD.do_df_rccd(w, ctx=ctx, maxit=maxit, return_T2=True, _e_hist=hist)
    assert len(hist) == maxit + 1 == len(ref_e)
    assert np.abs(np.asarray(hist) - np.asarray(ref_e)).max() <= E_TOL
    assert abs(e - e_ref) <= E_TOL
    assert np.abs(T2 - T_ref).max() <= T_TOL


def test_df_rccd_maxit_is_honoured_and_errors(ctx):
    w, pqP, Jpqh, C, eps = make_wfn(9, 3, 14, 7)
    e0 = jb.DFRCCD.do_df_rccd(w, ctx=ctx, maxit=0)
    e2 = jb.DFRCCD.do_df_rccd(w, ctx=ctx, maxit=2)
    assert abs(e0 - odf.do_df_rccd(pqP, Jpqh, w.Cao, w.Cav, eps, maxit=0)) <= E_TOL
    assert abs(e2 - odf.do_df_rccd(pqP, Jpqh, w.Cao, w.Cav, eps, maxit=2)) <= E_TOL
    bare = jb.Wfn(3, 6, eps, w.Cao, w.Cav, None, Ca=w.Ca)
    with pytest.raises(jb.JuesError):
        jb.do_df_rmp2(bare, ctx=ctx)
    bad = jb.Wfn(3, 6, eps, w.Cao, w.Cav, None, Ca=w.Ca, df=(pqP[:, :, :5], Jpqh))
    with pytest.raises(jb.JuesError):
        jb.DFRCCD.do_df_rccd(bad, ctx=ctx)
