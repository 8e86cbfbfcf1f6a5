"""
CPU oracle for the density-fitted variants of the path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
Only ``tests/`` may import this module (see oracle/jues_oracle.py for the rule).

A numpy (float64) restatement of

* ``do_df_rmp2``                /root/reference/src/MollerPlesset/DF-RMP2.jl:1-46
* ``DF.make_b`` (make_bμν)      /root/reference/src/Backend/DF.jl:52-58
* ``DFRCCD.do_df_rccd`` and its helpers (make_df_rccd_integrals, T2_init!, ccenergy, form_Fae!, form_Fmi!,
  form_Wmnij!, form_WmBeJ!, form_WmBEj!, form_T2, cciter)
                                /root/reference/src/CoupledCluster/DF-RCCD.jl:11-274

every ``@tensor`` / ``@tensoropt`` line transliterated index for index.  The reference obtains its inputs
``pqP[p,q,P] = (pq|P)`` and ``Jpqh = (P|Q)^(-1/2)`` from psi4 (DF.jl:30-51, ``setup_df``); psi4 is not
available, so they are ARGUMENTS here exactly as ``setup_df`` returns them.  Both functions are dead at HEAD
in the reference (they read the non-existent field ``Wfn.uvsr`` only for ``eltype``); what is restated is
their arithmetic.

Pinning status
--------------
The reference's only constant for this path (test/TestMollerPlesset.jl:39, DF-RMP2 of H2O/STO-3G with the
def2-SVP-RI auxiliary basis = -0.04913505451294127) needs d- and f-type three-centre integrals that the
offline fixture generator (oracle/sto3g_fixture.py: s and p functions) cannot produce: "parity unpinned" by
that constant.  The restatement is pinned instead on the already pinned conventional oracle
(oracle/jues_oracle.py): for an EXACT factorisation (mu nu|lam sig) = sum_Q b[mu,nu,Q] b[lam,sig,Q]
DF-RMP2 must equal RMP2 and every DF-RCCD sweep must equal the RCCD sweep started from the MP2 guess
(tests/test_oracle_df.py, 1e-13).
"""
from __future__ import annotations

from typing import Callable, Optional

import numpy as np

__all__ = ["make_b", "do_df_rmp2", "make_df_rccd_integrals", "df_ccenergy", "df_T2_init", "df_cciter", "do_df_rccd"]


def _es(*args):
    return np.einsum(*args, optimize=True)


def make_b(pqP, Jpqh):
    """b[p,q,Q] = pqP[p,q,P] * Jpqh[P,Q]  (DF-RMP2.jl:13-15, DF.jl:53-56)."""
    return _es("pqP,PQ->pqQ", pqP, Jpqh)


def do_df_rmp2(pqP, Jpqh, C, nocc, nvir, eps) -> float:
    """DF-RMP2.jl:1-46.  C = refWfn.Ca (all MOs, occupied first), eps = refWfn.epsa."""
    Co, Cv = C[:, :nocc], C[:, nocc:nocc + nvir]
    b = make_b(pqP, Jpqh)
    binu = _es("mi,mnQ->inQ", Co, b)                      # :17-19
    bia = _es("na,inQ->iaQ", Cv, binu)                    # :20-22
    dmp2 = 0.0
    for i in range(nocc):
        for j in range(nocc):
            bAB = bia[i] @ bia[j].T                        # :33-35  bAB[a,b] = bi[a,Q] bj[b,Q]
            bBA = bAB.T
            D = eps[i] + eps[j] - eps[nocc:nocc + nvir][:, None] - eps[nocc:nocc + nvir][None, :]
            dmp2 += float(np.sum(bAB * (2 * bAB - bBA) / D))   # :37-43
    return dmp2


def make_df_rccd_integrals(pqP, Jpqh, Cao, Cav):
    """DF-RCCD.jl:77-89."""
    b = make_b(pqP, Jpqh)
    bov = _es("mi,mnQ,na->iaQ", Cao, b, Cav)
    bvo = _es("ma,mnQ,ni->aiQ", Cav, b, Cao)
    boo = _es("mi,mnQ,nj->ijQ", Cao, b, Cao)
    bvv = _es("ma,mnQ,nb->abQ", Cav, b, Cav)
    return bov, bvo, boo, bvv


def _Dijab(nocc, nvir, eps):
    eo, ev = eps[:nocc], eps[nocc:nocc + nvir]
    return (eo[:, None, None, None] + eo[None, :, None, None]
            - ev[None, None, :, None] - ev[None, None, None, :])


def df_ccenergy(T, bov):
    """DF-RCCD.jl:90-110."""
    iajb = _es("iaQ,jbQ->ijab", bov, bov)
    return float(np.sum(iajb * 2 * T) - np.sum(iajb * T.transpose(1, 0, 2, 3)))


def df_T2_init(bov, Dijab):
    """DF-RCCD.jl:111-135: T2[i,j,a,b] = bov[i,a,Q] bov[j,b,Q] / D (the MP2 guess)."""
    return _es("iaQ,jbQ->ijab", bov, bov) / Dijab


def df_cciter(T, bov, bvo, boo, bvv, Dijab):
    """cciter (DF-RCCD.jl:55-75): intermediates from the old T2, then form_T2 (:175-220)."""
    Tt = 2 * T - T.transpose(1, 0, 2, 3)
    # form_WmBeJ! (:248-258)
    WmBeJ = (_es("meQ,bjQ->mbej", bov, bvo)
             + _es("meQ,nfQ,njfb->mbej", bov, bov, Tt) / 2
             - _es("meQ,nfQ,njfb->mbej", bov, bov, T) / 2)
    # form_WmBEj! (:267-274)
    WmBEj = -_es("mjQ,beQ->mbej", boo, bvv) + _es("jnfb,neQ,mfQ->mbej", T, bov, bov) / 2
    # form_Wmnij! (:229-237)
    Wmnij = _es("miQ,njQ->mnij", boo, boo) + _es("ijef,meQ,nfQ->mnij", T, bov, bov) / 2
    # form_Fae! (:143-151), form_Fmi! (:159-167)
    Fae = -1 * _es("meQ,nfQ,mnaf->ae", bov, bov, Tt)
    Fmi = _es("meQ,nfQ,inef->mi", bov, bov, 2 * T - T.transpose(0, 1, 3, 2))
    # form_T2 (:197-218)
    mnef = _es("meQ,nfQ->mnef", bov, bov)
    R = (_es("iaQ,jbQ->ijab", bov, bov)
         + _es("ijae,be->ijab", T, Fae) + _es("jibe,ae->ijab", T, Fae)
         - _es("imab,mj->ijab", T, Fmi) - _es("mjab,mi->ijab", T, Fmi)
         + _es("mnab,mnij->ijab", T, Wmnij)
         + _es("ijef,mnab,mnef->ijab", T, T, mnef) / 2
         + _es("ijef,aeQ,bfQ->ijab", T, bvv, bvv)
         + _es("imae,mbej->ijab", T, WmBeJ) * 2
         - _es("miae,mbej->ijab", T, WmBeJ)
         + _es("imae,mbej->ijab", T, WmBEj)
         + _es("mibe,maej->ijab", T, WmBEj)
         + _es("mjae,mbei->ijab", T, WmBEj)
         + _es("jmbe,maei->ijab", T, WmBeJ) * 2
         - _es("mjbe,maei->ijab", T, WmBeJ)
         + _es("jmbe,maei->ijab", T, WmBEj))
    return R / Dijab


def do_df_rccd(pqP, Jpqh, Cao, Cav, eps, maxit: int = 40, return_T2: bool = False,
               callback: Optional[Callable] = None):
    """DFRCCD.do_df_rccd (DF-RCCD.jl:11-54): `maxit` sweeps (this driver honours its keyword, :28) from the
    MP2 guess; callback(it, energy, T2) with it = 0 for the guess."""
    nocc, nvir = Cao.shape[1], Cav.shape[1]
    bov, bvo, boo, bvv = make_df_rccd_integrals(pqP, Jpqh, Cao, Cav)
    D = _Dijab(nocc, nvir, eps)
    T = df_T2_init(bov, D)
    if callback is not None:
        callback(0, df_ccenergy(T, bov), T)
    for it in range(1, maxit + 1):
        T = df_cciter(T, bov, bvo, boo, bvv, D)
        if callback is not None:
            callback(it, df_ccenergy(T, bov), T)
    e = df_ccenergy(T, bov)
    return (e, T) if return_T2 else e
