"""
CPU oracle, part 2  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE (same rules as jues_oracle.py:
only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import it).

numpy restatement of the "next" rows of the hot-path scope table:

* ``get_fock``                 /root/reference/src/Backend/IntegralTransformation.jl:119-141
* ``AutoRCCSD.update_energy``  /root/reference/src/CoupledCluster/AutoRCCSD.jl:40-52
* ``AutoRCCSD.update_amp``     /root/reference/src/CoupledCluster/AutoRCCSD.jl:67-178
* ``AutoRCCSD.do_rccsd``       /root/reference/src/CoupledCluster/AutoRCCSD.jl:193-301
  (options: CoupledCluster.jl:36-43 ``cc_max_iter=50, cc_max_rms=1e-10, cc_e_conv=1e-10, do_pT, fcn``)
* ``compute_pT``               /root/reference/src/CoupledCluster/PerturbativeTriples.jl:35-138
* ``mRCCD.do_rccd`` (DIIS)     /root/reference/src/CoupledCluster/mRCCD.jl:37-120, 143-207

``update_amp`` is one 87-line ``@tensoropt`` block; it is restated here as a term table
(coefficient, factors with their index letters) evaluated with ``numpy.einsum`` -- the same
pairwise TTGT evaluation TensorOperations performs -- in the reference's own term order.

Pinning status: the reference holds no golden vectors for these functions.  The restatement is
pinned (tests/test_oracle_auto.py) by (i) the identity AutoRCCSD == RCCSD.jl per sweep when the
off-diagonal Fock blocks vanish (two independent derivations inside the reference), (ii) the
H2O/STO-3G known answer for RCCSD (test/TestCoupledCluster.jl:44-45) reached through
``do_auto_rccsd`` on the offline fixture, (iii) an independent all-index closed-shell (T) formula.
Per-sweep amplitudes on synthetic inputs: "parity unpinned" by reference tests.
"""
from __future__ import annotations

import math
from typing import Callable, Optional

import numpy as np

from .jues_oracle import Wfn, get_eri, _es, rccd_intermediates, rccd_residual, rccd_energy, form_Dijab

__all__ = ["get_fock", "auto_update_energy", "auto_update_amp", "do_auto_rccsd", "compute_pT",
           "do_mrccd", "CC_DEFAULTS", "AUTO_T1_TERMS", "AUTO_T2_TERMS", "AUTO_P_TERMS"]

# CoupledCluster.jl:36-43
CC_DEFAULTS = dict(cc_max_iter=50, cc_max_rms=1e-10, cc_e_conv=1e-10, diis=False, do_pT=False, fcn=0)


# --------------------------------------------------------------------------------------
# IntegralTransformation.jl:119-141
# --------------------------------------------------------------------------------------
def get_fock(wfn: Wfn, spin: str = "alpha") -> np.ndarray:
    if spin.lower() in ("alpha", "up", "a"):
        C, Co = wfn.Ca, wfn.Cao
    elif spin.lower() in ("beta", "down", "b"):
        C, Co = wfn.Cb, wfn.Cbo
    else:
        raise ValueError(f"Invalid Spin option given to JuES.IntegralTransformation.get_fock: {spin}")
    gao, hao = wfn.ao_eri, wfn.hao
    f = _es("mp,nq,mn->pq", C, C, hao)                                    # :136
    f = f + 2 * _es("mp,nq,lk,sk,mnls->pq", C, C, Co, Co, gao)            # :137
    f = f - _es("mp,nq,lk,sk,mlns->pq", C, C, Co, Co, gao)                # :138
    return f


# --------------------------------------------------------------------------------------
# AutoRCCSD.jl:40-52
# --------------------------------------------------------------------------------------
def auto_update_energy(T1, T2, fov, Voovv) -> float:
    e = 2.0 * float(np.sum(fov * T1))                                     # :43
    B = -1.0 * _es("lc,kd->lckd", T1, T1)                                 # :44
    B = B - T2.transpose(0, 2, 1, 3)                                      # :45  B[l,c,k,d] -= T2[l,k,c,d]
    B = B + 2.0 * T2.transpose(1, 2, 0, 3)                                # :46  B[l,c,k,d] += 2 T2[k,l,c,d]
    e += float(_es("lckd,klcd->", B, Voovv))                              # :47
    e += 2.0 * float(_es("lc,kd,lkcd->", T1, T1, Voovv))                  # :48
    return e


# --------------------------------------------------------------------------------------
# AutoRCCSD.jl:77-166 as a table.  "t" = T1, "T" = T2, "foo/fov/fvv" = off-diagonal Fock blocks,
# the rest are get_eri(...) classes in physicists' notation (AutoRCCSD.jl:239-240).
# --------------------------------------------------------------------------------------
AUTO_T1_TERMS = [                                    # newT1[i,a]            :78-103
    (+1.0, "fov:ia"),
    (-1.0, "foo:ik t:ka"),
    (+1.0, "fvv:ca t:ic"),
    (-1.0, "fov:kc t:ic t:ka"),
    (+2.0, "fov:kc T:ikac"),
    (-1.0, "fov:kc T:kiac"),
    (-1.0, "t:kc ovov:icka"),
    (+2.0, "t:kc oovv:kica"),
    (-1.0, "T:kicd ovvv:kadc"),
    (+2.0, "T:ikcd ovvv:kadc"),
    (-2.0, "T:klac ooov:klic"),
    (+1.0, "T:lkac ooov:klic"),
    (-2.0, "t:kc t:la ooov:lkic"),
    (-1.0, "t:kc t:id ovvv:kadc"),
    (+2.0, "t:kc t:id ovvv:kacd"),
    (+1.0, "t:kc t:la ooov:klic"),
    (-2.0, "t:kc T:ilad oovv:lkcd"),
    (-2.0, "t:kc T:liad oovv:klcd"),
    (+1.0, "t:kc T:liad oovv:lkcd"),
    (-2.0, "t:ic T:lkad oovv:lkcd"),
    (+1.0, "t:ic T:lkad oovv:klcd"),
    (-2.0, "t:la T:ikdc oovv:klcd"),
    (+1.0, "t:la T:ikcd oovv:klcd"),
    (+1.0, "t:kc t:id t:la oovv:lkcd"),
    (-2.0, "t:kc t:id t:la oovv:klcd"),
    (+4.0, "t:kc T:ilad oovv:klcd"),
]
AUTO_T2_TERMS = [                                    # newT2[i,j,a,b]        :105-129
    (+1.0, "oovv:ijab"),
    (+1.0, "t:ic t:jd vvvv:cdab"),
    (+1.0, "T:ijcd vvvv:cdab"),
    (+1.0, "t:ka t:lb oooo:ijkl"),
    (+1.0, "T:klab oooo:ijkl"),
    (-1.0, "t:ic t:jd t:ka ovvv:kbcd"),
    (-1.0, "t:ic t:jd t:kb ovvv:kadc"),
    (+1.0, "t:ic t:ka t:lb ooov:lkjc"),
    (+1.0, "t:jc t:ka t:lb ooov:klic"),
    (+1.0, "T:klac T:ijdb oovv:klcd"),
    (-2.0, "T:ikac T:ljbd oovv:klcd"),
    (-2.0, "T:lkac T:ijdb oovv:klcd"),
    (+1.0, "T:kiac T:ljdb oovv:lkcd"),
    (+1.0, "T:ikac T:ljbd oovv:lkcd"),
    (-2.0, "T:ikac T:jlbd oovv:lkcd"),
    (+1.0, "T:kiac T:ljbd oovv:klcd"),
    (-2.0, "T:kiac T:jlbd oovv:klcd"),
    (+1.0, "T:ijac T:lkbd oovv:klcd"),
    (-2.0, "T:ijac T:klbd oovv:klcd"),
    (+1.0, "T:kjac T:ildb oovv:lkcd"),
    (+4.0, "T:ikac T:jlbd oovv:klcd"),
    (+1.0, "T:ijdc T:lkab oovv:klcd"),
    (+1.0, "t:ic t:jd t:ka t:lb oovv:klcd"),
    (+1.0, "t:ic t:jd T:lkab oovv:lkcd"),
    (+1.0, "t:ka t:lb T:ijdc oovv:lkcd"),
]
AUTO_P_TERMS = [                                     # P_OoVv[i,j,a,b]       :130-165
    (-1.0, "foo:ik T:kjab"),
    (+1.0, "fvv:ca T:ijcb"),
    (-1.0, "t:kb ooov:jika"),
    (+1.0, "t:jc ovvv:icab"),
    (-1.0, "fov:kc t:ic T:kjab"),
    (-1.0, "fov:kc t:ka T:ijcb"),
    (-1.0, "T:kiac oovv:kjcb"),
    (-1.0, "t:ic t:ka oovv:kjcb"),
    (-1.0, "t:ic t:kb ovov:jcka"),
    (+2.0, "T:ikac oovv:kjcb"),
    (-1.0, "T:ikac ovov:jckb"),
    (-1.0, "T:kjac ovov:ickb"),
    (-2.0, "t:lb T:ikac ooov:lkjc"),
    (+1.0, "t:lb T:kiac ooov:lkjc"),
    (-1.0, "t:jc T:ikdb ovvv:kacd"),
    (-1.0, "t:jc T:kiad ovvv:kbdc"),
    (-1.0, "t:jc T:ikad ovvv:kbcd"),
    (+1.0, "t:jc T:lkab ooov:lkic"),
    (+1.0, "t:lb T:ikac ooov:kljc"),
    (-1.0, "t:ka T:ijdc ovvv:kbdc"),
    (+1.0, "t:ka T:ilcb ooov:lkjc"),
    (+2.0, "t:jc T:ikad ovvv:kbdc"),
    (-1.0, "t:kc T:ijad ovvv:kbdc"),
    (+2.0, "t:kc T:ijad ovvv:kbcd"),
    (+1.0, "t:kc T:ilab ooov:kljc"),
    (-2.0, "t:kc T:ilab ooov:lkjc"),
    (+1.0, "T:jkcd T:ilab oovv:klcd"),
    (-2.0, "t:kc t:jd T:ilab oovv:klcd"),
    (+1.0, "t:kc t:jd T:ilab oovv:lkcd"),
    (-2.0, "t:kc t:la T:ijdb oovv:klcd"),
    (+1.0, "t:kc t:la T:ijdb oovv:lkcd"),
    (+1.0, "t:ic t:ka T:ljbd oovv:klcd"),
    (-2.0, "t:ic t:ka T:jlbd oovv:klcd"),
    (+1.0, "t:ic t:ka T:ljdb oovv:lkcd"),
    (+1.0, "t:ic t:lb T:kjad oovv:klcd"),
    (-2.0, "T:ikdc T:ljab oovv:klcd"),
]


def _eval_terms(terms, out_idx, shape, env):
    acc = np.zeros(shape)
    for coef, spec in terms:
        names, idx = zip(*(f.split(":") for f in spec.split()))
        acc += coef * np.einsum(",".join(idx) + "->" + out_idx, *(env[n] for n in names), optimize=True)
    return acc


def auto_update_amp(T1, T2, f, V, d, D):
    """AutoRCCSD.update_amp (:67-178).  f = (fock_OO, fock_OV, fock_VV) with zero diagonals,
    V = (Voooo, Vooov, Voovv, Vovov, Vovvv, Vvvvv); returns newT1, newT2, r1, r2."""
    oooo, ooov, oovv, ovov, ovvv, vvvv = V
    foo, fov, fvv = f
    env = dict(t=T1, T=T2, foo=foo, fov=fov, fvv=fvv, oooo=oooo, ooov=ooov, oovv=oovv, ovov=ovov,
               ovvv=ovvv, vvvv=vvvv)
    newT1 = _eval_terms(AUTO_T1_TERMS, "ia", T1.shape, env)
    newT2 = _eval_terms(AUTO_T2_TERMS, "ijab", T2.shape, env)
    P = _eval_terms(AUTO_P_TERMS, "ijab", T2.shape, env)
    newT2 = newT2 + P + P.transpose(1, 0, 3, 2)                           # :166
    newT1 = newT1 / d                                                     # :170
    newT2 = newT2 / D                                                     # :171
    r1 = math.sqrt(float(np.sum((newT1 - T1) ** 2))) / T1.size            # :174
    r2 = math.sqrt(float(np.sum((newT2 - T2) ** 2))) / T2.size            # :175
    return newT1, newT2, r1, r2


def auto_setup(wfn: Wfn, fcn: int = 0):
    """AutoRCCSD.jl:206-249: slices, Fock blocks, integral classes, resolvents, MP2 guess."""
    nelec = wfn.nalpha + wfn.nbeta
    if nelec % 2 != 0:
        raise ValueError(f"Number of electrons must be even for RHF. Given {nelec}")
    nmo = wfn.nmo
    ndocc = nelec // 2
    o = slice(fcn, ndocc)
    v = slice(ndocc, nmo)
    f = get_fock(wfn, spin="alpha")
    fd = np.diag(f).copy()
    fock_Od, fock_Vd = fd[o], fd[v]
    f = f - np.diag(fd)
    fblocks = (f[o, o], f[o, v], f[v, v])
    V = tuple(get_eri(wfn, s, fcn=fcn) for s in ("OOOO", "OOOV", "OOVV", "OVOV", "OVVV", "VVVV"))
    d = fock_Od[:, None] - fock_Vd[None, :]
    D = (fock_Od[:, None, None, None] + fock_Od[None, :, None, None]
         - fock_Vd[None, None, :, None] - fock_Vd[None, None, None, :])
    return fblocks, V, d, D, fock_Od, fock_Vd


def do_auto_rccsd(wfn: Wfn, callback: Optional[Callable] = None, return_all: bool = False, **kwargs):
    """AutoRCCSD.do_rccsd (:193-301).  Options as CoupledCluster.defaults (unknown kwargs ignored,
    :199-205).  The reference returns nothing useful (its last expression is an @output); here the
    correlation energy is returned -- or, with return_all, a dict with everything it prints."""
    opt = dict(CC_DEFAULTS)
    opt.update({k: v for k, v in kwargs.items() if k in CC_DEFAULTS})
    fcn = int(opt["fcn"])
    f, V, d, D, fo, fv = auto_setup(wfn, fcn)
    T1 = f[1] / d                                                         # :246
    T2 = V[2] / D                                                         # :247
    Ecc = auto_update_energy(T1, T2, f[1], V[2])                          # :250
    e_hist, rms_hist = [Ecc], [1.0]
    if callback is not None:
        callback(0, Ecc, T1, T2)
    dE, rms, ite = 1.0, 1.0, 1
    while abs(dE) > opt["cc_e_conv"] or rms > opt["cc_max_rms"]:          # :270
        if ite > opt["cc_max_iter"]:
            break
        T1, T2, r1, r2 = auto_update_amp(T1, T2, f, V, d, D)
        rms = max(r1, r2)
        oldE = Ecc
        Ecc = auto_update_energy(T1, T2, f[1], V[2])
        dE = Ecc - oldE
        e_hist.append(Ecc)
        rms_hist.append(rms)
        if callback is not None:
            callback(ite, Ecc, T1, T2)
        ite += 1
    converged = abs(dE) < opt["cc_e_conv"] and rms < opt["cc_max_rms"]    # :288
    Ept = None
    if opt["do_pT"]:                                                      # :294-300
        Vvvvo = V[4].transpose(3, 1, 2, 0)
        Vvooo = V[1].transpose(3, 1, 0, 2)
        Vvovo = V[2].transpose(2, 0, 3, 1)
        Ept = compute_pT(T1=T1, T2=T2, Vvvvo=Vvvvo, Vvooo=Vvooo, Vvovo=Vvovo, fo=fo, fv=fv)
    if return_all:
        return dict(ecc=Ecc, ept=Ept, iterations=ite - 1, converged=converged, e_hist=np.array(e_hist),
                    rms_hist=np.array(rms_hist), T1=T1, T2=T2)
    return Ecc


# --------------------------------------------------------------------------------------
# PerturbativeTriples.jl:35-138
# --------------------------------------------------------------------------------------
def compute_pT(*, T1, T2, Vvvvo, Vvooo, Vvovo, fo, fv) -> float:
    """Restricted sums i >= j >= k and a >= b >= c exactly as the reference; the a,b,c loop nest
    (:117-131) is evaluated as a masked array expression (terms summed with fsum per (i,j,k))."""
    o, v = T1.shape
    a_, b_, c_ = np.meshgrid(np.arange(v), np.arange(v), np.arange(v), indexing="ij")
    mask = (a_ >= b_) & (b_ >= c_)
    dab = (a_ == b_).astype(float)
    dbc = (b_ == c_).astype(float)
    fvsum = fv[:, None, None] + fv[None, :, None] + fv[None, None, :]
    Et = 0.0
    for i in range(o):
        for j in range(i + 1):
            dij = float(i == j)
            for k in range(j + 1):
                # :96-101 (letters as in the reference; every line is one pair of contractions)
                W = (_es("bda,cd->abc", Vvvvo[:, :, :, i], T2[k, j]) - _es("cl,lab->abc", Vvooo[:, k, j, :], T2[i])
                     + _es("cda,bd->abc", Vvvvo[:, :, :, i], T2[j, k]) - _es("bl,lac->abc", Vvooo[:, j, k, :], T2[i])
                     + _es("adc,bd->abc", Vvvvo[:, :, :, k], T2[j, i]) - _es("bl,lca->abc", Vvooo[:, j, i, :], T2[k])
                     + _es("bdc,ad->abc", Vvvvo[:, :, :, k], T2[i, j]) - _es("al,lcb->abc", Vvooo[:, i, j, :], T2[k])
                     + _es("cdb,ad->abc", Vvvvo[:, :, :, j], T2[i, k]) - _es("al,lbc->abc", Vvooo[:, i, k, :], T2[j])
                     + _es("adb,cd->abc", Vvvvo[:, :, :, j], T2[k, i]) - _es("cl,lba->abc", Vvooo[:, k, i, :], T2[j]))
                # :103
                Vt = (W + Vvovo[:, j, :, k][None, :, :] * T1[i][:, None, None]
                      + Vvovo[:, i, :, k][:, None, :] * T1[j][None, :, None]
                      + Vvovo[:, i, :, j][:, :, None] * T1[k][None, None, :])
                djk = float(j == k)
                p = lambda A, s: A.transpose(*s)     # A'[a,b,c] = A[perm(a,b,c)]
                Wabc, Vabc = W, Vt
                Wacb, Vacb = p(W, (0, 2, 1)), p(Vt, (0, 2, 1))
                Wbac, Vbac = p(W, (1, 0, 2)), p(Vt, (1, 0, 2))
                Wbca, Vbca = p(W, (2, 0, 1)), p(Vt, (2, 0, 1))     # X[b,c,a] as a function of (a,b,c)
                Wcab, Vcab = p(W, (1, 2, 0)), p(Vt, (1, 2, 0))
                Wcba, Vcba = p(W, (2, 1, 0)), p(Vt, (2, 1, 0))
                Dd = fo[i] + fo[j] + fo[k] - fvsum                                            # :121
                X = Wabc * Vabc + Wacb * Vacb + Wbac * Vbac + Wbca * Vbca + Wcab * Vcab + Wcba * Vcba
                Y = Vabc + Vbca + Vcab
                Z = Vacb + Vbac + Vcba
                E = (Y - 2 * Z) * (Wabc + Wbca + Wcab) + (Z - 2 * Y) * (Wacb + Wbac + Wcba) + 3 * X
                contrib = E * (2 - dij - djk) / (Dd * (1 + dab + dbc))                        # :127
                Et += math.fsum(contrib[mask].tolist())
    return Et


# --------------------------------------------------------------------------------------
# mRCCD.jl:37-120 (driver), :143-207 (cciter with DIIS)
# --------------------------------------------------------------------------------------
def do_mrccd(wfn: Wfn, maxit: int = 40, return_T2: bool = False, callback: Optional[Callable] = None,
             return_all: bool = False):
    """mRCCD.do_rccd: zero initial amplitudes (T2_init!, :239-256), DIIS with at most 6 vectors kept
    in **Float32** (:64-65,171,175), B normalised by max|B| (:194), coefficients from inv(B)*resid in
    Float32 (:198), new T2 = sum_k Float32(c_k) * Float32 amplitudes (:200-202), stop when the
    2-norm of (T2new - T2old) drops below 1e-7 (:106,206).  The residual itself is RCCD.jl's
    (the GEMM chain of mRCCD.jl:265-488 equals it to 1e-16 -- SURVEY.md App. C)."""
    nocc, nvir = wfn.nalpha, wfn.nvira
    oovv = get_eri(wfn, "OOVV")
    vvvv = get_eri(wfn, "VVVV")
    ovvo = get_eri(wfn, "OVVO")
    ovov = get_eri(wfn, "OVOV")
    oooo = get_eri(wfn, "OOOO")
    Dijab = form_Dijab(nocc, nvir, wfn.epsa)
    T2 = np.zeros((nocc, nocc, nvir, nvir))
    vals = [T2.astype(np.float32)]
    errs = [np.zeros(0, np.float32)]
    max_diis = 6
    rms_hist, e_hist = [], []
    nit = 0
    for i in range(maxit):
        Fae, Fmi, Wabef, Wmnij, WmBeJ, WmBEj = rccd_intermediates(T2, oovv, ovov, ovvo, oooo, vvvv)
        Td = rccd_residual(T2, Fae, Fmi, WmBeJ, WmBEj, Wabef, Wmnij, oovv) / Dijab
        err64 = (Td - T2).ravel()
        vals.append(Td.astype(np.float32))
        errs.append(err64.astype(np.float32))
        if i == 0:
            del errs[0]
        if len(vals) > max_diis:
            del vals[0]
            del errs[0]
        n = len(vals) - 1
        B = -np.ones((n + 1, n + 1), np.float32)
        B[-1, -1] = 0
        for n1, e1 in enumerate(errs):
            for n2, e2 in enumerate(errs):
                B[n1, n2] = np.dot(e1, e2)
        B[:n, :n] /= np.abs(B[:n, :n]).max()
        resid = np.zeros(n + 1, np.float32)
        resid[-1] = -1
        ci = np.linalg.inv(B).astype(np.float32) @ resid
        Tn = np.zeros_like(Td)
        for num in range(n):
            Tn += (np.float32(ci[num]) * vals[num + 1]).astype(np.float64)
        T2 = Tn
        r2 = float(np.linalg.norm(err64))
        rms_hist.append(r2)
        e_hist.append(rccd_energy(T2, oovv))
        nit = i + 1
        if callback is not None:
            callback(nit, e_hist[-1], None, T2)
        if r2 < 1e-7:
            break
    e = rccd_energy(T2, oovv)
    if return_all:
        return dict(ecc=e, iterations=nit, rms_hist=np.array(rms_hist), e_hist=np.array(e_hist), T2=T2)
    return (e, T2) if return_T2 else e
