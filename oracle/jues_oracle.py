"""
CPU oracle for the JuES hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product path
(``jues.jl_b200``) never imports it and fails loudly when the CUDA library is
missing.

What this is
------------
A numpy (float64) restatement of the reference's algorithm for

* ``tei_transform``                    /root/reference/src/Backend/Transformation.jl:39-93
* ``get_eri``                          /root/reference/src/Backend/IntegralTransformation.jl:38-101
* ``do_rmp2``                          /root/reference/src/MollerPlesset/RMP2.jl:11-45
* ``form_Dijab``                       /root/reference/src/CoupledCluster/Denominators.jl:14-33
* ``RCCD.do_rccd`` and helpers         /root/reference/src/CoupledCluster/RCCD.jl:33-467
* ``RCCSD.do_rccsd`` and helpers       /root/reference/src/CoupledCluster/RCCSD.jl:33-289

The reference is pure Julia and delegates its arithmetic to TensorOperations.jl
v2.2.0 (Manifest.toml:322-326) -> Strided 0.3.5 -> OpenBLAS dgemm, none of which is
under /root/reference, and there is no ``julia`` binary in this image, so the
reference cannot be executed here.  ``@tensor`` / ``@tensoropt`` expressions are
Einstein summations evaluated pairwise as permute+reshape+dgemm (TTGT);
``numpy.einsum(..., optimize=True)`` is the same strategy, so every contraction
below is the reference's expression transliterated index-for-index.

Pinning status
--------------
The reference holds no golden vectors for this path.  The
oracle is pinned against the reference's own known-answer constants for
H2O/STO-3G (test/TestMollerPlesset.jl:35, test/TestCoupledCluster.jl:41-45) through
the offline integral fixture in ``oracle/sto3g_fixture.py`` (agreement ~1e-8 Eh, the
accuracy of an independent SCF/integral code -- see tests/test_oracle_known_answers.py)
and by internal cross-checks (two independent factorisations of the same equations).
Amplitude-level / per-iteration parity on synthetic inputs is "parity unpinned" by
reference tests: it is defined by this restatement.

Array conventions: numpy arrays indexed exactly like the Julia arrays (0-based);
memory order is irrelevant to the oracle.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, Optional

import numpy as np

__all__ = [
    "Wfn", "tei_transform", "get_eri", "do_rmp2", "form_Dijab", "form_Dia",
    "make_rccd_integrals", "do_rccd", "make_rccsd_integrals", "do_rccsd",
    "rccd_iteration", "rccsd_iteration", "rccd_energy", "rccsd_energy",
]


def _es(*args):
    return np.einsum(*args, optimize=True)


# --------------------------------------------------------------------------------------
# Wfn: the input record (Wavefunction.jl:67-88).  Only the fields the path reads.
# --------------------------------------------------------------------------------------
@dataclass
class Wfn:
    nalpha: int
    nvira: int
    epsa: np.ndarray          # (nmo,)
    Cao: np.ndarray           # (nbf, nalpha)
    Cav: np.ndarray           # (nbf, nvira)
    ao_eri: np.ndarray        # (nbf,)*4, chemists' (mu nu|lam sig)
    # beta fields mirror alpha for a restricted reference (Wavefunction.jl:106-113)
    nbeta: int = field(default=-1)
    nvirb: int = field(default=-1)
    # read only by get_fock / AutoRCCSD (Wavefunction.jl:67,77,83; IntegralTransformation.jl:119-141)
    hao: Optional[np.ndarray] = None      # (nbf, nbf) core Hamiltonian
    Ca: Optional[np.ndarray] = None       # (nbf, nmo) = [Cao Cav] (Wavefunction.jl:113-118)
    energy: float = 0.0                   # reference (SCF) energy, only printed

    def __post_init__(self):
        if self.nbeta < 0:
            self.nbeta = self.nalpha
        if self.nvirb < 0:
            self.nvirb = self.nvira
        if self.Ca is None:
            self.Ca = np.concatenate([np.asarray(self.Cao), np.asarray(self.Cav)], axis=1)

    @property
    def Cbo(self):
        return self.Cao

    @property
    def Cbv(self):
        return self.Cav

    @property
    def Cb(self):
        return self.Ca

    @property
    def nmo(self):
        return self.nalpha + self.nvira


# --------------------------------------------------------------------------------------
# Transformation.jl:39-93 -- four sequential quarter transforms, sigma first.
# --------------------------------------------------------------------------------------
def tei_transform(gao, C1, C2=None, C3=None, C4=None, name="default"):
    """out[i,a,j,b] = sum C1[mu,i] C2[nu,a] C3[lam,j] C4[sig,b] gao[mu,nu,lam,sig].

    One-C form (Transformation.jl:15-20) when C2..C4 are omitted.
    """
    if C2 is None:
        C2 = C3 = C4 = C1
    # quarter 1 (Transformation.jl:71-73)   (mu,nu,lam,sig) -> (mu,nu,lam,b)
    temp = _es("sb,mnls->mnlb", C4, gao)
    # quarter 2 (:77-79)                    (mu,nu,lam,b)   -> (mu,nu,j,b)
    temp2 = _es("lj,mnlb->mnjb", C3, temp)
    # quarter 3 (:83-85)                    (mu,nu,j,b)     -> (mu,a,j,b)
    temp = _es("na,mnjb->majb", C2, temp2)
    # quarter 4 (:89-91)                    (mu,a,j,b)      -> (i,a,j,b)
    temp2 = _es("mi,majb->iajb", C1, temp)
    return temp2


# --------------------------------------------------------------------------------------
# IntegralTransformation.jl:38-101
# --------------------------------------------------------------------------------------
def get_eri(wfn: Wfn, eri_string: str, notation: str = "phys", fcn: int = 0):
    if len(eri_string.encode()) != 4:
        raise ValueError(
            f"Invalid string given to JuES.IntegralTransformation.get_eri: {eri_string}")
    if notation == "phys":
        eri_string = "".join(eri_string[k] for k in (0, 2, 1, 3))      # :46-48
    C = []
    for s in eri_string:                                                # :52-68
        if s == "o":
            C.append(wfn.Cbo[:, fcn:])
        elif s == "O":
            C.append(wfn.Cao[:, fcn:])
        elif s == "v":
            C.append(wfn.Cbv)
        elif s == "V":
            C.append(wfn.Cav)
    C1, C2, C3, C4 = C                                                  # :70 (errors if != 4)
    V = _es("sb,lj,na,mi,mnls->iajb", C4, C3, C2, C1, wfn.ao_eri)       # :94
    if notation == "phys":
        V = V.transpose(0, 2, 1, 3)                                     # :96-98
    return np.ascontiguousarray(V)


# --------------------------------------------------------------------------------------
# RMP2.jl:11-45
# --------------------------------------------------------------------------------------
def do_rmp2(wfn: Wfn, strict_order: bool = False, **kwargs) -> float:
    """kwargs are accepted and ignored, as in RMP2.jl:11."""
    nocc = wfn.nalpha
    nvir = wfn.nvira
    eps = wfn.epsa
    moeri = get_eri(wfn, "OOVV")                                         # :24
    if strict_order:
        # literal loop nest b,a,j,i with sequential accumulation (:27-41)
        dmp2 = 0.0
        for b in range(nvir):
            for a in range(nvir):
                for j in range(nocc):
                    for i in range(nocc):
                        dmp2 += (moeri[i, j, a, b]
                                 * (2 * moeri[i, j, a, b] - moeri[i, j, b, a])
                                 ) / (eps[i] + eps[j] - eps[nocc + a] - eps[nocc + b])
        return dmp2
    eo = eps[:nocc]
    ev = eps[nocc:nocc + nvir]
    D = eo[:, None, None, None] + eo[None, :, None, None] - ev[None, None, :, None] - ev[None, None, None, :]
    terms = moeri * (2 * moeri - moeri.transpose(0, 1, 3, 2)) / D
    return math.fsum(terms.ravel().tolist()) if terms.size <= 4_000_000 else float(terms.sum())


# --------------------------------------------------------------------------------------
# Denominators.jl:14-33, RCCSD.jl:175-186
# --------------------------------------------------------------------------------------
def form_Dijab(nocc, nvir, eps):
    eo = eps[:nocc]
    ev = eps[nocc:nocc + nvir]
    return (eo[:, None, None, None] + eo[None, :, None, None]
            - ev[None, None, :, None] - ev[None, None, None, :])


def form_Dia(nocc, nvir, eps):
    return eps[:nocc, None] - eps[None, nocc:nocc + nvir]


def _phys(gao, C1, C2, C3, C4):
    """permutedims(tei_transform(...), [1,3,2,4])  (RCCD.jl:91-95, RCCSD.jl:118-132)."""
    return np.ascontiguousarray(tei_transform(gao, C1, C2, C3, C4).transpose(0, 2, 1, 3))


# ======================================================================================
# RCCD  (RCCD.jl)
# ======================================================================================
def make_rccd_integrals(gao, Cao, Cav):
    """RCCD.jl:85-97.  Note `ovov` is built from (Cao,Cao,Cav,Cav) (:89)."""
    oovv = _phys(gao, Cao, Cav, Cao, Cav)
    vvvv = _phys(gao, Cav, Cav, Cav, Cav)
    ovvo = _phys(gao, Cao, Cav, Cav, Cao)
    ovov = _phys(gao, Cao, Cao, Cav, Cav)
    oooo = _phys(gao, Cao, Cao, Cao, Cao)
    return oovv, ovov, ovvo, oooo, vvvv


def rccd_energy(T2, oovv):
    """RCCD.jl:99-117: sum ijab (2 T[i,j,a,b] - T[j,i,a,b])."""
    return float(np.sum(oovv * (2 * T2 - T2.transpose(1, 0, 2, 3))))


def rccd_intermediates(T2, oovv, ovov, ovvo, oooo, vvvv):
    Wabef = _es("mnab,mnef->abef", T2, oovv) / 2 + vvvv                      # :358-360
    WmBeJ = (ovvo
             + _es("mnef,njfb->mbej", oovv, 2 * T2 - T2.transpose(1, 0, 2, 3)) / 2
             - _es("nmef,njfb->mbej", oovv, T2) / 2)                        # :399-402
    WmBEj = -ovov.transpose(0, 1, 3, 2) + _es("jnfb,nmef->mbej", T2, oovv) / 2.0   # :444-446
    Wmnij = oooo + _es("ijef,mnef->mnij", T2, oovv) / 2                      # :320-322
    Fae = -1 * _es("mnef,mnaf->ae", oovv, 2 * T2 - T2.transpose(1, 0, 2, 3))  # :178-180
    Fmi = _es("mnef,inef->mi", oovv, 2 * T2 - T2.transpose(0, 1, 3, 2))      # :215-217
    return Fae, Fmi, Wabef, Wmnij, WmBeJ, WmBEj


def rccd_residual(T, Fae, Fmi, WmBeJ, WmBEj, Wabef, Wmnij, oovv):
    """RCCD.jl:242-256 (15 terms), before the division by Dijab."""
    R = (oovv
         + _es("ijae,be->ijab", T, Fae) + _es("jibe,ae->ijab", T, Fae)
         - _es("imab,mj->ijab", T, Fmi) - _es("mjab,mi->ijab", T, Fmi)
         + _es("mnab,mnij->ijab", T, Wmnij)
         + _es("ijef,abef->ijab", T, Wabef)
         + _es("imae,mbej->ijab", T, WmBeJ) * 2
         - _es("miae,mbej->ijab", T, WmBeJ)
         + _es("imae,mbej->ijab", T, WmBEj)
         + _es("mibe,maej->ijab", T, WmBEj)
         + _es("mjae,mbei->ijab", T, WmBEj)
         + _es("jmbe,maei->ijab", T, WmBeJ) * 2
         - _es("mjbe,maei->ijab", T, WmBeJ)
         + _es("jmbe,maei->ijab", T, WmBEj))
    return R


def rccd_iteration(T2, ints, Dijab):
    """cciter (RCCD.jl:119-143): intermediates from old T2, then form_T2 ./ Dijab (:301)."""
    oovv, ovov, ovvo, oooo, vvvv = ints
    Fae, Fmi, Wabef, Wmnij, WmBeJ, WmBEj = rccd_intermediates(T2, oovv, ovov, ovvo, oooo, vvvv)
    R = rccd_residual(T2, Fae, Fmi, WmBeJ, WmBEj, Wabef, Wmnij, oovv)
    return R / Dijab


def rccd_guess(ovov, Dijab):
    """T2_init! (RCCD.jl:145-160): T2[i,j,a,b] = ovov[i,a,j,b]/D  -- the reference's
    quirk: this is (ij|ab)/D, not the MP2 guess (ia|jb)/D."""
    return ovov.transpose(0, 2, 1, 3) / Dijab


def do_rccd(wfn: Wfn, maxit: int = 40, guess: str = "reference",
            callback: Optional[Callable] = None, return_T2: bool = False, **kwargs):
    """RCCD.do_rccd (RCCD.jl:33-83).  The reference hard-wires maxit=40 and ignores all
    kwargs; `maxit`, `guess` and `callback` exist here so tests can look inside.

    callback(it, energy, T2) is called after every iteration (it = 1..maxit) and once
    with it=0 for the guess.
    """
    nocc, nvir = wfn.nalpha, wfn.nvira
    ints = make_rccd_integrals(wfn.ao_eri, wfn.Cao, wfn.Cav)
    oovv, ovov = ints[0], ints[1]
    Dijab = form_Dijab(nocc, nvir, wfn.epsa)
    if guess == "reference":
        T2 = rccd_guess(ovov, Dijab)
    elif guess == "mp2":
        T2 = oovv / Dijab
    else:
        raise ValueError(guess)
    if callback is not None:
        callback(0, rccd_energy(T2, oovv), T2)
    for it in range(1, maxit + 1):
        T2 = rccd_iteration(T2, ints, Dijab)
        if callback is not None:
            callback(it, rccd_energy(T2, oovv), T2)
    e = rccd_energy(T2, oovv)
    return (e, T2) if return_T2 else e


# ======================================================================================
# RCCSD  (RCCSD.jl)
# ======================================================================================
_RCCSD_NAMES = ("vvvv", "ovvv", "vovv", "vvov", "vvvo", "oovv", "ovvo", "vovo", "ovov",
                "voov", "ooov", "oovo", "ovoo", "vooo", "oooo")


def make_rccsd_integrals(gao, Cao, Cav):
    """RCCSD.jl:117-142: 15 transforms + permutes.  Returns a dict keyed by the
    reference's variable names."""
    o, v = Cao, Cav
    I = {}
    I["vvvv"] = _phys(gao, v, v, v, v)
    I["ovvv"] = _phys(gao, o, v, v, v)
    I["vovv"] = _phys(gao, v, v, o, v)
    I["vvov"] = _phys(gao, v, o, v, v)
    I["vvvo"] = _phys(gao, v, v, v, o)
    I["oovv"] = _phys(gao, o, v, o, v)
    I["ovvo"] = _phys(gao, o, v, v, o)
    I["vovo"] = _phys(gao, v, v, o, o)
    I["ovov"] = _phys(gao, o, o, v, v)
    I["voov"] = _phys(gao, v, o, o, v)
    I["ooov"] = _phys(gao, o, o, o, v)
    I["oovo"] = _phys(gao, o, v, o, o)
    I["ovoo"] = _phys(gao, o, o, v, o)
    I["vooo"] = _phys(gao, v, o, o, o)
    I["oooo"] = _phys(gao, o, o, o, o)
    I["vvov"] = np.ascontiguousarray(I["vvov"].transpose(3, 0, 1, 2))     # :133
    I["vvvo"] = np.ascontiguousarray(I["vvvo"].transpose(2, 0, 1, 3))     # :134
    I["vovv"] = np.ascontiguousarray(I["vovv"].transpose(1, 0, 2, 3))     # :135
    I["vooo"] = np.ascontiguousarray(I["vooo"].transpose(1, 0, 2, 3))     # :136
    return I


def rccsd_energy(oovv, T1, T2):
    """ccenergy (RCCSD.jl:143-149) with tiatia[m,n,a,f] = T1[m,a] T1[n,f] (:81-83)."""
    tt = _es("ma,nf->mnaf", T1, T1)
    X = 2 * T2 + 2 * tt - T2.transpose(1, 0, 2, 3) - tt.transpose(1, 0, 2, 3)
    return float(np.sum(oovv * X))


def rccsd_intermediates(I, T1, T2):
    tt = _es("ma,nf->mnaf", T1, T1)
    oovv, ovvv, vovv = I["oovv"], I["ovvv"], I["vovv"]
    ooov, oovo, oooo = I["ooov"], I["oovo"], I["oooo"]
    # form_Fae! (:187-194)
    Fae = (_es("mf,maef->ae", T1, 2 * vovv - ovvv)
           - _es("mnaf,mnef->ae", T2 + 0.5 * tt, 2 * oovv - oovv.transpose(1, 0, 2, 3)))
    # form_Fmi! (:195-203)
    Fmi = (_es("ne,mnie->mi", T1, 2 * ooov - ooov.transpose(1, 0, 2, 3))
           + _es("inef,mnef->mi", T2 + 0.5 * tt, 2 * oovv - oovv.transpose(0, 1, 3, 2)))
    # form_Fme! (:204-210)
    Fme = _es("nf,mnef->me", T1, 2 * oovv - oovv.transpose(1, 0, 2, 3))
    # form_Wmnij! (:211-219)
    Wmnij = (oooo + _es("je,mnie->mnij", T1, ooov) + _es("ie,mnej->mnij", T1, oovo)
             + 0.5 * _es("ijef,mnef->mnij", T2 + tt, oovv))
    # form_Wabef! (:220-228)
    Wabef = (I["vvvv"] - _es("mb,maef->abef", T1, vovv) - _es("ma,mbef->abef", T1, ovvv)
             + 0.5 * _es("mnab,mnef->abef", T2 + tt, oovv))
    # form_WmBeJ! (:229-238)
    WmBeJ = (I["ovvo"] + _es("jf,mbef->mbej", T1, ovvv) - _es("nb,mnej->mbej", T1, oovo)
             - 0.5 * _es("jnfb,mnef->mbej", T2 + 2 * tt, oovv)
             + 0.5 * _es("njfb,mnef->mbej", T2, 2 * oovv - oovv.transpose(1, 0, 2, 3)))
    # form_WmBEj! (:239-247)
    WmBEj = (-1 * I["vovo"].transpose(1, 0, 2, 3)
             - _es("jf,mbfe->mbej", T1, ovvv) + _es("nb,nmej->mbej", T1, oovo)
             + _es("jnfb,nmef->mbej", 0.5 * T2 + tt, oovv))
    return Fae, Fmi, Fme, Wmnij, Wabef, WmBeJ, WmBEj


def rccsd_T1_residual(I, T1, T2, Fae, Fmi, Fme):
    """form_T1 (:248-259), before ./ Dia."""
    ooov, vovv = I["ooov"], I["vovv"]
    R1 = (_es("ie,ae->ia", T1, Fae) - _es("ma,mi->ia", T1, Fmi)
          + _es("me,imae->ia", Fme, 2 * T2 - T2.transpose(1, 0, 2, 3))
          + _es("me,amie->ia", T1, 2 * I["voov"] - I["ovov"].transpose(1, 0, 2, 3))
          - _es("mnae,mnie->ia", T2, 2 * ooov - ooov.transpose(1, 0, 2, 3))
          + _es("imef,maef->ia", T2, 2 * vovv - vovv.transpose(0, 1, 3, 2)))
    return R1


def rccsd_T2_residual(I, T1, T2, Fae, Fmi, Fme, Wabef, Wmnij, WmBeJ, WmBEj):
    """form_T2 (:260-289), before ./ Dijab.  Term order follows the reference."""
    tt = _es("ma,nf->mnaf", T1, T1)
    ijab, mbej, amej = I["oovv"], I["ovvo"], I["vovo"]
    abej, abie, mbij, amij = I["vvvo"], I["vvov"], I["ovoo"], I["vooo"]
    Fae_t = Fae - 0.5 * _es("mb,me->be", T1, Fme)
    Fmi_t = Fmi + 0.5 * _es("je,me->mj", T1, Fme)
    R2 = (ijab
          + _es("ijae,be->ijab", T2, Fae_t)
          + _es("ijeb,ae->ijab", T2, Fae_t)
          - _es("imab,mj->ijab", T2, Fmi_t)
          - _es("mjab,mi->ijab", T2, Fmi_t)
          + _es("mnab,mnij->ijab", T2 + tt, Wmnij)
          + _es("ijef,abef->ijab", T2 + tt, Wabef)
          + (_es("imae,mbej->ijab", T2 - T2.transpose(1, 0, 2, 3), WmBeJ)
             - _es("imea,mbej->ijab", tt, mbej))
          + _es("imae,mbej->ijab", T2, WmBeJ + WmBEj)
          + (_es("mibe,maej->ijab", T2, WmBEj)
             - _es("imeb,amej->ijab", tt, amej))
          + (_es("mjae,mbei->ijab", T2, WmBEj)
             - _es("jmea,bmei->ijab", tt, amej))
          + (_es("jmbe,maei->ijab", T2 - T2.transpose(1, 0, 2, 3), WmBeJ)
             - _es("jmeb,maei->ijab", tt, mbej))
          + _es("jmbe,maei->ijab", T2, WmBeJ + WmBEj)
          + _es("ie,eabj->ijab", T1, abej)
          + _es("je,eabi->ijab", T1, abie)
          - _es("ma,mbij->ijab", T1, mbij)
          - _es("mb,maij->ijab", T1, amij))
    return R2


def rccsd_iteration(I, T1, T2, Dia, Dijab):
    """cciter (RCCSD.jl:150-173): both new amplitudes come from the OLD (T1,T2)."""
    Fae, Fmi, Fme, Wmnij, Wabef, WmBeJ, WmBEj = rccsd_intermediates(I, T1, T2)
    R1 = rccsd_T1_residual(I, T1, T2, Fae, Fmi, Fme)
    R2 = rccsd_T2_residual(I, T1, T2, Fae, Fmi, Fme, Wabef, Wmnij, WmBeJ, WmBEj)
    return R1 / Dia, R2 / Dijab


def do_rccsd(wfn: Wfn, maxit: int = 40, callback: Optional[Callable] = None,
             return_T: bool = False, **kwargs):
    """RCCSD.do_rccsd (RCCSD.jl:33-116): MP2 guess (:80), T1=0 (:70), `maxit` Jacobi
    sweeps (reference: 40, hard-wired, kwargs ignored), energy each iteration (:104).

    callback(it, energy, T1, T2): it=0 is the guess (the "@MP2" line, :84).
    """
    nocc, nvir = wfn.nalpha, wfn.nvira
    I = make_rccsd_integrals(wfn.ao_eri, wfn.Cao, wfn.Cav)
    Dia = form_Dia(nocc, nvir, wfn.epsa)
    Dijab = form_Dijab(nocc, nvir, wfn.epsa)
    T1 = np.zeros((nocc, nvir))
    T2 = I["oovv"] / Dijab
    if callback is not None:
        callback(0, rccsd_energy(I["oovv"], T1, T2), T1, T2)
    for it in range(1, maxit + 1):
        T1, T2 = rccsd_iteration(I, T1, T2, Dia, Dijab)
        if callback is not None:
            callback(it, rccsd_energy(I["oovv"], T1, T2), T1, T2)
    e = rccsd_energy(I["oovv"], T1, T2)
    return (e, T1, T2) if return_T else e
