"""
numpy model of the FACTORISED RCCD / RCCSD iteration that the CUDA library executes
(jues.jl_b200/csrc/cc_*.cu).  Test infrastructure only: it exists so that the algebra of
the GPU algorithm (no materialised Wabef, six unique integral classes, symmetrised half
residual  R = oovv + Lpp + Lhh + (1 + P(ij)(ab)) H) can be checked against the literal oracle
on the CPU, term by term, before and independently of any kernel.

Every statement is one pairwise contraction (= one GEMM on the device) or an element-wise /
permutation operation, in the same order and with the same index strings as the C++.
"""
import numpy as np


def es(s, a, b):
    return np.einsum(s, a, b, optimize=True)


def unique_integrals(gao, Cao, Cav):
    """The six symmetry-unique MO classes, phys-ordered like the reference's arrays:
    V=oovv[i,j,a,b]=(ia|jb), J=ovov[m,b,j,e]=(mj|be), ooov[m,n,i,e]=(mi|ne),
    ovvv[m,a,e,f]=(me|af), oooo[m,n,i,j]=(mi|nj), vvvv[a,b,e,f]=(ae|bf)."""
    def phys(C1, C2, C3, C4):
        t = np.einsum("mi,na,lj,sb,mnls->iajb", C1, C2, C3, C4, gao, optimize=True)
        return np.ascontiguousarray(t.transpose(0, 2, 1, 3))
    o, v = Cao, Cav
    return dict(V=phys(o, v, o, v), J=phys(o, o, v, v), ooov=phys(o, o, o, v),
                ovvv=phys(o, v, v, v), oooo=phys(o, o, o, o), vvvv=phys(v, v, v, v))


def P(H):
    """(ij)(ab) image: P(H)[i,j,a,b] = H[j,i,b,a]."""
    return H.transpose(1, 0, 3, 2)


# ------------------------------------------------------------------------------------------
# RCCD
# ------------------------------------------------------------------------------------------
def rccd_iteration(I, T, D):
    V, J, oooo, vvvv = I["V"], I["J"], I["oooo"], I["vvvv"]
    Vt = 2 * V - V.transpose(1, 0, 2, 3)                      # static
    ovvo = V.transpose(0, 3, 2, 1)                             # ovvo[m,b,e,j] = V[m,j,e,b]
    # small intermediates
    Tt = 2 * T - T.transpose(1, 0, 2, 3)
    Fae = -es("mnef,mnaf->ae", V, Tt)
    Fmi = es("mnef,inef->mi", Vt, T)
    X = 0.5 * es("mnef,ijef->mnij", V, T)
    Wpp = oooo + 2 * X                                          # Wmnij + X
    # ring intermediates
    WmBeJ = ovvo + 0.5 * es("mnef,njfb->mbej", Vt, T) - 0.5 * es("mnef,jnfb->mbej", V, T)
    WmBEj = -J.transpose(0, 1, 3, 2) + 0.5 * es("nmef,jnfb->mbej", V, T)
    # ladders
    Lpp = es("ijef,abef->ijab", T, vvvv)
    Lhh = es("mnij,mnab->ijab", Wpp, T)
    # half residual
    H = es("ijae,be->ijab", T, Fae) - es("imab,mj->ijab", T, Fmi)
    H += es("imae,mbej->ijab", Tt, WmBeJ)
    H += es("imae,mbej->ijab", T, WmBEj)
    H += es("mibe,maej->ijab", T, WmBEj)
    R = V + Lpp + Lhh + H + P(H)
    return R / D


# ------------------------------------------------------------------------------------------
# RCCSD
# ------------------------------------------------------------------------------------------
def rccsd_iteration(I, t, T, Dia, D, fock=None):
    """fock = (foo, fov, fvv): off-diagonal Fock blocks (zero diagonals) of a non-canonical
    reference, index order as AutoRCCSD.jl:78-80,130-131 uses them (foo[i,k], fov[k,c], fvv[c,a]).
    None = canonical orbitals (RCCSD.jl)."""
    V, J, ooov, ovvv, oooo, vvvv = (I[k] for k in ("V", "J", "ooov", "ovvv", "oooo", "vvvv"))
    # ---- static combinations (built once in the library) ----
    Vt = 2 * V - V.transpose(1, 0, 2, 3)
    ovvo = V.transpose(0, 3, 2, 1)                 # [m,b,e,j] = (me|bj)
    oovo = ooov.transpose(1, 0, 3, 2)              # oovo[m,n,e,j] = ooov[n,m,j,e]
    ooov_t = 2 * ooov - ooov.transpose(1, 0, 2, 3)
    Ot = 2 * ovvv.transpose(0, 1, 3, 2) - ovvv     # 2 vovv - ovvv, [m,a,e,f]
    # ---- amplitude combinations ----
    tt = np.einsum("ma,nf->mnaf", t, t)
    tau = T + tt
    tauh = T + 0.5 * tt
    Tt = 2 * T - T.transpose(1, 0, 2, 3)
    # ---- one- and two-index intermediates ----
    Fme = es("mnef,nf->me", Vt, t)
    Fae = es("maef,mf->ae", Ot, t) - es("mnaf,mnef->ae", tauh, Vt)
    Fmi = es("mnie,ne->mi", ooov_t, t) + es("inef,mnef->mi", tauh, Vt)
    R1f = 0.0
    if fock is not None:
        foo, fov, fvv = fock
        Fae = Fae + fvv.T - 0.5 * es("me,ma->ae", fov, t)       # Fae[a,e] += f[e,a] - 1/2 f[m,e] t[m,a]
        Fmi = Fmi + foo.T + 0.5 * es("me,ie->mi", fov, t)       # Fmi[m,i] += f[i,m] + 1/2 f[m,e] t[i,e]
        Fme = Fme + fov
        R1f = fov
    Fae_t = Fae - 0.5 * es("mb,me->be", t, Fme)
    Fmi_t = Fmi + 0.5 * es("je,me->mj", t, Fme)
    X = 0.5 * es("mnef,ijef->mnij", V, tau)
    Wpp = oooo + es("mnie,je->mnij", ooov, t) + es("mnej,ie->mnij", oovo, t) + 2 * X
    # ---- ring intermediates ----
    WmBeJ = (ovvo + es("mbef,jf->mbej", ovvv, t) - es("mnej,nb->mbej", oovo, t)
             - 0.5 * es("mnef,jnfb->mbej", V, T + 2 * tt) + 0.5 * es("mnef,njfb->mbej", Vt, T))
    WmBEj = (-J.transpose(0, 1, 3, 2) - es("mbfe,jf->mbej", ovvv, t) + es("nmej,nb->mbej", oovo, t)
             + es("nmef,jnfb->mbej", V, 0.5 * T + tt))
    # ---- T1 ----
    R1 = (R1f + es("ie,ae->ia", t, Fae) - es("ma,mi->ia", t, Fmi) + es("imae,me->ia", Tt, Fme)
          + es("imae,me->ia", 2 * V, t) - es("maie,me->ia", J, t)
          - es("mnae,mnie->ia", T, ooov_t) + es("imef,maef->ia", T, Ot))
    # ---- T2: ladders ----
    Lpp = es("ijef,abef->ijab", tau, vvvv)
    Lhh = es("mnij,mnab->ijab", Wpp, tau)
    Yp = es("ijef,mbef->ijmb", tau, ovvv)
    # ---- T2: half residual ----
    H = es("ijae,be->ijab", T, Fae_t) - es("imab,mj->ijab", T, Fmi_t)
    H -= es("ijmb,ma->ijab", Yp, t)
    H += es("imae,mbej->ijab", Tt, WmBeJ)
    H += es("imae,mbej->ijab", T, WmBEj)
    H += es("mibe,maej->ijab", T, WmBEj)
    # rank-1 ring corrections: - t[ie] t[ma] ovvo[mbej] - t[ie] t[mb] vovo[amej]
    Z1 = es("ma,mbej->abej", t, ovvo)
    H -= es("ie,abej->ijab", t, Z1)
    Z2 = es("mb,maje->baje", t, J)                 # vovo[a,m,e,j] = (ae|mj) = J[m,a,j,e]
    H -= es("ie,baje->ijab", t, Z2)
    # t . (vvvo) and t . (ovoo):  vvvo[e,a,b,j] = ovvv[j,a,b,e];  ovoo[m,b,i,j] = ooov[m,j,i,b]
    H += es("ie,jabe->ijab", t, ovvv)
    H -= es("ma,mjib->ijab", t, ooov)
    R2 = V + Lpp + Lhh + H + P(H)
    return R1 / Dia, R2 / D


# ------------------------------------------------------------------------------------------
# symmetric / antisymmetric particle-particle ladder (tensor_ops.cu: pack_vvvv_sa, pack_tau_sa,
# unpack_ladder_sa; cc.cu: sa_ladder)
# ------------------------------------------------------------------------------------------
def sa_ladder(tau, W4):
    """sum_ef tau[i,j,e,f] W4[e,f,a,b] (W4[e,f,a,b] = <ef|ab> = W4[f,e,b,a]) through the packed pair
    space P(e,f) = e(e+1)/2 + f, e >= f, exactly as the device does it: two (o^2 x np)(np x np)
    products, np = v(v+1)/2, instead of one (o^2 x v^2)(v^2 x v^2)."""
    o, v = tau.shape[0], tau.shape[2]
    pairs = [(e, f) for e in range(v) for f in range(e + 1)]
    n = len(pairs)
    Tp, Tm = np.zeros((o * o, n)), np.zeros((o * o, n))
    Wp, Wm = np.zeros((n, n)), np.zeros((n, n))
    for P, (e, f) in enumerate(pairs):
        x, y = tau[:, :, e, f].ravel(order="F"), tau[:, :, f, e].ravel(order="F")
        Tp[:, P] = x if e == f else x + y
        Tm[:, P] = x - y
        for Q, (a, b) in enumerate(pairs):
            xx, yy = W4[e, f, a, b], W4[f, e, a, b]
            Wp[P, Q] = 2 * xx if e == f else xx + yy
            Wm[P, Q] = xx - yy
    Lp, Lm = Tp @ Wp, Tm @ Wm
    out = np.zeros((o, o, v, v))
    for a in range(v):
        for b in range(v):
            hi, lo = max(a, b), min(a, b)
            Q = hi * (hi + 1) // 2 + lo
            s = 0.5 if a > b else (-0.5 if a < b else 0.0)
            out[:, :, a, b] = (0.5 * Lp[:, Q] + s * Lm[:, Q]).reshape(o, o, order="F")
    return out
