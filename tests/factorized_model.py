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
def rccd_iteration(I, T, D, relaid=False, df_wmbej=False):
    """relaid: the second-half-of-round-2 form of the sweep (cc.cu, `relaid`): the Fmi term enters through its
    (ij)(ab) image and the three ring products are summed in their GEMM-native layouts before they are added
    to H.  df_wmbej: the ring intermediate of DF-RCCD.jl:248-258 (<mn|ef> where RCCD.jl:402 has <nm|ef>)."""
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
    WmBeJ = ovvo + 0.5 * es("mnef,njfb->mbej", V if df_wmbej else Vt, T) - 0.5 * es("mnef,jnfb->mbej", V, T)
    WmBEj = -J.transpose(0, 1, 3, 2) + 0.5 * es("nmef,jnfb->mbej", V, T)
    # ladders
    Lpp = es("ijef,abef->ijab", T, vvvv)
    Lhh = es("mnij,mnab->ijab", Wpp, T)
    # half residual
    if relaid:
        H = es("ijae,be->ijab", T, Fae) - es("mi,mjab->ijab", Fmi, T)      # image of - T[imab] Fmi[mj]
        Ra = es("imae,mbej->iajb", Tt, WmBeJ) + es("imae,mbej->iajb", T, WmBEj)   # [ia|jb], as the GEMM writes it
        Rb = es("mjae,mbei->jaib", T, WmBEj)                                       # [ja|ib]
        H += Ra.transpose(0, 2, 1, 3) + Rb.transpose(2, 0, 1, 3)                   # ring_combine
    else:
        H = es("ijae,be->ijab", T, Fae) - es("imab,mj->ijab", T, Fmi)
        H += es("imae,mbej->ijab", Tt, WmBeJ)
        H += es("imae,mbej->ijab", T, WmBEj)
        H += es("mibe,maej->ijab", T, WmBEj)
    R = V + Lpp + Lhh + H + P(H)
    return R / D


# ------------------------------------------------------------------------------------------
# RCCSD
# ------------------------------------------------------------------------------------------
def rccsd_iteration(I, t, T, Dia, D, fock=None, relaid=False):
    """fock = (foo, fov, fvv): off-diagonal Fock blocks (zero diagonals) of a non-canonical
    reference, index order as AutoRCCSD.jl:78-80,130-131 uses them (foo[i,k], fov[k,c], fvv[c,a]).
    None = canonical orbitals (RCCSD.jl).  relaid: the form of the sweep without output permutation passes
    (cc.cu, `relaid`): Fmi term through its image, ring products combined in their GEMM-native layouts,
    everything contracted with t[m,a] over m summed first, one static operand 2<am|ef> - <ma|ef>."""
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
    if relaid:
        H = es("ijae,be->ijab", T, Fae_t) - es("mi,mjab->ijab", Fmi_t, T)    # image of - T[imab] Fmi[mj]
        Ra = es("imae,mbej->iajb", Tt, WmBeJ) + es("imae,mbej->iajb", T, WmBEj)
        Rb = es("mjae,mbei->jaib", T, WmBEj)
        H += Ra.transpose(0, 2, 1, 3) + Rb.transpose(2, 0, 1, 3)                # ring_combine
        # everything contracted with t[m,a] over m: <mj|ib> + tau[ijef] <ef|mb> + t[ie] <mj|eb>
        Ysum = ooov.transpose(2, 1, 0, 3) + Yp + es("ie,mjeb->ijmb", t, V)
        H -= es("ijmb,ma->ijab", Ysum, t)
        Z2 = es("ie,maje->imaj", t, J)
        H -= es("imaj,mb->ijab", Z2, t)
        H += es("ie,jabe->ijab", t, ovvv)
    else:
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
# unpack_ladder_sa; cc.cu: sa_ladder) -- same index functions as the kernels
# ------------------------------------------------------------------------------------------
def sa_partner(z, t, v):
    """Partner w of z at slot t of z's column block (or -1 for the pad slot)."""
    n_low = z // 2 + 1
    w = (z & 1) + 2 * t if t < n_low else z + 1 + 2 * (t - n_low)
    return w if w < v else -1


def sa_column(a, b, v):
    """Column Q of the unordered pair {a,b} in the output-pair space, hv = v/2+1 slots per z."""
    hv = v // 2 + 1
    hi, lo = max(a, b), min(a, b)
    z, w = (lo, hi) if (hi - lo) & 1 else (hi, lo)
    n_low = z // 2 + 1
    t = (w - (z & 1)) // 2 if w <= z else n_low + (w - z - 1) // 2
    return z * hv + t


def sa_pack_tau(tau):
    o, v = tau.shape[0], tau.shape[2]
    pairs = [(e, f) for e in range(v) for f in range(e + 1)]          # P = e(e+1)/2 + f
    Tp, Tm = np.zeros((o * o, len(pairs))), np.zeros((o * o, len(pairs)))
    for P, (e, f) in enumerate(pairs):
        x, y = tau[:, :, e, f].ravel(order="F"), tau[:, :, f, e].ravel(order="F")
        Tp[:, P] = x if e == f else x + y
        Tm[:, P] = x - y
    return Tp, Tm


def sa_pack_vvvv(W4slab, v, b0):
    """[W+ | W-] for the column block of z in [b0, b0+vs) from the slab W4slab[e,f,w,z-b0] = <ef|wz>."""
    vs = W4slab.shape[3]
    hv = v // 2 + 1
    pairs = [(e, f) for e in range(v) for f in range(e + 1)]
    Wp, Wm = np.zeros((len(pairs), vs * hv)), np.zeros((len(pairs), vs * hv))
    for zl in range(vs):
        z = b0 + zl
        for t in range(hv):
            w = sa_partner(z, t, v)
            if w < 0:
                continue
            for P, (e, f) in enumerate(pairs):
                x, y = W4slab[e, f, w, zl], W4slab[f, e, w, zl]
                Wp[P, zl * hv + t] = 2 * x if e == f else x + y
                Wm[P, zl * hv + t] = x - y if w > z else y - x
    return Wp, Wm


def sa_unpack(Lp, Lm, o, v, b0, vs):
    out = np.zeros((o, o, v, vs))
    for a in range(v):
        for bl in range(vs):
            b = b0 + bl
            Q = sa_column(a, b, v)
            s = 0.5 if a > b else (-0.5 if a < b else 0.0)
            out[:, :, a, bl] = (0.5 * Lp[:, Q] + s * Lm[:, Q]).reshape(o, o, order="F")
    return out


def sa_ladder(tau, W4, nranks=1):
    """sum_ef tau[i,j,e,f] W4[e,f,a,b] (W4[e,f,a,b] = <ef|ab> = W4[f,e,b,a]) as the device does it:
    per rank two (o^2 x np)(np x nq) products over its block of the output pairs, all-gather of the
    blocks, unpack of the rank's slab; returns the full (o,o,v,v) result assembled from the slabs."""
    o, v = tau.shape[0], tau.shape[2]
    assert v % (2 * nranks) == 0
    vs = v // nranks
    Tp, Tm = sa_pack_tau(tau)
    Lp_blocks, Lm_blocks = [], []
    for r in range(nranks):
        Wp, Wm = sa_pack_vvvv(W4[:, :, :, r * vs:(r + 1) * vs], v, r * vs)
        Lp_blocks.append(Tp @ Wp)
        Lm_blocks.append(Tm @ Wm)
    Lp, Lm = np.concatenate(Lp_blocks, axis=1), np.concatenate(Lm_blocks, axis=1)      # all-gather
    return np.concatenate([sa_unpack(Lp, Lm, o, v, r * vs, vs) for r in range(nranks)], axis=3)
