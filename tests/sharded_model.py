"""
numpy model of the SHARDED RCCD / RCCSD sweep (one rank's share), the algorithm
jues.jl_b200/csrc/cc.cu executes when the context has nranks > 1 (and, with one slab covering
everything, when nranks == 1).  Test infrastructure only.

Sharding: the output virtual index b of T2new[i,j,a,b] is split into equal
slabs S_r; rank r holds only the last-index slabs of the v^4 and ov^3 integral classes
    W4[e,f,a,b] = <ef|ab>,  OA[e,f,m,b] = <ef|mb> = ovvv[m,b,e,f],  OB[a,j,e,b] = <aj|eb> = (ae|jb)
for b in S_r, builds every ring intermediate only for its slab, and exchanges
    * one sum-all-reduce of the small partial intermediates (Fae, Fmi, Wpp, R1),
    * one all-gather of the packed ladder blocks [L+ | L-] (o^2 v^2 doubles in total),
    * one all-gather of the half residual H' (o^2 v^2 in total),
    * one all-gather of the new T2 slab.
`comm` supplies allreduce(array)->array and allgather_last(array)->array (concatenate along the
last axis over ranks); with comm=None the rank is alone.
"""
import numpy as np


def es(s, a, b):
    return np.einsum(s, a, b, optimize=True)


class SoloComm:
    def allreduce(self, x):
        return x

    def allgather_last(self, x):
        return x


def slab_bounds(v, nranks, rank):
    """Equal slabs of the (padded) virtual extent: v is padded up to a multiple of 2*nranks."""
    vp = -(-v // (2 * nranks)) * (2 * nranks)
    vs = vp // nranks
    return vp, rank * vs, (rank + 1) * vs


def pad_virtuals(I, v, vp):
    """Zero-pad every virtual axis of the unique integral classes from v to vp."""
    def pad(x, axes):
        w = [(0, 0)] * x.ndim
        for a in axes:
            w[a] = (0, vp - v)
        return np.pad(x, w)
    return dict(V=pad(I["V"], (2, 3)), J=pad(I["J"], (1, 3)), ooov=pad(I["ooov"], (3,)),
                ovvv=pad(I["ovvv"], (1, 2, 3)), oooo=I["oooo"], vvvv=pad(I["vvvv"], (0, 1, 2, 3)))


def rank_integrals(I, b0, b1):
    """What one rank keeps: replicated small classes + last-index slabs of the big ones."""
    S = slice(b0, b1)
    R = dict(V=I["V"], J=I["J"], oooo=I["oooo"], ooov=I["ooov"])
    R["W4"] = np.ascontiguousarray(I["vvvv"].transpose(2, 3, 0, 1)[:, :, :, S])        # <ef|ab>
    R["OA"] = np.ascontiguousarray(I["ovvv"].transpose(2, 3, 0, 1)[:, :, :, S])        # [e,f,m,b]
    R["OB"] = np.ascontiguousarray(np.einsum("jabe->ajeb", I["ovvv"])[:, :, :, S])      # (ae|jb)
    return R


def sweep(R, t, T, eo, ev, b0, b1, comm=None, singles=True, fock=None):
    """One Jacobi sweep; returns (t_new, T_new) replicated on every rank.
    fock = (foo, fov, fvv): off-diagonal Fock blocks of a non-canonical reference (AutoRCCSD),
    replicated on every rank, added once after the all-reduce exactly as cc.cu does."""
    comm = comm or SoloComm()
    S = slice(b0, b1)
    V, J, oooo, ooov = R["V"], R["J"], R["oooo"], R["ooov"]
    W4 = R["W4"]
    o, v = V.shape[0], V.shape[2]
    Vt = 2 * V - V.transpose(1, 0, 2, 3)
    Tt = 2 * T - T.transpose(1, 0, 2, 3)
    if singles:
        OA, OB = R["OA"], R["OB"]
        oovo = ooov.transpose(1, 0, 3, 2)
        ooov_t = 2 * ooov - ooov.transpose(1, 0, 2, 3)
        tt = np.einsum("ma,nf->mnaf", t, t)
        tau, tauh = T + tt, T + 0.5 * tt
        tS = t[:, S]
    else:
        tau = tauh = T
    # ---- partial (f in slab) small intermediates, summed over ranks -----------------------------
    FaeT = -es("mnaf,mnef->ea", tauh[..., S], Vt[..., S])                 # stored [e,a]
    Fmi = es("mnef,inef->mi", Vt[..., S], tauh[..., S])
    Wpp = es("mnef,ijef->mnij", V[..., S], tau[..., S])
    R1 = np.zeros((o, v))
    if singles:
        FaeT += 2 * es("amef,mf->ea", OB, tS) - es("eamf,mf->ea", OA, tS)
        R1 -= es("mnae,mnie->ia", T[..., S], ooov_t[..., S])
        R1 += 2 * es("imef,amef->ia", T[..., S], OB) - es("imef,eamf->ia", T[..., S], OA)
    buf = comm.allreduce(np.concatenate([FaeT.ravel(), Fmi.ravel(), Wpp.ravel(), R1.ravel()]))
    FaeT = buf[:v * v].reshape(v, v)
    Fmi = buf[v * v:v * v + o * o].reshape(o, o)
    Wpp = buf[v * v + o * o:v * v + o * o + o ** 4].reshape(o, o, o, o) + oooo
    R1 = buf[v * v + o * o + o ** 4:].reshape(o, v)
    t_new = t
    if fock is not None:
        foo, fov, fvv = fock
        FaeT = FaeT + fvv - 0.5 * es("me,ma->ea", fov, t)          # FaeT[e,a] = Fae[a,e]
        Fmi = Fmi + foo.T + 0.5 * es("me,ie->mi", fov, t)
        R1 = R1 + fov
    if singles:
        Fme = es("mnef,nf->me", Vt, t)
        if fock is not None:
            Fme = Fme + fock[1]
        Fmi = Fmi + es("mnie,ne->mi", ooov_t, t)
        Wpp = Wpp + es("mnie,je->mnij", ooov, t) + es("mnej,ie->mnij", oovo, t)
        R1 = (R1 + es("ie,ea->ia", t, FaeT) - es("ma,mi->ia", t, Fmi) + es("imae,me->ia", Tt, Fme)
              + 2 * es("imae,me->ia", V, t) - es("maie,me->ia", J, t))
        t_new = R1 / (eo[:, None] - ev[None, :])
        FaeT_t = FaeT - 0.5 * es("me,mb->eb", Fme, t)
        Fmi_t = Fmi + 0.5 * es("je,me->mj", t, Fme)
    else:
        FaeT_t, Fmi_t = FaeT, Fmi
    # ---- ring intermediates for the slab, layout [m,e,j,b] ----------------------------------------
    WJ = V[..., S].transpose(0, 2, 1, 3) + 0.5 * es("mnef,njfb->mejb", Vt, T[..., S])
    WE = -J[..., S].copy()
    if singles:
        Tp2, Tph = T + 2 * tt, 0.5 * T + tt
        WJ += es("efmb,jf->mejb", OA, t) - es("mnej,nb->mejb", oovo, tS) - 0.5 * es("mnef,jnfb->mejb", V, Tp2[..., S])
        WE += -es("femb,jf->mejb", OA, t) + es("nmej,nb->mejb", oovo, tS) + es("nmef,jnfb->mejb", V, Tph[..., S])
    else:
        WJ -= 0.5 * es("mnef,jnfb->mejb", V, T[..., S])
        WE += 0.5 * es("nmef,jnfb->mejb", V, T[..., S])
    # ---- ladders and half residual for the slab -----------------------------------------------------
    # particle-particle ladder through the packed symmetric/antisymmetric pair space: this rank's block of
    # the output pairs, all-gather of the blocks, unpack of the rank's slab (cc.cu: sa_ladder)
    import factorized_model as fm
    Tp, Tm = fm.sa_pack_tau(tau)
    Wp, Wm = fm.sa_pack_vvvv(W4, v, b0)        # static in the library (built once)
    Lp = comm.allgather_last(Tp @ Wp)
    Lm = comm.allgather_last(Tm @ Wm)
    Lpp = fm.sa_unpack(Lp, Lm, o, v, b0, b1 - b0)
    Lhh = es("mnij,mnab->ijab", Wpp, tau[..., S])
    H = es("ijae,eb->ijab", T, FaeT_t[:, S]) - es("imab,mj->ijab", T[..., S], Fmi_t)
    H += es("imae,mejb->ijab", Tt, WJ) + es("imae,mejb->ijab", T, WE)
    H += es("mjae,meib->ijab", T, WE)                 # the (ij)(ab) image of T[mibe] WmBEj[maej]
    if singles:
        Yp = es("ijef,efmb->ijmb", tau, OA)
        H -= es("ijmb,ma->ijab", Yp, t)
        Q1 = es("ie,mjeb->imjb", t, V[..., S])
        H -= es("imjb,ma->ijab", Q1, t)
        Q2 = es("ie,maje->imaj", t, J)
        H -= es("imaj,mb->ijab", Q2, tS)
        H += es("ie,ajeb->ijab", t, OB)
        H -= es("ma,mjib->ijab", t, ooov[..., S])
    Hfull = comm.allgather_last(H)
    PH = Hfull.transpose(1, 0, 3, 2)[..., S]
    D = eo[:, None, None, None] + eo[None, :, None, None] - ev[None, None, :, None] - ev[None, None, None, S]
    Tn = (V[..., S] + Lpp + Lhh + H + PH) / D
    return t_new, comm.allgather_last(Tn)
