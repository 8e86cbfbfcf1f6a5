"""
numpy model of the (T) algorithm the CUDA library executes (jues.jl_b200/csrc/pt.cu) -- test
infrastructure only.  Same data layouts (column-major device arrays are modelled by numpy arrays
indexed in the same letter order), same GEMM operands, same batching over the third occupied
index, so that every index convention of the device code is checked on the CPU against the
literal oracle (oracle/jues_oracle_auto.compute_pT, PerturbativeTriples.jl:35-138).

Device tensors (K = v + o; see jues.jl_b200/csrc/pt.h):
    Acat[a,b,p,kap]:  kap <  v: <p kap|ab> = ovvv[p,kap,a,b]     kap >= v: -T2[p,l,a,b], l = kap - v
    Bq[c,q,kap,r]:    kap <  v: T2[r,q,c,kap]                     kap >= v: <qr|lc> = ooov[q,r,l,c]
    Br[c,r,kap,q]:    the same numbers, q and r exchanged in the layout
    Vv[a,b,i,j] = <ij|ab>
One "X" block per ordered occupied triple (p,q,r) -- ONE GEMM, the contractions over d and l
concatenated along K:
    X(p,q,r)[(a,b),c] = sum_kap Acat[(a,b),p,kap] B[c,kap,(q,r)]
and
    W_ijk[a,b,c] = X(i,j,k)[a,b,c] + X(i,k,j)[a,c,b] + X(k,i,j)[c,a,b]
                 + X(k,j,i)[c,b,a] + X(j,k,i)[b,c,a] + X(j,i,k)[b,a,c]        (:96-101)
The blocks of consecutive k come out of one launch: rows (c,q) of Bq / (c,r) of Br give N = v*nb,
rows (a,b,p) of Acat give M = v^2*nb (x_blocks below mirrors pt.cu's pointer arithmetic on flat
column-major buffers).
"""
import math

import numpy as np


def device_tensors(T1, T2, ovvv, ooov, oovv):
    o, v = T1.shape
    K = v + o
    Acat = np.zeros((v, v, o, K))
    Acat[:, :, :, :v] = ovvv.transpose(2, 3, 0, 1)               # [a,b,p,d] = ovvv[p,d,a,b]
    Acat[:, :, :, v:] = -T2.transpose(2, 3, 0, 1)                 # [a,b,p,l] = -T2[p,l,a,b]
    Bq = np.zeros((v, o, K, o))
    Bq[:, :, :v, :] = T2.transpose(2, 1, 3, 0)                    # [c,q,d,r] = T2[r,q,c,d]
    Bq[:, :, v:, :] = ooov.transpose(3, 0, 2, 1)                  # [c,q,l,r] = ooov[q,r,l,c]
    Br = np.ascontiguousarray(Bq.transpose(0, 3, 2, 1))           # [c,r,kap,q]
    F = lambda x: np.asfortranarray(x).ravel(order="F")           # flat column-major device buffers
    return dict(o=o, v=v, Acat=F(Acat), Bq=F(Bq), Br=F(Br),
                Vv=np.ascontiguousarray(oovv.transpose(2, 3, 0, 1)), t=T1)


def x_blocks(Dv, p0, ps, q0, qs, r0, rs, nb):
    """pt.cu: x_blocks -- one GEMM on strided views of the flat buffers.  Returns the nb blocks as
    an array [n][a,b,c]."""
    o, v = Dv["o"], Dv["v"]
    v2, K = v * v, v + o
    M = v2 * nb if ps else v2
    N = v if ps else v * nb
    A0 = p0 * v2
    lda = v2 * o
    A = np.array([[Dv["Acat"][A0 + m + lda * k] for k in range(K)] for m in range(M)])
    B0 = (v * r0 + v * o * K * q0) if rs else (v * q0 + v * o * K * r0)
    Bb = Dv["Br"] if rs else Dv["Bq"]
    ldb = v * o
    B = np.array([[Bb[B0 + n + ldb * k] for k in range(K)] for n in range(N)])
    C = A @ B.T                                                    # (M, N), column-major out[m + M*n]
    if ps:
        return C.reshape(v, v, nb, v, order="F").transpose(2, 0, 1, 3)   # out[(a,b,n),c]
    return C.reshape(v, v, v, nb, order="F").transpose(3, 0, 1, 2)       # out[(a,b),(c,n)]


def X_block(Dv, p, q, r):
    return x_blocks(Dv, p, 0, q, 0, r, 1, 1)[0]


def pt_energy(Dv, fo, fv, nocc=None, rank=0, nranks=1):
    """This rank's share of E(T): occupied pairs (i >= j) are dealt round-robin (pt.cu); the shares are
    summed over ranks by the caller."""
    t, Vv = Dv["t"], Dv["Vv"]
    o, v = t.shape
    nocc = o if nocc is None else nocc
    a_, b_, c_ = np.meshgrid(np.arange(v), np.arange(v), np.arange(v), indexing="ij")
    mask = (a_ >= b_) & (b_ >= c_)
    wab = 1.0 / (1.0 + (a_ == b_) + (b_ == c_))
    fvs = fv[:, None, None] + fv[None, :, None] + fv[None, None, :]
    slots = []
    pair = -1
    for i in range(nocc):
        for j in range(i + 1):
            pair += 1
            if pair % nranks != rank:
                continue
            # one batch over k = 0..j (the device chunks it when memory is short): six launches
            nb = j + 1
            F1, F2 = x_blocks(Dv, i, 0, j, 0, 0, 1, nb), x_blocks(Dv, i, 0, 0, 1, j, 0, nb)
            F3, F4 = x_blocks(Dv, 0, 1, i, 0, j, 0, nb), x_blocks(Dv, 0, 1, j, 0, i, 0, nb)
            F5, F6 = x_blocks(Dv, j, 0, 0, 1, i, 0, nb), x_blocks(Dv, j, 0, i, 0, 0, 1, nb)
            e_pair = 0.0
            for k in range(j + 1):
                X1, X2, X3, X4, X5, X6 = F1[k], F2[k], F3[k], F4[k], F5[k], F6[k]
                W = (X1 + X2.transpose(0, 2, 1) + X3.transpose(1, 2, 0) + X4.transpose(2, 1, 0)
                     + X5.transpose(2, 0, 1) + X6.transpose(1, 0, 2))
                V = (W + Vv[:, :, j, k][None, :, :] * t[i][:, None, None]
                     + Vv[:, :, i, k][:, None, :] * t[j][None, :, None]
                     + Vv[:, :, i, j][:, :, None] * t[k][None, None, :])
                P = lambda A, s: A.transpose(*s)
                Wabc, Wacb, Wbac = W, P(W, (0, 2, 1)), P(W, (1, 0, 2))
                Wbca, Wcab, Wcba = P(W, (2, 0, 1)), P(W, (1, 2, 0)), P(W, (2, 1, 0))
                Vabc, Vacb, Vbac = V, P(V, (0, 2, 1)), P(V, (1, 0, 2))
                Vbca, Vcab, Vcba = P(V, (2, 0, 1)), P(V, (1, 2, 0)), P(V, (2, 1, 0))
                Xs = Wabc * Vabc + Wacb * Vacb + Wbac * Vbac + Wbca * Vbca + Wcab * Vcab + Wcba * Vcba
                Y = Vabc + Vbca + Vcab
                Z = Vacb + Vbac + Vcba
                E = (Y - 2 * Z) * (Wabc + Wbca + Wcab) + (Z - 2 * Y) * (Wacb + Wbac + Wcba) + 3 * Xs
                occ = 2.0 - (i == j) - (j == k)
                Dd = fo[i] + fo[j] + fo[k] - fvs
                e_pair += float(np.sum((E * occ * wab / Dd)[mask]))
            slots.append(e_pair)
    return math.fsum(slots)


def pt_energy_allsum(T1, T2, ovvv, ooov, oovv, fo, fv):
    """Independent check: the textbook closed-shell (T) energy with unrestricted sums
    E = sum_{ijk,abc} (4 W_abc + W_bca + W_cab)(V_abc - V_cba) / (3 D)   (Rendell-Lee-Komornicki)."""
    o, v = T1.shape
    Et = 0.0
    fvs = fv[:, None, None] + fv[None, :, None] + fv[None, None, :]
    for i in range(o):
        for j in range(o):
            for k in range(o):
                def term(p, q, r):   # sum_d <pd|ab>... in letters: W^{pqr}_{abc} single permutation member
                    return (np.einsum("dab,cd->abc", ovvv[p], T2[r, q])
                            - np.einsum("lc,lab->abc", ooov[q, r], T2[p]))
                W = (term(i, j, k) + term(i, k, j).transpose(0, 2, 1) + term(k, i, j).transpose(1, 2, 0)
                     + term(k, j, i).transpose(2, 1, 0) + term(j, k, i).transpose(2, 0, 1)
                     + term(j, i, k).transpose(1, 0, 2))
                V = (W + oovv[j, k][None, :, :] * T1[i][:, None, None]
                     + oovv[i, k][:, None, :] * T1[j][None, :, None]
                     + oovv[i, j][:, :, None] * T1[k][None, None, :])
                D = fo[i] + fo[j] + fo[k] - fvs
                Et += float(np.sum((4 * W + W.transpose(2, 0, 1) + W.transpose(1, 2, 0))
                                   * (V - V.transpose(2, 1, 0)) / (3 * D)))
    return Et
