"""
numpy model of the (T) algorithm the CUDA library executes (jues.jl_b200/csrc/pt.cu) -- test
infrastructure only.  Same data layouts (column-major device arrays are modelled by numpy arrays
indexed in the same letter order), same GEMM operands, same batching over the third occupied
index, so that every index convention of the device code is checked on the CPU against the
literal oracle (oracle/jues_oracle_auto.compute_pT, PerturbativeTriples.jl:35-138).

Device tensors:
    OAp[a,b,p,d] = <pd|ab> = ovvv[p,d,a,b]   (the CC driver's OA[e,f,m,b] = <ef|mb> is this array)
    Ov[l,c,q,r]  = <qr|lc> = ooov[q,r,l,c]
    Vv[a,b,i,j]  = <ij|ab> = oovv[i,j,a,b]
    Tq[a,b,j,i]  = T2[i,j,a,b]
One "X" block per ordered occupied triple (p,q,r):
    X(p,q,r)[a,b,c] = sum_d OAp[a,b,p,d] Tq[c,d,q,r]  -  sum_l Tq[a,b,l,p] Ov[l,c,q,r]
(two GEMMs, M = v^2, N = v, K = v resp. o) and
    W_ijk[a,b,c] = X(i,j,k)[a,b,c] + X(i,k,j)[a,c,b] + X(k,i,j)[c,a,b]
                 + X(k,j,i)[c,b,a] + X(j,k,i)[b,c,a] + X(j,i,k)[b,a,c]        (:96-101)
"""
import math

import numpy as np


def device_tensors(T1, T2, ovvv, ooov, oovv):
    return dict(OAp=np.ascontiguousarray(ovvv.transpose(2, 3, 0, 1)),
                Ov=np.ascontiguousarray(ooov.transpose(2, 3, 0, 1)),
                Vv=np.ascontiguousarray(oovv.transpose(2, 3, 0, 1)),
                Tq=np.ascontiguousarray(T2.transpose(2, 3, 1, 0)), t=T1)


def X_block(Dv, p, q, r):
    v = Dv["Tq"].shape[0]
    A1 = Dv["OAp"][:, :, p, :].reshape(v * v, v)          # [(a,b), d]
    B1 = Dv["Tq"][:, :, q, r]                              # [c, d]  (N x K: transB)
    A2 = Dv["Tq"][:, :, :, p].reshape(v * v, -1)           # [(a,b), l]
    B2 = Dv["Ov"][:, :, q, r]                              # [l, c]
    return (A1 @ B1.T - A2 @ B2).reshape(v, v, v)


def pt_energy(Dv, fo, fv, nocc=None):
    t, Vv = Dv["t"], Dv["Vv"]
    o, v = t.shape
    nocc = o if nocc is None else nocc
    a_, b_, c_ = np.meshgrid(np.arange(v), np.arange(v), np.arange(v), indexing="ij")
    mask = (a_ >= b_) & (b_ >= c_)
    wab = 1.0 / (1.0 + (a_ == b_) + (b_ == c_))
    fvs = fv[:, None, None] + fv[None, :, None] + fv[None, None, :]
    slots = []
    for i in range(nocc):
        for j in range(i + 1):
            # one batch over k = 0..j (the device chunks it when memory is short)
            e_pair = 0.0
            for k in range(j + 1):
                X1, X2, X3 = X_block(Dv, i, j, k), X_block(Dv, i, k, j), X_block(Dv, k, i, j)
                X4, X5, X6 = X_block(Dv, k, j, i), X_block(Dv, j, k, i), X_block(Dv, j, i, k)
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
