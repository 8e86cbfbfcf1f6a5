"""
numpy model of the SHARDED ONE-PASS integral transformation (csrc/transform.cu: tei_transform_sharded, and
its use in csrc/cc.cu: CC::build_integrals / rmp2_dev).  Test infrastructure only.

Rank r of P owns the block B_r of the AO index nu of g'[mu,lam,nu,sig] = gao[mu,nu,lam,sig] (1/P of the AO
tensor) and ends with  M[p,q,r,s] = <pq|rs>  for all p, q, r and ITS share S_r of the last-slot columns:
    1. three local quarter transforms of its block (mu->p, lam->q, sig->s for the columns of EVERY rank),
    2. one personalised exchange (block [p,q,nu in B_r,s in S_d] -> rank d),
    3. the last quarter nu->r over the re-assembled nu range.
`comm.alltoall(list_of_arrays) -> list_of_arrays` supplies the exchange (SoloComm for one rank).
"""
import numpy as np


class SoloComm:
    rank, size = 0, 1

    def alltoall(self, parts):
        return parts

    def allgather_last(self, x):
        return x


def ao_share(np_, nranks, rank):
    per = -(-np_ // nranks)
    lo = min(np_, per * rank)
    return lo, min(np_, per * (rank + 1)) - lo


def sharded_transform(gao, Cp, Cq, Cr, Cs_parts, comm):
    """gao: full chemists' AO tensor (every rank reads only its nu block of it).  Cs_parts[d]: the last-slot
    coefficient columns of rank d.  Returns M[p,q,r,s in S_me]."""
    n = gao.shape[0]
    me, P = comm.rank, comm.size
    lo, nb = ao_share(n, P, me)
    gp = gao[:, lo:lo + nb, :, :].transpose(0, 2, 1, 3)            # g'[mu,lam,nu in B,sig]
    x = np.einsum("mp,mlns->plns", Cp, gp, optimize=True)
    x = np.einsum("lq,plns->pqns", Cq, x, optimize=True)
    send = [np.ascontiguousarray(np.einsum("st,pqns->pqnt", Cs_parts[d], x, optimize=True)) for d in range(P)]
    recv = comm.alltoall(send)                                      # recv[d]: [p,q,nu in B_d,s in S_me]
    y = np.concatenate(recv, axis=2)                                # complete nu range, rank order = nu order
    return np.einsum("nr,pqns->pqrs", Cr, y, optimize=True)


def cc_classes(gao, Cao, Cav_padded, comm, singles=True):
    """The integral classes of one rank exactly as CC::build_integrals cuts them out of M (virtual extent
    already padded to a multiple of 2P): replicated V, J, oooo, ooov; last-index slabs W4, OA, OB."""
    o, v = Cao.shape[1], Cav_padded.shape[1]
    P, me = comm.size, comm.rank
    vs = v // P
    oP = -(-o // P) * P
    os_ = oP // P
    Cop = np.pad(Cao, ((0, 0), (0, oP - o)))
    C = np.hstack([Cao, Cav_padded])
    Cs = [np.hstack([Cop[:, d * os_:(d + 1) * os_], Cav_padded[:, d * vs:(d + 1) * vs]]) for d in range(P)]
    M = sharded_transform(gao, C, C, C, Cs, comm)
    O, Vv, so, sv = slice(0, o), slice(o, o + v), slice(0, os_), slice(os_, os_ + vs)
    R = dict(V=comm.allgather_last(M[O, O, Vv, sv]), J=comm.allgather_last(M[O, Vv, O, sv]),
             oooo=comm.allgather_last(M[O, O, O, so])[:, :, :, :o], W4=M[Vv, Vv, Vv, sv])
    if singles:
        R.update(ooov=comm.allgather_last(M[O, O, O, sv]), OA=M[Vv, Vv, O, sv], OB=M[Vv, O, Vv, sv])
    return R


def mp2_slab(gao, Cao, Cav_padded, comm):
    """<ij|a b_S> of one rank as rmp2_dev asks for it."""
    P = comm.size
    vs = Cav_padded.shape[1] // P
    return sharded_transform(gao, Cao, Cao, Cav_padded, [Cav_padded[:, d * vs:(d + 1) * vs] for d in range(P)], comm)
