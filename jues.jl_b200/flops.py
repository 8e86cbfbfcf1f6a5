"""Flop model of the path: dense FP64, one multiply-add = 2 flop.

Host-side accounting only (bench.py, tools/): `roofline.achieved` is the flops the kernels
EXECUTED (counted by the library per GEMM launch, `Context.counters()["gemm_flops"]`), which must
not exceed the algorithmic counts below; the literal counts of the reference's own evaluation
order (`*_ref`) are shown beside them as "effective vs the reference algorithm" figures.
"""
from __future__ import annotations

__all__ = ["tei_flops_ref", "tei_flops_best", "full_transform_flops", "rccd_iter_alg", "rccd_iter_ref",
           "rccsd_iter_alg", "rccsd_iter_ref", "ladder_flops", "ring_flops", "hh_ladder_flops",
           "rccd_transforms_ref", "rccsd_transforms_ref", "mp2_energy_bytes", "pt_flops", "fock_flops"]


def tei_flops_ref(N, d1, d2, d3, d4):
    """F_tei in the reference's fixed order sigma, lambda, nu, mu (Transformation.jl:68-91)."""
    return 2 * N * (N**3 * d4 + N**2 * d3 * d4 + N * d2 * d3 * d4 + d1 * d2 * d3 * d4)


def tei_flops_best(N, d1, d2, d3, d4, streamed=False):
    """Cheapest of the 24 contraction orders (what tei_transform_dev picks); a streamed AO tensor
    must contract its last index first."""
    from itertools import permutations
    d = (d1, d2, d3, d4)
    best = None
    for perm in permutations(range(4)):
        if streamed and perm[0] != 3:
            continue
        e = [N, N, N, N]
        f = 0
        for ax in perm:
            f += 2 * e[0] * e[1] * e[2] * e[3] * d[ax]
            e[ax] = d[ax]
        best = f if best is None else min(best, f)
    return best


def full_transform_flops(N):
    return 8 * N**5


def ladder_flops(o, v):
    return 2 * o**2 * v**4


def ring_flops(o, v):
    return 2 * o**3 * v**3


def hh_ladder_flops(o, v):
    return 2 * o**4 * v**2


def rccd_iter_ref(o, v):
    return 4 * o**2 * v**4 + 22 * o**3 * v**3 + 4 * o**4 * v**2 + 6 * o**2 * v**3 + 6 * o**3 * v**2


def rccd_iter_alg(o, v):
    """No Wabef: T.Wabef + T.Wmnij = T.vvvv + (oooo + 2X).T, X = 1/2 T.oovv."""
    return 2 * o**2 * v**4 + 22 * o**3 * v**3 + 4 * o**4 * v**2 + 6 * o**2 * v**3 + 6 * o**3 * v**2


def rccsd_iter_ref(o, v):
    """The reference's literal sweep (Wabef built and applied, 13 ring-type terms)."""
    return 4 * o**2 * v**4 + 4 * o * v**4 + 26 * o**3 * v**3 + 4 * o**4 * v**2 + 16 * o**2 * v**3


def rccsd_iter_alg(o, v):
    """F_alg of SURVEY.md section 8a (factorised RCCSD sweep; also AutoRCCSD's sweep: the one-body
    Fock terms add only O(o v^2 + o^2 v) flops)."""
    return 2 * o**2 * v**4 + 22 * o**3 * v**3 + 4 * o**4 * v**2 + 24 * o**2 * v**3 + 24 * o**3 * v**2


def rccsd_iter_exec(o, v):
    """FP64 operations the GEMM launches of ONE factorised RCCSD sweep of the library execute on one rank
    (csrc/cc.cu with the packed ladder; the library counts them per launch, this is the same sum in closed
    form, within 0.1 % at nbf=120 and 0.5 % over 8 ranks): packed pp-ladder 4 o^2 np nq with np = v(v+1)/2
    summed pairs and nq = v(v/2+1) output pairs, seven (ov)^3-sized products (six rings + tau.<ef|mb>), and the
    lower-order terms.  This is the unit of work both bench arms divide by their seconds, so that the ratio of
    their `value`s is the ratio of their times."""
    np_, nq = v * (v + 1) // 2, v * (v // 2 + 1)
    return (4 * o**2 * np_ * nq + 14 * o**3 * v**3 + 12 * o**2 * v**3 + 18 * o**3 * v**2 + 4 * o**4 * v**2
            + 2 * o * v**3 + 2 * o**2 * v**2)


def cc_transform_exec(N):
    """FP64 operations of the one-pass transform that yields every integral class of a CC run (8 N^5 over all
    ranks; exact on one rank, +1 % padding over 8)."""
    return 8 * N**5


def rccd_transforms_ref(N, o, v):
    """The 5 literal transforms of make_rccd_integrals (RCCD.jl:85-97)."""
    return (tei_flops_ref(N, o, v, o, v) + tei_flops_ref(N, v, v, v, v) + tei_flops_ref(N, o, v, v, o)
            + tei_flops_ref(N, o, o, v, v) + tei_flops_ref(N, o, o, o, o))


def rccsd_transforms_ref(N, o, v):
    """The 15 literal transforms of make_rccsd_integrals (RCCSD.jl:117-142)."""
    O, V = o, v
    slots = [(V, V, V, V), (O, V, V, V), (V, V, O, V), (V, O, V, V), (V, V, V, O), (O, V, O, V), (O, V, V, O),
             (V, V, O, O), (O, O, V, V), (V, O, O, V), (O, O, O, V), (O, V, O, O), (O, O, V, O), (V, O, O, O),
             (O, O, O, O)]
    return sum(tei_flops_ref(N, *s) for s in slots)


def mp2_energy_bytes(o, v):
    """Algorithmic bytes of the MP2 energy reduction: every (ia|jb) read once."""
    return 8 * o**2 * v**2


def pt_flops(o, v):
    """(T) (PerturbativeTriples.jl:96-101): per triple i >= j >= k six pairs of contractions
    v^3 x v and v^3 x o; o(o+1)(o+2)/6 triples.  Same count here (pt.cu)."""
    return o * (o + 1) * (o + 2) // 6 * 6 * 2 * v**3 * (v + o)


def fock_flops(N, nmo, o):
    """get_fock as two transforms (C,C,Co,Co) and (C,Co,C,Co) in their cheapest orders + C^T h C."""
    return tei_flops_best(N, nmo, nmo, o, o) + tei_flops_best(N, nmo, o, nmo, o) + 4 * N * N * nmo
