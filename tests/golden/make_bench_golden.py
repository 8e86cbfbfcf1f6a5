"""Generate tests/golden/bench_ehist_nbf{N}_nocc20.npz: the ORACLE's (oracle/jues_oracle.py, literal
reference algorithm: RCCSD.jl:150-289 with the materialised Wabef) 40-sweep RCCSD trace on the exact
inputs bench.py runs at 1 / 2 / 4 / 8 GPUs (nbf = 120 / 144 / 172 / 212, nocc = 20, seed 2024, counter-based
synthetic ERIs): energy, ||T1||_2, ||T2||_2 per sweep and 64 sampled elements of the final T2.  The 15
integral classes are slices of ONE full MO transform (same numbers as RCCSD.jl:117-142's 15 separate
transforms, which would take hours of numpy here).  bench.py and the `-m gpu` tests compare against these
files; nothing under oracle/ runs on the product path.

    python tests/golden/make_bench_golden.py 120 144 172 212
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import jues.jl_b200.synth as synth             # noqa: E402  (input generator only, pure numpy)
from oracle import jues_oracle as orc          # noqa: E402

NOCC, SEED, MAXIT = 20, 2024, 40


def classes_from_mo(mo, o, v):
    """The reference's 15 arrays (RCCSD.jl:117-142) as slices of the chemists'-order MO tensor."""
    O, V = slice(0, o), slice(o, o + v)
    sl = {"o": O, "v": V}

    def phys(a, b, c, d):                                # permutedims(tei_transform(C_a,C_b,C_c,C_d),[1,3,2,4])
        return np.ascontiguousarray(mo[sl[a], sl[b], sl[c], sl[d]].transpose(0, 2, 1, 3))
    I = {"vvvv": phys("v", "v", "v", "v"), "ovvv": phys("o", "v", "v", "v"), "vovv": phys("v", "v", "o", "v"),
         "vvov": phys("v", "o", "v", "v"), "vvvo": phys("v", "v", "v", "o"), "oovv": phys("o", "v", "o", "v"),
         "ovvo": phys("o", "v", "v", "o"), "vovo": phys("v", "v", "o", "o"), "ovov": phys("o", "o", "v", "v"),
         "voov": phys("v", "o", "o", "v"), "ooov": phys("o", "o", "o", "v"), "oovo": phys("o", "v", "o", "o"),
         "ovoo": phys("o", "o", "v", "o"), "vooo": phys("v", "o", "o", "o"), "oooo": phys("o", "o", "o", "o")}
    I["vvov"] = np.ascontiguousarray(I["vvov"].transpose(3, 0, 1, 2))     # RCCSD.jl:133
    I["vvvo"] = np.ascontiguousarray(I["vvvo"].transpose(2, 0, 1, 3))     # :134
    I["vovv"] = np.ascontiguousarray(I["vovv"].transpose(1, 0, 2, 3))     # :135
    I["vooo"] = np.ascontiguousarray(I["vooo"].transpose(1, 0, 2, 3))     # :136
    return I


def main():
    here = os.path.dirname(os.path.abspath(__file__))
    for nbf in [int(x) for x in sys.argv[1:]] or [120]:
        t0 = time.time()
        o, v = NOCC, nbf - NOCC
        Cao, Cav, eps = synth.orbitals(nbf, NOCC, SEED)
        g = synth.counter_eri(nbf, SEED, synth.counter_scale(nbf))
        C = np.hstack([Cao, Cav])
        mo = np.einsum("mp,mnls->pnls", C, g, optimize=True)
        del g
        mo = np.einsum("nq,pnls->pqls", C, mo, optimize=True)
        mo = np.einsum("lr,pqls->pqrs", C, mo, optimize=True)
        mo = np.einsum("st,pqrs->pqrt", C, mo, optimize=True)
        I = classes_from_mo(mo, o, v)
        del mo
        Dia, D = orc.form_Dia(o, v, eps), orc.form_Dijab(o, v, eps)
        T1, T2 = np.zeros((o, v)), I["oovv"] / D
        e, n1, n2 = [orc.rccsd_energy(I["oovv"], T1, T2)], [0.0], [float(np.linalg.norm(T2))]
        print(f"nbf={nbf}: integrals {time.time() - t0:.0f} s, E0={e[0]:.16f}", flush=True)
        for it in range(1, MAXIT + 1):
            T1, T2 = orc.rccsd_iteration(I, T1, T2, Dia, D)
            e.append(orc.rccsd_energy(I["oovv"], T1, T2))
            n1.append(float(np.linalg.norm(T1))); n2.append(float(np.linalg.norm(T2)))
            print(f"  sweep {it:2d} E={e[-1]:.16f}  ({time.time() - t0:.0f} s)", flush=True)
        rng = np.random.default_rng(7)
        idx = np.stack([rng.integers(0, o, 64), rng.integers(0, o, 64), rng.integers(0, v, 64), rng.integers(0, v, 64)], 1)
        np.savez(os.path.join(here, f"bench_ehist_nbf{nbf}_nocc{NOCC}.npz"), nbf=nbf, nocc=NOCC, seed=SEED,
                 e_hist=np.array(e), t1_norm=np.array(n1), t2_norm=np.array(n2), t2_idx=idx,
                 t2_samples=T2[idx[:, 0], idx[:, 1], idx[:, 2], idx[:, 3]], T1=T1)


if __name__ == "__main__":
    main()
