"""Generate tests/golden/bench_ehist_nbf{N}_nocc20.npz: the ORACLE's (oracle/jues_oracle.py, literal
reference algorithm: RCCSD.jl:150-289 with the materialised Wabef) 40-sweep RCCSD trace on the exact
inputs bench.py runs at 1 / 2 / 4 / 8 GPUs (nbf = 120 / 144 / 172 / 212, nocc = 20, seed 2024, counter-based
synthetic ERIs): energy, ||T1||_2, ||T2||_2 per sweep and 64 sampled elements of the final T2.  The 15
integral classes are slices of ONE full MO transform (same numbers as RCCSD.jl:117-142's 15 separate
transforms, which would take hours of numpy here).  bench.py and the `-m gpu` tests compare against these
files; nothing under oracle/ runs on the product path.

    python tests/golden/make_bench_golden.py 120 144 172
    python tests/golden/make_bench_golden.py --factorized 212

`--factorized`: for shapes whose literal sweep (three v^4 temporaries next to <vv|vv>) does not fit this
machine's 62 GB, the trace comes from tests/factorized_model.py -- the numpy statement of the SAME equations
without the materialised Wabef -- after the script has re-derived the committed nbf=120 trace of the literal
oracle with it (|dE| <= 1e-12 asserted below): still CPU numpy, still independent of the CUDA code.
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


def full_mo_lean(nbf, C):
    """(pq|rs) with at most two N^4 arrays alive: the AO tensor is generated in sigma slabs and contracted
    on the fly, the other three indices are rotated to the back one at a time."""
    n = nbf
    scale = synth.counter_scale(nbf)
    x = np.zeros((n * n * n, n))
    for lo in range(0, n, 8):
        hi = min(n, lo + 8)
        gs = synth.counter_eri(nbf, SEED, scale, sig_range=(lo, hi))        # Fortran (n,n,n,hi-lo)
        x += gs.reshape(n * n * n, hi - lo, order="F") @ C[lo:hi, :]
        del gs
    # the row index of x is mu + n*(nu + n*lam) (mu fastest), so the C-order reshape reads [lam, nu, mu, s]
    x4 = x.reshape(n, n, n, n)
    del x
    x4 = np.ascontiguousarray(x4.transpose(3, 0, 1, 2))        # [s, lam, nu, mu]
    x4 = (x4.reshape(n * n * n, n) @ C).reshape(n, n, n, n)     # mu -> p:   [s, lam, nu, p]
    x4 = np.ascontiguousarray(x4.transpose(3, 0, 1, 2))        # [p, s, lam, nu]
    x4 = (x4.reshape(n * n * n, n) @ C).reshape(n, n, n, n)     # nu -> q:   [p, s, lam, q]
    x4 = np.ascontiguousarray(x4.transpose(0, 3, 1, 2))        # [p, q, s, lam]
    x4 = (x4.reshape(n * n * n, n) @ C).reshape(n, n, n, n)     # lam -> r:  [p, q, s, r]
    return x4.transpose(0, 1, 3, 2)                            # (pq|rs) view


def factorized_classes(mo, o, v):
    """tests/factorized_model.py: unique_integrals, as slices of the MO tensor."""
    O, V = slice(0, o), slice(o, o + v)
    sl = {"o": O, "v": V}

    def phys(a, b, c, d):
        return np.ascontiguousarray(mo[sl[a], sl[b], sl[c], sl[d]].transpose(0, 2, 1, 3))
    return dict(V=phys("o", "v", "o", "v"), J=phys("o", "o", "v", "v"), ooov=phys("o", "o", "o", "v"),
                ovvv=phys("o", "v", "v", "v"), oooo=phys("o", "o", "o", "o"), vvvv=phys("v", "v", "v", "v"))


def run_factorized(nbf, here, check_against=None):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import factorized_model as fm
    t0 = time.time()
    o, v = NOCC, nbf - NOCC
    Cao, Cav, eps = synth.orbitals(nbf, NOCC, SEED)
    mo = full_mo_lean(nbf, np.hstack([Cao, Cav]))
    I = factorized_classes(mo, o, v)
    del mo
    Dia, D = orc.form_Dia(o, v, eps), orc.form_Dijab(o, v, eps)
    T1, T2 = np.zeros((o, v)), I["V"] / D
    e, n1, n2 = [orc.rccsd_energy(I["V"], T1, T2)], [0.0], [float(np.linalg.norm(T2))]
    print(f"nbf={nbf} (factorised model): integrals {time.time() - t0:.0f} s, E0={e[0]:.16f}", flush=True)
    for it in range(1, MAXIT + 1):
        T1, T2 = fm.rccsd_iteration(I, T1, T2, Dia, D)
        e.append(orc.rccsd_energy(I["V"], T1, T2))
        n1.append(float(np.linalg.norm(T1))); n2.append(float(np.linalg.norm(T2)))
        print(f"  sweep {it:2d} E={e[-1]:.16f}  ({time.time() - t0:.0f} s)", flush=True)
    if check_against is not None:
        ref = np.load(check_against)
        d = float(np.abs(np.array(e) - ref["e_hist"]).max())
        print(f"  factorised model vs committed literal-oracle trace: max|dE| = {d:.2e}", flush=True)
        assert d <= 1e-12
        return
    rng = np.random.default_rng(7)
    idx = np.stack([rng.integers(0, o, 64), rng.integers(0, o, 64), rng.integers(0, v, 64), rng.integers(0, v, 64)], 1)
    np.savez(os.path.join(here, f"bench_ehist_nbf{nbf}_nocc{NOCC}.npz"), nbf=nbf, nocc=NOCC, seed=SEED,
             e_hist=np.array(e), t1_norm=np.array(n1), t2_norm=np.array(n2), t2_idx=idx,
             t2_samples=T2[idx[:, 0], idx[:, 1], idx[:, 2], idx[:, 3]], T1=T1,
             model="tests/factorized_model.py (validated against the literal oracle's nbf=120 trace by this script)")


def main():
    here = os.path.dirname(os.path.abspath(__file__))
    if len(sys.argv) > 1 and sys.argv[1] == "--factorized":
        run_factorized(120, here, check_against=os.path.join(here, f"bench_ehist_nbf120_nocc{NOCC}.npz"))
        for nbf in [int(x) for x in sys.argv[2:]]:
            run_factorized(nbf, here)
        return
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
