"""Oracle self-consistency on the CPU: two independent factorisations of the RCCD / RCCSD sweep
(the literal transcription in oracle/jues_oracle.py and the Wabef-free, symmetrised form the GPU
executes, tests/factorized_model.py) agree to rounding; the reference's quirks are reproduced."""
import numpy as np
import pytest

import jues.jl_b200 as jb
from oracle import jues_oracle as orc
import factorized_model as fm


def inputs(N, o, seed, scale=None):
    g, Cao, Cav, eps = jb.synth.dense_inputs(N, o, seed=seed, scale=scale)
    return g, Cao, Cav, eps, orc.Wfn(o, N - o, eps, Cao, Cav, g)


@pytest.mark.parametrize("N,o,seed", [(8, 3, 1), (10, 3, 7), (14, 4, 3)])
def test_factorized_rccd_equals_literal(N, o, seed):
    g, Cao, Cav, eps, w = inputs(N, o, seed)
    v = N - o
    I6 = fm.unique_integrals(g, Cao, Cav)
    ints = orc.make_rccd_integrals(g, Cao, Cav)
    D = orc.form_Dijab(o, v, eps)
    T = orc.rccd_guess(ints[1], D)
    for _ in range(4):
        Tr = orc.rccd_iteration(T, ints, D)
        Tm = fm.rccd_iteration(I6, T, D)
        assert np.abs(Tr - Tm).max() < 1e-15
        T = Tr


@pytest.mark.parametrize("N,o,seed", [(8, 3, 1), (10, 3, 7), (14, 4, 3)])
def test_factorized_rccsd_equals_literal(N, o, seed):
    g, Cao, Cav, eps, w = inputs(N, o, seed)
    v = N - o
    I6 = fm.unique_integrals(g, Cao, Cav)
    I = orc.make_rccsd_integrals(g, Cao, Cav)
    D, Dia = orc.form_Dijab(o, v, eps), orc.form_Dia(o, v, eps)
    t1, T = np.zeros((o, v)), I["oovv"] / D
    for _ in range(5):
        a, b = orc.rccsd_iteration(I, t1, T, Dia, D)
        c, d = fm.rccsd_iteration(I6, t1, T, Dia, D)
        assert np.abs(a - c).max() < 1e-15 and np.abs(b - d).max() < 1e-15
        t1, T = a, b


@pytest.mark.parametrize("N,o,seed", [(8, 3, 1), (12, 4, 9)])
def test_relaid_sweep_equals_literal(N, o, seed):
    """The form of the sweep without output permutation passes (cc.cu `relaid`: Fmi term through its (ij)(ab)
    image, ring products combined in their GEMM-native layouts, terms contracted with t[m,a] summed first) is the
    same iteration.  The image trick relies on T2[i,j,a,b] = T2[j,i,b,a], which every iterate satisfies."""
    g, Cao, Cav, eps, w = inputs(N, o, seed)
    v = N - o
    I6 = fm.unique_integrals(g, Cao, Cav)
    I = orc.make_rccsd_integrals(g, Cao, Cav)
    ints = orc.make_rccd_integrals(g, Cao, Cav)
    D, Dia = orc.form_Dijab(o, v, eps), orc.form_Dia(o, v, eps)
    t1, T = np.zeros((o, v)), I["oovv"] / D
    for _ in range(5):
        a, b = orc.rccsd_iteration(I, t1, T, Dia, D)
        c, d = fm.rccsd_iteration(I6, t1, T, Dia, D, relaid=True)
        assert np.abs(a - c).max() < 1e-14 and np.abs(b - d).max() < 1e-14
        t1, T = a, b
    T = orc.rccd_guess(ints[1], D)
    for _ in range(4):
        Tr = orc.rccd_iteration(T, ints, D)
        assert np.abs(Tr - fm.rccd_iteration(I6, T, D, relaid=True)).max() < 1e-14
        T = Tr


def test_integral_class_identities():
    """The 15 classes of make_rccsd_integrals (RCCSD.jl:117-142) reduce to six."""
    g, Cao, Cav, eps, w = inputs(9, 3, 5)
    I = orc.make_rccsd_integrals(g, Cao, Cav)
    U = fm.unique_integrals(g, Cao, Cav)
    close = lambda a, b: np.abs(a - b).max() < 1e-14
    assert close(I["oovv"], U["V"]) and close(I["ovov"], U["J"]) and close(I["ovvv"], U["ovvv"])
    assert close(I["ovvo"], U["V"].transpose(0, 3, 2, 1))
    assert close(I["vovo"], U["J"].transpose(1, 0, 3, 2))
    assert close(I["voov"], U["V"].transpose(2, 1, 0, 3))
    assert close(I["oovo"], U["ooov"].transpose(1, 0, 3, 2))
    assert close(I["ovoo"], U["ooov"].transpose(0, 3, 2, 1))
    assert close(I["vooo"], U["ooov"].transpose(0, 3, 1, 2))
    assert close(I["vovv"], U["ovvv"].transpose(0, 1, 3, 2))
    assert close(I["vvvo"], U["ovvv"].transpose(3, 1, 2, 0))
    assert close(I["vvov"], U["ovvv"].transpose(3, 2, 1, 0))


def test_rccd_quirk_guess_and_mp2_guess_converge_to_same_energy():
    """RCCD.jl:45,145-160 starts from (ij|ab)/D, not the MP2 amplitudes."""
    g, Cao, Cav, eps, w = inputs(10, 3, 7, scale=0.03)
    hist_q, hist_m = [], []
    eq = orc.do_rccd(w, guess="reference", callback=lambda it, e, T: hist_q.append(e))
    em = orc.do_rccd(w, guess="mp2", callback=lambda it, e, T: hist_m.append(e))
    assert abs(hist_q[0] - hist_m[0]) > 1e-6           # different starting points
    assert abs(hist_m[0] - orc.do_rmp2(w)) < 1e-13      # MP2 guess has the MP2 energy
    assert abs(eq - em) < 1e-12


def test_rccsd_guess_energy_is_mp2():
    g, Cao, Cav, eps, w = inputs(10, 3, 2)
    hist = []
    orc.do_rccsd(w, maxit=1, callback=lambda it, e, T1, T2: hist.append(e))
    assert abs(hist[0] - orc.do_rmp2(w)) < 1e-13


def test_mp2_summation_orders_agree():
    g, Cao, Cav, eps, w = inputs(12, 4, 3)
    assert abs(orc.do_rmp2(w) - orc.do_rmp2(w, strict_order=True)) < 1e-14


def test_get_eri_notation_and_frozen_core():
    g, Cao, Cav, eps, w = inputs(8, 3, 4)
    phys = orc.get_eri(w, "OOVV")
    chem = orc.get_eri(w, "OVOV", notation="chem")
    assert np.allclose(phys, chem.transpose(0, 2, 1, 3), atol=1e-15)
    assert orc.get_eri(w, "OOVV", fcn=1).shape == (2, 2, 5, 5)
    with pytest.raises(ValueError):
        orc.get_eri(w, "OOV")


def test_tei_transform_equals_direct_einsum():
    g, Cao, Cav, eps, w = inputs(7, 2, 9)
    ref = np.einsum("mi,na,lj,sb,mnls->iajb", Cao, Cav, Cao, Cav, g, optimize=True)
    assert np.abs(orc.tei_transform(g, Cao, Cav, Cao, Cav) - ref).max() < 1e-15
    Cf = np.hstack([Cao, Cav])
    assert np.abs(orc.tei_transform(g, Cf) - orc.tei_transform(g, Cf, Cf, Cf, Cf)).max() == 0.0


def test_counter_eri_symmetry_and_slabs():
    N = 7
    g = jb.synth.counter_eri(N, seed=3, scale=0.1)
    for perm in [(1, 0, 2, 3), (0, 1, 3, 2), (2, 3, 0, 1)]:
        assert np.array_equal(g, g.transpose(perm))
    slab = jb.synth.counter_eri(N, seed=3, scale=0.1, sig_range=(2, 5))
    assert np.array_equal(slab, g[:, :, :, 2:5])
    assert np.abs(g).max() <= 0.1 and np.abs(g.mean()) < 0.01


def test_flop_model_matches_the_survey_table():
    """jues.jl_b200/flops.py against SURVEY.md section 8a (C3 column: N=120, o=20, v=100)."""
    from importlib import import_module
    f = import_module("jues.jl_b200.flops")
    assert f.full_transform_flops(120) == 8 * 120**5
    assert abs(f.tei_flops_ref(120, 20, 100, 20, 100) / 5.51e10 - 1) < 2e-3
    assert abs(f.tei_flops_best(120, 20, 100, 20, 100) / 1.18e10 - 1) < 2e-3
    assert abs(f.rccsd_transforms_ref(120, 20, 100) / 7.29e11 - 1) < 2e-3
    assert abs(f.rccd_transforms_ref(120, 20, 100) / 2.99e11 - 1) < 2e-3
    assert abs(f.rccd_iter_alg(20, 100) / 2.65e11 - 1) < 3e-3 and abs(f.rccd_iter_ref(20, 100) / 3.45e11 - 1) < 3e-3
    assert abs(f.rccsd_iter_alg(20, 100) / 2.74e11 - 1) < 3e-3
    assert f.ladder_flops(60, 400) == 2 * 60**2 * 400**4
    assert f.tei_flops_best(50, 5, 45, 5, 45, streamed=True) >= f.tei_flops_best(50, 5, 45, 5, 45)


def test_symmetric_antisymmetric_ladder_identity():
    """The packed-pair pp-ladder the device executes (tests/factorized_model.py mirrors the kernels'
    index functions) == the plain tau.vvvv contraction, for 1, 2 and 3 ranks' column blocks."""
    import numpy as np
    import factorized_model as fm
    from jues.jl_b200 import synth
    from oracle import jues_oracle as orc
    for (N, o, nr) in [(9, 3, 1), (11, 3, 2), (10, 4, 3), (5, 3, 1)]:
        g, Cao, Cav, eps = synth.dense_inputs(N, o, seed=4)
        I = fm.unique_integrals(g, Cao, Cav)
        w = orc.Wfn(o, N - o, eps, Cao, Cav, g)
        e, T1, T2 = orc.do_rccsd(w, maxit=2, return_T=True)
        tau = T2 + np.einsum("ia,jb->ijab", T1, T1)
        W4 = np.ascontiguousarray(I["vvvv"].transpose(2, 3, 0, 1))   # W4[e,f,a,b] = <ef|ab> = vvvv[a,b,e,f]
        v = N - o
        vp = -(-v // (2 * nr)) * (2 * nr)                            # padded like the device
        taup = np.pad(tau, [(0, 0), (0, 0), (0, vp - v), (0, vp - v)])
        W4p = np.pad(W4, [(0, vp - v)] * 4)
        ref = np.einsum("ijef,efab->ijab", taup, W4p)
        got = fm.sa_ladder(taup, W4p, nranks=nr)
        assert np.abs(got - ref).max() < 1e-14 * max(1.0, np.abs(ref).max()), (N, o, nr)
    # every unordered pair has exactly one column, pad slots excepted
    for v in (2, 4, 6, 10):
        cols = {}
        for a in range(v):
            for b in range(a + 1):
                cols.setdefault(fm.sa_column(a, b, v), []).append((a, b))
        assert all(len(x) == 1 for x in cols.values()) and max(cols) < v * (v // 2 + 1)
        for z in range(v):
            for t in range(v // 2 + 1):
                w = fm.sa_partner(z, t, v)
                if w >= 0:
                    assert fm.sa_column(z, w, v) == z * (v // 2 + 1) + t
