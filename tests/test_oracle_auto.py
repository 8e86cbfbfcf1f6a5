"""CPU checks of the section-8f oracle (oracle/jues_oracle_auto.py) and of the numpy models of the
device algorithms (tests/factorized_model.py with Fock terms, tests/sharded_model.py,
tests/pt_model.py).  The reference holds no golden vectors for AutoRCCSD / compute_pT / mRCCD; the
pins available are
  * AutoRCCSD.update_amp == RCCSD.jl's cciter per sweep for canonical orbitals (two derivations of
    the same equations inside the reference),
  * the reference's RCCSD known answer for H2O/STO-3G (test/TestCoupledCluster.jl:44-45) reached
    through do_auto_rccsd on the offline fixture,
  * an independent textbook (T) formula with unrestricted sums, and the CCSD(T) literature value of
    the same H2O/STO-3G geometry (Crawford's programming project #6: E(T) = -0.000099877272).
"""
import os

import numpy as np
import pytest

import factorized_model as fm
import pt_model as pm
import sharded_model as sm
from jues.jl_b200 import synth
from oracle import jues_oracle as orc
from oracle import jues_oracle_auto as oa

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def canonical_wfn(N, o, seed):
    g, Cao, Cav, eps = synth.dense_inputs(N, o, seed=seed)
    return orc.Wfn(o, N - o, eps, Cao, Cav, g, hao=synth.core_hamiltonian(g, Cao, Cav, eps))


def noncanonical_wfn(N, o, seed, ov_mix=0.0):
    g, h, Ca, eps = synth.noncanonical_inputs(N, o, seed=seed, ov_mix=ov_mix)
    return orc.Wfn(o, N - o, eps, Ca[:, :o].copy(), Ca[:, o:].copy(), g, hao=h, Ca=Ca)


@pytest.fixture(scope="module")
def h2o():
    d = np.load(os.path.join(GOLD, "h2o_sto3g.npz"))
    C, eps, g, o = d["C"], d["eps"], d["g"], int(d["nocc"])
    return orc.Wfn(o, C.shape[1] - o, eps, C[:, :o].copy(), C[:, o:].copy(), g, hao=d["H"])


def test_get_fock_is_diagonal_for_the_scf_solution(h2o):
    f = oa.get_fock(h2o)
    assert np.abs(f - np.diag(h2o.epsa)).max() < 1e-11
    w = canonical_wfn(9, 3, 5)
    assert np.abs(oa.get_fock(w) - np.diag(w.epsa)).max() < 1e-13
    with pytest.raises(ValueError):
        oa.get_fock(w, spin="gamma")


def test_auto_rccsd_equals_rccsd_jl_per_sweep():
    """Canonical orbitals: AutoRCCSD.update_amp (:67-178) and RCCSD.jl cciter (:150-173) are the same map."""
    w = canonical_wfn(10, 3, 7)
    a, b = [], []
    orc.do_rccsd(w, maxit=6, callback=lambda it, e, T1, T2: a.append((e, T1.copy(), T2.copy())))
    oa.do_auto_rccsd(w, cc_max_iter=6, cc_e_conv=0.0, cc_max_rms=0.0,
                     callback=lambda it, e, T1, T2: b.append((e, T1.copy(), T2.copy())))
    assert len(a) == len(b) == 7
    for (e0, t0, T0), (e1, t1, T1) in zip(a, b):
        assert abs(e0 - e1) < 1e-14 and np.abs(t0 - t1).max() < 1e-14 and np.abs(T0 - T1).max() < 1e-14


def test_auto_rccsd_known_answer_h2o(h2o):
    """test/TestCoupledCluster.jl:44-45 (RCCSD -0.070680102078571) through AutoRCCSD's own loop."""
    r = oa.do_auto_rccsd(h2o, do_pT=True, return_all=True)
    assert r["converged"] and r["iterations"] <= 50
    assert abs(r["ecc"] - (-0.070680102078571)) < 2e-9
    assert abs(r["ept"] - (-0.000099877272)) < 2e-9          # Crawford project #6, same geometry/basis


def test_convergence_control_and_defaults():
    w = canonical_wfn(10, 3, 7)
    r = oa.do_auto_rccsd(w, return_all=True)
    assert r["converged"] and 1 < r["iterations"] < 50
    assert abs(r["e_hist"][-1] - r["e_hist"][-2]) < 1e-10 and r["rms_hist"][-1] < 1e-10
    r3 = oa.do_auto_rccsd(w, cc_max_iter=3, return_all=True)
    assert r3["iterations"] == 3 and not r3["converged"]
    loose = oa.do_auto_rccsd(w, cc_e_conv=1e-4, cc_max_rms=1e-4, return_all=True)
    assert loose["iterations"] < r["iterations"]
    assert oa.do_auto_rccsd(w, unknown_option=1) == r["ecc"]       # unknown kwargs ignored (:199-205)
    odd = orc.Wfn(3, 6, w.epsa[:9], w.Cao, w.Cav[:, :6], w.ao_eri, nbeta=2, hao=w.hao)
    with pytest.raises(ValueError):
        oa.do_auto_rccsd(odd)


def test_frozen_core_equals_dropping_the_orbital():
    """fcn=1 with canonical orbitals == RCCSD.jl on the Wfn without the lowest occupied orbital."""
    w = canonical_wfn(10, 4, 11)
    r = oa.do_auto_rccsd(w, fcn=1, cc_max_iter=5, cc_e_conv=0.0, cc_max_rms=0.0, return_all=True)
    wf = orc.Wfn(3, 6, w.epsa[1:], w.Cao[:, 1:].copy(), w.Cav, w.ao_eri)
    e, T1, T2 = orc.do_rccsd(wf, maxit=5, return_T=True)
    assert abs(e - r["ecc"]) < 1e-13 and np.abs(T2 - r["T2"]).max() < 1e-13 and r["T1"].shape == (3, 6)


def test_ccsd_energy_is_invariant_to_occupied_and_virtual_rotations():
    e0 = oa.do_auto_rccsd(canonical_wfn(10, 3, 7))
    e1 = oa.do_auto_rccsd(noncanonical_wfn(10, 3, 7))
    assert abs(e0 - e1) < 5e-10


@pytest.mark.parametrize("N,o,seed,mix", [(10, 3, 7, 0.0), (9, 4, 11, 0.02), (12, 2, 5, 0.03)])
def test_factorised_sweep_with_fock_equals_literal_update_amp(N, o, seed, mix):
    """The device algorithm (numpy models: factorised, and sharded with 1 and 2 slabs) against the
    87-term literal update_amp for a non-canonical reference, sweep by sweep."""
    w = noncanonical_wfn(N, o, seed, ov_mix=mix)
    f, V, d, D, fo, fv = oa.auto_setup(w)
    I = fm.unique_integrals(w.ao_eri, w.Cao, w.Cav)
    T1, T2 = f[1] / d, V[2] / D
    t, T = T1.copy(), T2.copy()
    ts, Ts = T1.copy(), T2.copy()
    v = N - o
    vp, _, _ = sm.slab_bounds(v, 2, 0)
    Ip = sm.pad_virtuals(I, v, vp)
    pad2 = lambda x, ax: np.pad(x, [(0, vp - v) if k in ax else (0, 0) for k in range(x.ndim)])
    fp = (f[0], pad2(f[1], (1,)), pad2(f[2], (0, 1)))
    evp = np.concatenate([fv, np.full(vp - v, fv.max() + 1e3)])
    ts, Ts = pad2(ts, (1,)), pad2(Ts, (2, 3))
    for it in range(4):
        T1, T2, r1, r2 = oa.auto_update_amp(T1, T2, f, V, d, D)
        t, T = fm.rccsd_iteration(I, t, T, d, D, fock=f)
        assert np.abs(t - T1).max() < 1e-14 and np.abs(T - T2).max() < 1e-14
        # sharded model: two ranks (threads), all-reduce = sum, all-gather = concatenate
        outs = run_two_slabs(Ip, ts, Ts, fo, evp, vp, fp)
        ts, Ts = outs
        assert np.abs(ts[:, :v] - T1).max() < 1e-14 and np.abs(Ts[:, :, :v, :v] - T2).max() < 1e-14


def run_two_slabs(Ip, t, T, eo, ev, vp, fock):
    """Evaluate sharded_model.sweep for 2 ranks without processes: generator-style comm that is
    replayed until every collective has both contributions."""
    import threading
    nr = 2
    barrier = threading.Barrier(nr)
    box = {}
    res = [None] * nr

    class Comm:
        def __init__(self, r):
            self.r, self.n = r, 0

        def _xchg(self, x):
            key = self.n
            self.n += 1
            box[(key, self.r)] = x
            barrier.wait()
            vals = [box[(key, q)] for q in range(nr)]
            barrier.wait()
            return vals

        def allreduce(self, x):
            vals = self._xchg(x)
            return vals[0] + vals[1]

        def allgather_last(self, x):
            return np.concatenate(self._xchg(x), axis=-1)

    def work(r):
        _, b0, b1 = sm.slab_bounds(vp, nr, r)
        R = sm.rank_integrals(Ip, b0, b1)
        res[r] = sm.sweep(R, t, T, eo, ev, b0, b1, comm=Comm(r), singles=True, fock=fock)

    th = [threading.Thread(target=work, args=(r,)) for r in range(nr)]
    [x.start() for x in th]
    [x.join() for x in th]
    assert np.array_equal(res[0][1], res[1][1])
    return res[0]


@pytest.mark.parametrize("N,o,seed", [(8, 3, 7), (9, 2, 11), (10, 4, 5)])
def test_pT_oracle_against_textbook_formula_and_device_model(N, o, seed):
    g, Cao, Cav, eps = synth.dense_inputs(N, o, seed=seed, scale=1.5 / N)
    w = orc.Wfn(o, N - o, eps, Cao, Cav, g)
    e, T1, T2 = orc.do_rccsd(w, maxit=6, return_T=True)
    ooov, oovv, ovvv = (orc.get_eri(w, s) for s in ("OOOV", "OOVV", "OVVV"))
    fo, fv = eps[:o], eps[o:]
    ref = oa.compute_pT(T1=T1, T2=T2, Vvvvo=ovvv.transpose(3, 1, 2, 0), Vvooo=ooov.transpose(3, 1, 0, 2),
                        Vvovo=oovv.transpose(2, 0, 3, 1), fo=fo, fv=fv)
    assert abs(ref) > 1e-4
    assert abs(pm.pt_energy_allsum(T1, T2, ovvv, ooov, oovv, fo, fv) - ref) < 1e-14
    assert abs(pm.pt_energy(pm.device_tensors(T1, T2, ovvv, ooov, oovv), fo, fv) - ref) < 1e-14


def test_pT_padding_contributes_nothing():
    """The device pads o and v to even counts with zero amplitudes/integrals and far-away energies."""
    N, o = 8, 3
    g, Cao, Cav, eps = synth.dense_inputs(N, o, seed=3, scale=1.5 / N)
    w = orc.Wfn(o, N - o, eps, Cao, Cav, g)
    e, T1, T2 = orc.do_rccsd(w, maxit=4, return_T=True)
    ooov, oovv, ovvv = (orc.get_eri(w, s) for s in ("OOOV", "OOVV", "OVVV"))
    fo, fv = eps[:o], eps[o:]
    ref = pm.pt_energy(pm.device_tensors(T1, T2, ovvv, ooov, oovv), fo, fv)
    po = lambda x, axes_o, axes_v: np.pad(x, [(0, 1) if k in axes_o or k in axes_v else (0, 0) for k in range(x.ndim)])
    Dv = pm.device_tensors(po(T1, (0,), (1,)), po(T2, (0, 1), (2, 3)), po(ovvv, (0,), (1, 2, 3)),
                           po(ooov, (0, 1, 2), (3,)), po(oovv, (0, 1), (2, 3)))
    fop = np.concatenate([fo, [fo.min() - 1e3]])
    fvp = np.concatenate([fv, [fv.max() + 1e3]])
    assert abs(pm.pt_energy(Dv, fop, fvp, nocc=o) - ref) < 1e-15


def test_mrccd_diis_reaches_the_rccd_energy():
    """mRCCD (zero guess, Float32 DIIS vectors) converges to RCCD.jl's fixed point within the
    accuracy its own Float32 storage allows."""
    w = canonical_wfn(10, 3, 7)
    ref = orc.do_rccd(w, maxit=60, guess="mp2")
    r = oa.do_mrccd(w, return_all=True)
    assert r["iterations"] <= 40
    assert abs(r["ecc"] - ref) < 1e-6
    assert r["iterations"] < 40 and r["rms_hist"][-1] < 1e-7
