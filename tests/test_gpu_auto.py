"""GPU parity tests of the section-8f entry points (get_fock, AutoRCCSD.do_rccsd with convergence
control / frozen core / non-canonical Fock, compute_pT) through the C ABI against the CPU oracle
(oracle/jues_oracle_auto.py) on identical seeded inputs.

Tolerances (BASELINE.json north_star): energies 1e-10 Eh, amplitudes 1e-9 max-abs PER SWEEP; the
number of sweeps to convergence and the converged flag must be identical."""
import os

import numpy as np
import pytest

import jues.jl_b200 as jb
from oracle import jues_oracle as orc
from oracle import jues_oracle_auto as oa

pytestmark = pytest.mark.gpu

E_TOL = 1e-10
AMP_TOL = 1e-9
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def pair(o, v, eps, Ca, g, h):
    """(product Wfn, oracle Wfn) of the same inputs."""
    kw = dict(hao=h, Ca=Ca)
    return (jb.Wfn(o, v, eps, Ca[:, :o].copy(), Ca[:, o:].copy(), g, **kw),
            orc.Wfn(o, v, eps, Ca[:, :o].copy(), Ca[:, o:].copy(), g, **kw))


def canonical(N, o, seed, scale=None):
    g, Cao, Cav, eps = jb.synth.dense_inputs(N, o, seed=seed, scale=scale)
    h = jb.synth.core_hamiltonian(g, Cao, Cav, eps)
    return pair(o, N - o, eps, np.concatenate([Cao, Cav], axis=1), g, h)


def noncanonical(N, o, seed, ov_mix=0.0):
    g, h, Ca, eps = jb.synth.noncanonical_inputs(N, o, seed=seed, ov_mix=ov_mix)
    return pair(o, N - o, eps, Ca, g, h)


def run_with_capture(ctx, fn):
    cap = []
    ctx.set_amplitude_callback(lambda it, e, T1, T2: cap.append((it, e, T1, T2)))
    try:
        res = fn()
    finally:
        ctx.set_amplitude_callback(None)
    return res, cap


# ------------------------------------------------------------------------------------------
# get_fock
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("N,o,seed,mix", [(7, 5, 1, 0.0), (12, 3, 7, 0.0), (15, 4, 5, 0.05), (24, 5, 2024, 0.02)])
def test_get_fock(ctx, N, o, seed, mix):
    w, wo = noncanonical(N, o, seed, ov_mix=mix)
    f = jb.get_fock(w, ctx=ctx)
    ref = oa.get_fock(wo)
    assert f.shape == (N, N)
    assert np.abs(f - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max())
    fb = jb.get_fock(w, spin="Beta", ctx=ctx)
    assert np.array_equal(f, fb)
    with pytest.raises(jb.JuesError):
        jb.get_fock(w, spin="gamma", ctx=ctx)


def test_get_fock_from_a_device_tensor_and_streamed(ctx):
    w, wo = noncanonical(10, 3, 3, ov_mix=0.03)
    ref = oa.get_fock(wo)
    gd = jb.DeviceFourTensor.from_array(w.ao_eri, ctx=ctx)
    wd = jb.Wfn(w.nalpha, w.nvira, w.epsa, w.Cao, w.Cav, gd, hao=w.hao, Ca=w.Ca)
    assert np.abs(jb.get_fock(wd, ctx=ctx) - ref).max() <= 1e-12
    os.environ["JUES_B200_FORCE_STREAM"] = "1"
    try:
        fs = jb.get_fock(w, ctx=ctx)
    finally:
        del os.environ["JUES_B200_FORCE_STREAM"]
    assert np.abs(fs - ref).max() <= 1e-12


# ------------------------------------------------------------------------------------------
# AutoRCCSD.do_rccsd
# ------------------------------------------------------------------------------------------
def check_every_sweep(ctx, w, wo, **opts):
    ref = []
    r = oa.do_auto_rccsd(wo, return_all=True,
                         callback=lambda it, e, T1, T2: ref.append((it, e, T1.copy(), T2.copy())), **opts)
    got, cap = run_with_capture(ctx, lambda: jb.AutoRCCSD.do_rccsd(w, ctx=ctx, _return_all=True, **opts))
    assert got["iterations"] == r["iterations"] and got["converged"] == r["converged"]
    assert len(cap) == len(ref) == r["iterations"] + 1
    for (it, eg, T1g, T2g), (itr, er, T1r, T2r) in zip(cap, ref):
        assert it == itr
        assert abs(eg - er) <= E_TOL, (it, eg, er)
        assert np.abs(T1g - T1r).max() <= AMP_TOL, it
        assert np.abs(T2g - T2r).max() <= AMP_TOL, it
    assert abs(got["ecc"] - r["ecc"]) <= E_TOL
    assert np.abs(got["e_hist"] - r["e_hist"]).max() <= E_TOL
    assert got["rms_hist"].shape == r["rms_hist"].shape
    if r["iterations"]:
        assert np.abs(got["rms_hist"][1:] - r["rms_hist"][1:]).max() <= 1e-12
    assert np.abs(got["T1"] - r["T1"]).max() <= AMP_TOL and np.abs(got["T2"] - r["T2"]).max() <= AMP_TOL
    return got, r


@pytest.mark.parametrize("N,o,seed", [(7, 5, 1), (12, 3, 7), (24, 5, 2024), (15, 4, 5)])
def test_auto_rccsd_canonical_every_sweep(ctx, N, o, seed):
    w, wo = canonical(N, o, seed)
    got, r = check_every_sweep(ctx, w, wo)
    assert got["converged"]
    # canonical orbitals: the converged energy is RCCSD.jl's
    assert abs(got["ecc"] - orc.do_rccsd(wo)) <= 1e-9


@pytest.mark.parametrize("N,o,seed,mix", [(12, 3, 7, 0.0), (15, 4, 5, 0.0), (9, 4, 11, 0.02), (24, 5, 2024, 0.01),
                                          (13, 5, 3, 0.03)])
def test_auto_rccsd_noncanonical_every_sweep(ctx, N, o, seed, mix):
    """Off-diagonal f_oo, f_vv (and f_ov when mix != 0) enter the amplitude equations."""
    w, wo = noncanonical(N, o, seed, ov_mix=mix)
    check_every_sweep(ctx, w, wo)


def test_auto_rccsd_energy_invariant_to_orbital_rotations(ctx):
    w0, _ = canonical(14, 4, 21)
    w1, _ = noncanonical(14, 4, 21)
    e0 = jb.AutoRCCSD.do_rccsd(w0, ctx=ctx)
    e1 = jb.AutoRCCSD.do_rccsd(w1, ctx=ctx)
    assert abs(e0 - e1) < 5e-10


@pytest.mark.parametrize("fcn", [1, 2])
def test_auto_rccsd_frozen_core(ctx, fcn):
    w, wo = noncanonical(14, 5, 17)
    got, r = check_every_sweep(ctx, w, wo, fcn=fcn)
    assert got["T1"].shape == (5 - fcn, 9)


def test_auto_rccsd_options(ctx):
    w, wo = canonical(12, 3, 7)
    got, r = check_every_sweep(ctx, w, wo, cc_max_iter=3)
    assert got["iterations"] == 3 and not got["converged"]
    got, r = check_every_sweep(ctx, w, wo, cc_e_conv=1e-5, cc_max_rms=1e-6)
    full = jb.AutoRCCSD.do_rccsd(w, ctx=ctx, _return_all=True)
    assert got["converged"] and got["iterations"] < full["iterations"]
    got0, r0 = check_every_sweep(ctx, w, wo, cc_max_iter=0)
    assert got0["iterations"] == 0 and abs(got0["ecc"] - orc.do_rmp2(wo)) <= E_TOL   # the MP2 guess (:250)
    # unknown options are ignored, the scalar return is the correlation energy
    assert abs(jb.AutoRCCSD.do_rccsd(w, ctx=ctx, doprint=True, bogus=3) - full["ecc"]) <= 1e-13


def test_auto_rccsd_argument_errors(ctx):
    w, wo = canonical(10, 3, 7)
    with pytest.raises(jb.JuesError):
        jb.AutoRCCSD.do_rccsd(w, ctx=ctx, fcn=3)
    with pytest.raises(jb.JuesError):
        jb.AutoRCCSD.do_rccsd(w, ctx=ctx, cc_max_iter=-1)
    odd = jb.Wfn(3, 7, w.epsa, w.Cao, w.Cav, w.ao_eri, nbeta=2, hao=w.hao, Ca=w.Ca)
    with pytest.raises(jb.JuesError):
        jb.AutoRCCSD.do_rccsd(odd, ctx=ctx)
    nohao = jb.Wfn(3, 7, w.epsa, w.Cao, w.Cav, w.ao_eri)
    with pytest.raises(jb.JuesError):
        jb.AutoRCCSD.do_rccsd(nohao, ctx=ctx)


def test_auto_rccsd_device_tensor_and_streamed_sources(ctx):
    w, wo = noncanonical(12, 3, 9, ov_mix=0.01)
    r = oa.do_auto_rccsd(wo, return_all=True)
    gd = jb.DeviceFourTensor.from_array(w.ao_eri, ctx=ctx)
    wd = jb.Wfn(w.nalpha, w.nvira, w.epsa, w.Cao, w.Cav, gd, hao=w.hao, Ca=w.Ca)
    a = jb.AutoRCCSD.do_rccsd(wd, ctx=ctx, _return_all=True)
    os.environ["JUES_B200_FORCE_STREAM"] = "1"
    try:
        b = jb.AutoRCCSD.do_rccsd(w, ctx=ctx, _return_all=True)
    finally:
        del os.environ["JUES_B200_FORCE_STREAM"]
    for x in (a, b):
        assert x["iterations"] == r["iterations"]
        assert abs(x["ecc"] - r["ecc"]) <= E_TOL and np.abs(x["T2"] - r["T2"]).max() <= AMP_TOL


def test_auto_rccsd_h2o_sto3g_known_answers(ctx):
    """End to end on the device: the reference's RCCSD constant (test/TestCoupledCluster.jl:44-45)
    through AutoRCCSD's convergence loop, and the CCSD(T) literature value of the same system."""
    d = np.load(os.path.join(GOLD, "h2o_sto3g.npz"))
    C, eps, g, o = d["C"], d["eps"], d["g"], int(d["nocc"])
    w, wo = pair(o, C.shape[1] - o, eps, C, g, d["H"])
    got = jb.AutoRCCSD.do_rccsd(w, ctx=ctx, do_pT=True, _return_all=True)
    ref = oa.do_auto_rccsd(wo, do_pT=True, return_all=True)
    assert got["converged"] and got["iterations"] == ref["iterations"]
    assert abs(got["ecc"] - ref["ecc"]) <= E_TOL and abs(got["ept"] - ref["ept"]) <= E_TOL
    assert abs(got["ecc"] - (-0.070680102078571)) < 2e-9
    assert abs(got["ept"] - (-0.000099877272)) < 2e-9
    e, ept = jb.AutoRCCSD.do_rccsd(w, ctx=ctx, do_pT=True, fcn=1)
    rf = oa.do_auto_rccsd(wo, do_pT=True, fcn=1, return_all=True)
    assert abs(e - rf["ecc"]) <= E_TOL and abs(ept - rf["ept"]) <= E_TOL


# ------------------------------------------------------------------------------------------
# (T)
# ------------------------------------------------------------------------------------------
def pt_inputs(N, o, seed, sweeps=6):
    g, Cao, Cav, eps = jb.synth.dense_inputs(N, o, seed=seed, scale=1.5 / N)
    wo = orc.Wfn(o, N - o, eps, Cao, Cav, g)
    e, T1, T2 = orc.do_rccsd(wo, maxit=sweeps, return_T=True)
    ooov, oovv, ovvv = (orc.get_eri(wo, s) for s in ("OOOV", "OOVV", "OVVV"))
    return dict(T1=T1, T2=T2, Vvvvo=ovvv.transpose(3, 1, 2, 0), Vvooo=ooov.transpose(3, 1, 0, 2),
                Vvovo=oovv.transpose(2, 0, 3, 1), fo=eps[:o].copy(), fv=eps[o:].copy())


@pytest.mark.parametrize("N,o,seed", [(8, 3, 7), (9, 2, 11), (10, 4, 5), (7, 5, 1), (16, 5, 2), (20, 6, 4)])
def test_compute_pT(ctx, N, o, seed):
    kw = pt_inputs(N, o, seed)
    ref = oa.compute_pT(**kw)
    got = jb.compute_pT(ctx=ctx, **kw)
    assert abs(ref) > 1e-6
    assert abs(got - ref) <= E_TOL, (got, ref)


@pytest.mark.parametrize("kb", [1, 2, 3])
def test_compute_pT_chunked_batches(ctx, kb, monkeypatch):
    """Force small batches over the third occupied index (the memory-limited path)."""
    kw = pt_inputs(12, 5, 13)
    ref = oa.compute_pT(**kw)
    monkeypatch.setenv("JUES_B200_PT_KB", str(kb))
    assert abs(jb.compute_pT(ctx=ctx, **kw) - ref) <= E_TOL


def test_compute_pT_is_deterministic_and_checks_shapes(ctx):
    kw = pt_inputs(10, 3, 3)
    a = jb.compute_pT(ctx=ctx, **kw)
    b = jb.compute_pT(ctx=ctx, **kw)
    assert a == b
    bad = dict(kw)
    bad["Vvooo"] = kw["Vvooo"][:, :, :, :-1]
    with pytest.raises(jb.JuesError):
        jb.compute_pT(ctx=ctx, **bad)


@pytest.mark.parametrize("N,o,seed,fcn", [(10, 3, 7, 0), (14, 5, 17, 1), (16, 4, 2, 0)])
def test_auto_rccsd_with_triples(ctx, N, o, seed, fcn):
    w, wo = canonical(N, o, seed, scale=1.0 / N)
    ref = oa.do_auto_rccsd(wo, do_pT=True, fcn=fcn, return_all=True)
    got = jb.AutoRCCSD.do_rccsd(w, ctx=ctx, do_pT=True, fcn=fcn, _return_all=True)
    assert got["iterations"] == ref["iterations"]
    assert abs(got["ecc"] - ref["ecc"]) <= E_TOL
    assert abs(got["ept"] - ref["ept"]) <= E_TOL, (got["ept"], ref["ept"])
    e, ept = jb.AutoRCCSD.do_rccsd(w, ctx=ctx, do_pT=True, fcn=fcn)
    assert e == got["ecc"] and ept == got["ept"]


# ------------------------------------------------------------------------------------------
# mRCCD (DIIS with Float32 vectors)
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("N,o,seed", [(10, 3, 7), (16, 4, 2), (24, 5, 2024)])
def test_mrccd_diis(ctx, N, o, seed):
    """The reference keeps its DIIS vectors -- hence the extrapolated amplitudes -- in Float32
    (mRCCD.jl:64-65,171,200-202), so parity with the CPU path is bounded by Float32 rounding of the
    amplitudes (~6e-8 relative) and of the B matrix: energies within 1e-6 Eh, amplitudes within
    1e-6, same number of sweeps +-1, and the converged energy equals the DIIS-free RCCD one."""
    w, wo = canonical(N, o, seed)
    ref = oa.do_mrccd(wo, return_all=True)
    got = jb.mRCCD.do_rccd(w, ctx=ctx, _return_all=True)
    assert abs(got["iterations"] - ref["iterations"]) <= 1 and got["iterations"] < 40
    assert abs(got["ecc"] - ref["ecc"]) <= 1e-6
    assert np.abs(got["T2"] - ref["T2"]).max() <= 1e-6
    assert got["rms_hist"][-1] < 1e-7
    # first sweep from zero amplitudes is the MP2 amplitudes rounded to Float32
    assert abs(got["e_hist"][0] - orc.do_rmp2(wo)) <= 1e-6
    assert abs(got["rms_hist"][0] - ref["rms_hist"][0]) <= 1e-12
    assert abs(got["ecc"] - orc.do_rccd(wo, maxit=60, guess="mp2")) <= 1e-6
    e, T2 = jb.mRCCD.do_rccd(w, ctx=ctx, return_T2=True)
    assert e == got["ecc"] and np.array_equal(T2, got["T2"])           # deterministic
    assert jb.mRCCD.do_rccd(w, ctx=ctx, maxit=2, _return_all=True)["iterations"] == 2
