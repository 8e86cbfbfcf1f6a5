"""Pin the oracle against the reference's OWN known-answer tests.

The reference holds no golden vectors for the hot path, but its (stale, psi4-backed) tests assert
three correlation energies for H2O/STO-3G (R=1.1 A, 104 deg) and three AO ERIs for H2/STO-3G.
The inputs are rebuilt offline by oracle/sto3g_fixture.py (independent integral + RHF code);
the fixtures are committed under tests/golden/ with that generating script.  Agreement is bounded
by the basis-table digits, bohr constant and SCF convergence of the original psi4 run (~1e-9),
NOT by the oracle's algebra, so the tolerance here is 2e-9 Eh."""
import os

import numpy as np
import pytest

from oracle import jues_oracle as orc

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 2e-9


@pytest.fixture(scope="module")
def h2o():
    d = np.load(os.path.join(GOLD, "h2o_sto3g.npz"))
    C, eps, g, o = d["C"], d["eps"], d["g"], int(d["nocc"])
    return orc.Wfn(o, C.shape[1] - o, eps, C[:, :o].copy(), C[:, o:].copy(), g), d


def test_scf_energy_of_fixture(h2o):
    # Crawford "programming projects" value for this geometry / basis
    assert abs(float(h2o[1]["escf"]) - (-74.942079928192)) < 1e-8


def test_rmp2_known_answer(h2o):
    """test/TestMollerPlesset.jl:35"""
    assert abs(orc.do_rmp2(h2o[0]) - (-0.04914964480386458)) < TOL


def test_rccd_known_answer(h2o):
    """test/TestCoupledCluster.jl:41-42"""
    assert abs(orc.do_rccd(h2o[0]) - (-0.07015050066089029)) < TOL


def test_rccsd_known_answer(h2o):
    """test/TestCoupledCluster.jl:44-45"""
    assert abs(orc.do_rccsd(h2o[0]) - (-0.070680102078571)) < TOL


def test_h2_ao_eris_known_answer():
    """test/TestWavefunction.jl:33-35 (1-based indices there)."""
    g = np.load(os.path.join(GOLD, "h2_sto3g.npz"))["g"]
    assert abs(g[0, 0, 0, 0] - 0.7746059439198979) < 2e-9
    assert abs(g[1, 0, 1, 1] - 0.3093089669634818) < 2e-9
    assert abs(g[0, 0, 1, 1] - 0.4780413730018048) < 2e-9


def test_fixture_regenerates(tmp_path):
    """The committed fixture is what the committed script produces."""
    from oracle import sto3g_fixture as fx
    S, T, V, g, enuc = fx.integrals(fx.h2o_geometry())
    d = np.load(os.path.join(GOLD, "h2o_sto3g.npz"))
    assert np.abs(g - d["g"]).max() < 1e-12
    assert np.abs(T + V - d["H"]).max() < 1e-12


@pytest.mark.gpu
def test_gpu_reproduces_known_answers(ctx, h2o):
    """End to end on the device: the reference's three constants through the CUDA path (C1 of
    BASELINE.json: RCCD on H2O/STO-3G), plus GPU == oracle to 1e-10 on the same inputs."""
    import jues.jl_b200 as jb
    wo = h2o[0]
    w = jb.Wfn(wo.nalpha, wo.nvira, wo.epsa, wo.Cao, wo.Cav, wo.ao_eri)
    for got, ref, const in [(jb.do_rmp2(w, ctx=ctx), orc.do_rmp2(wo), -0.04914964480386458),
                            (jb.RCCD.do_rccd(w, ctx=ctx), orc.do_rccd(wo), -0.07015050066089029),
                            (jb.RCCSD.do_rccsd(w, ctx=ctx), orc.do_rccsd(wo), -0.070680102078571)]:
        assert abs(got - ref) < 1e-10
        assert abs(got - const) < TOL
