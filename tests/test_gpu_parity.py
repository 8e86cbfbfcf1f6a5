"""GPU parity tests: the CUDA path (through the C ABI / the reference-shaped Python API) against
the CPU oracle on identical seeded inputs.

Tolerances (BASELINE.json north_star): correlation energies within 1e-10 Eh, amplitudes within
1e-9 max-abs PER ITERATION; transformed integrals within 1e-12 * max|g|."""
import numpy as np
import pytest

import jues.jl_b200 as jb
from oracle import jues_oracle as orc

pytestmark = pytest.mark.gpu

E_TOL = 1e-10
AMP_TOL = 1e-9


def make_wfns(N, o, seed=2024, scale=None):
    g, Cao, Cav, eps = jb.synth.dense_inputs(N, o, seed=seed, scale=scale)
    return (jb.Wfn(o, N - o, eps, Cao, Cav, g), orc.Wfn(o, N - o, eps, Cao, Cav, g))


# ------------------------------------------------------------------------------------------
# BLAS.gemm! replacement
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(1, 1, 1), (7, 5, 3), (25, 361, 361), (130, 250, 17), (257, 129, 100),
                                   (400, 1000, 333)])
@pytest.mark.parametrize("tA", ["N", "T"])
@pytest.mark.parametrize("tB", ["N", "T"])
def test_gemm(ctx, shape, tA, tB):
    M, N, K = shape
    rng = np.random.default_rng(M * 1000 + N * 10 + K)
    A = rng.standard_normal((K, M) if tA == "T" else (M, K))
    B = rng.standard_normal((N, K) if tB == "T" else (K, N))
    C0 = rng.standard_normal((M, N))
    ref = -0.5 * (A.T if tA == "T" else A) @ (B.T if tB == "T" else B) + 2.0 * C0
    out = ctx.gemm(tA, tB, -0.5, A, B, 2.0, C0.copy(order="F"))
    assert np.abs(out - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max())


NCFG = 10     # tile configurations of csrc/dgemm.cu (kCfgs)


@pytest.mark.parametrize("cfg", range(NCFG))
@pytest.mark.parametrize("split", [1, 3])
def test_gemm_every_tile_configuration(ctx, cfg, split, monkeypatch):
    """Force each tile configuration (and split-K) of the DGEMM on ragged shapes, all transposes."""
    monkeypatch.setenv("JUES_B200_GEMM_CFG", str(cfg + NCFG * (split - 1)))
    rng = np.random.default_rng(cfg)
    for (M, N, K) in [(131, 257, 100), (400, 300, 77), (30, 17, 200)]:
        for tA in "NT":
            for tB in "NT":
                A = rng.standard_normal((K, M) if tA == "T" else (M, K))
                B = rng.standard_normal((N, K) if tB == "T" else (K, N))
                ref = (A.T if tA == "T" else A) @ (B.T if tB == "T" else B)
                out = ctx.gemm(tA, tB, 1.0, A, B)
                assert np.abs(out - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max()), (cfg, split, M, N, K, tA, tB)
                # accumulating epilogue (old values of C fetched in batches)
                C0 = rng.standard_normal((M, N))
                out = ctx.gemm(tA, tB, 0.5, A, B, -1.5, C0.copy(order="F"))
                ref2 = 0.5 * ref - 1.5 * C0
                assert np.abs(out - ref2).max() <= 1e-12 * max(1.0, np.abs(ref2).max()), (cfg, split, M, N, K, tA, tB, "beta")


def test_gemm_skinny_splitk(ctx):
    """o x o output with a very long K (the Fmi / Fae shape) takes the split-K path."""
    rng = np.random.default_rng(5)
    A = rng.standard_normal((20, 40000))
    B = rng.standard_normal((40000, 20))
    out = ctx.gemm("N", "N", 1.0, A, B)
    ref = A @ B
    assert np.abs(out - ref).max() <= 1e-11 * np.abs(ref).max()
    out2 = ctx.gemm("N", "N", 1.0, A, B)
    assert np.array_equal(out, out2), "split-K reduction must be deterministic"


def test_gemm_rejects_bad_arguments(ctx):
    with pytest.raises(jb.JuesError):
        ctx.gemm("N", "N", 1.0, np.zeros((3, 4)), np.zeros((5, 2)))
    with pytest.raises(jb.JuesError):
        ctx.gemm("X", "N", 1.0, np.zeros((3, 4)), np.zeros((4, 2)))


# ------------------------------------------------------------------------------------------
# tei_transform / get_eri
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("N,o", [(7, 5), (10, 3), (24, 5)])
def test_tei_transform_full_one_C(ctx, N, o):
    """tei_transform(gao, C) (Transformation.jl:15-20): BASELINE config 2 at N=24."""
    w, wo = make_wfns(N, o, seed=11)
    Cfull = np.hstack([w.Cao, w.Cav])
    got = jb.tei_transform(w.ao_eri, Cfull, "default", ctx=ctx)
    ref = orc.tei_transform(w.ao_eri, Cfull)
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() <= 1e-12 * np.abs(w.ao_eri).max()


@pytest.mark.parametrize("slots", ["ovov", "oovv", "vvvv", "ovvv", "vooo", "ovvo"])
def test_tei_transform_subsets(ctx, slots):
    w, wo = make_wfns(13, 4, seed=3)
    Cs = [w.Cao if s == "o" else w.Cav for s in slots]
    got = jb.tei_transform(w.ao_eri, *Cs, "x", ctx=ctx)
    ref = orc.tei_transform(w.ao_eri, *Cs)
    assert np.abs(got - ref).max() <= 1e-12 * np.abs(w.ao_eri).max()


@pytest.mark.parametrize("string", ["OOVV", "OVOV", "VVVV", "OOOV", "oOvV"])
@pytest.mark.parametrize("notation", ["phys", "chem"])
def test_get_eri(ctx, string, notation):
    w, wo = make_wfns(12, 3, seed=5)
    got = jb.get_eri(w, string, notation=notation, ctx=ctx)
    ref = orc.get_eri(wo, string, notation=notation)
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() <= 1e-12 * np.abs(w.ao_eri).max()


def test_get_eri_frozen_core_and_errors(ctx):
    w, wo = make_wfns(12, 4, seed=6)
    got = jb.get_eri(w, "OOVV", fcn=1, ctx=ctx)
    ref = orc.get_eri(wo, "OOVV", fcn=1)
    assert got.shape == (3, 3, 8, 8)
    assert np.abs(got - ref).max() <= 1e-12 * np.abs(w.ao_eri).max()
    with pytest.raises(jb.JuesError):
        jb.get_eri(w, "OOV", ctx=ctx)          # IntegralTransformation.jl:41-43
    with pytest.raises(jb.JuesError):
        jb.get_eri(w, "OOXV", ctx=ctx)


# ------------------------------------------------------------------------------------------
# RMP2
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("N,o,seed", [(7, 5, 1), (24, 5, 7), (24, 5, 2024), (31, 6, 11), (60, 10, 2024)])
def test_rmp2(ctx, N, o, seed):
    w, wo = make_wfns(N, o, seed=seed)
    e = jb.do_rmp2(w, ctx=ctx, doprint=False)
    ref = orc.do_rmp2(wo)
    assert abs(e - ref) <= E_TOL, (e, ref)


# ------------------------------------------------------------------------------------------
# RCCD / RCCSD: every sweep
# ------------------------------------------------------------------------------------------
def run_with_capture(ctx, fn):
    cap = []
    ctx.set_amplitude_callback(lambda it, e, T1, T2: cap.append((it, e, T1, T2)))
    try:
        res = fn()
    finally:
        ctx.set_amplitude_callback(None)
    return res, cap


@pytest.mark.parametrize("N,o,seed,guess", [(7, 5, 1, "reference"), (12, 3, 7, "reference"),
                                            (24, 5, 2024, "reference"), (24, 5, 11, "mp2"),
                                            (15, 4, 5, "reference")])
def test_rccd_every_iteration(ctx, N, o, seed, guess):
    w, wo = make_wfns(N, o, seed=seed)
    ref = []
    orc.do_rccd(wo, guess=guess, callback=lambda it, e, T2: ref.append((it, e, T2.copy())))
    hist = []
    e, cap = run_with_capture(ctx, lambda: jb.RCCD.do_rccd(w, ctx=ctx, _guess=guess, _e_hist=hist))
    assert len(cap) == 41 and len(ref) == 41 and len(hist) == 41     # guess + 40 sweeps (RCCD.jl:34)
    for (it, eg, _, T2g), (itr, er, T2r) in zip(cap, ref):
        assert it == itr
        assert abs(eg - er) <= E_TOL, (it, eg, er)
        assert abs(hist[it] - er) <= E_TOL
        assert np.abs(T2g - T2r).max() <= AMP_TOL, it
    assert abs(e - ref[-1][1]) <= E_TOL


@pytest.mark.parametrize("N,o,seed", [(7, 5, 1), (12, 3, 7), (24, 5, 2024), (15, 4, 5), (40, 8, 11)])
def test_rccsd_every_iteration(ctx, N, o, seed):
    w, wo = make_wfns(N, o, seed=seed)
    ref = []
    orc.do_rccsd(wo, callback=lambda it, e, T1, T2: ref.append((it, e, T1.copy(), T2.copy())))
    hist = []
    e, cap = run_with_capture(ctx, lambda: jb.RCCSD.do_rccsd(w, ctx=ctx, _e_hist=hist, maxit=3))
    assert len(cap) == 41 and len(ref) == 41     # public kwargs (maxit=3) are ignored: RCCSD.jl:36-50
    for (it, eg, T1g, T2g), (itr, er, T1r, T2r) in zip(cap, ref):
        assert abs(eg - er) <= E_TOL, (it, eg, er)
        assert np.abs(T1g - T1r).max() <= AMP_TOL, it
        assert np.abs(T2g - T2r).max() <= AMP_TOL, it
    assert abs(e - ref[-1][1]) <= E_TOL


def test_rccsd_returns_amplitudes(ctx):
    w, wo = make_wfns(12, 3, seed=9)
    e, T1, T2 = jb.RCCSD.do_rccsd(w, ctx=ctx, _return_T=True)
    er, T1r, T2r = orc.do_rccsd(wo, return_T=True)
    assert abs(e - er) <= E_TOL
    assert np.abs(T1 - T1r).max() <= AMP_TOL and np.abs(T2 - T2r).max() <= AMP_TOL
    # T2[i,j,a,b] == T2[j,i,b,a]
    assert np.abs(T2 - T2.transpose(1, 0, 3, 2)).max() <= 1e-14


def test_c3_shape_rccsd_energy_trace(ctx):
    """BASELINE config 3 (N=120, o=20) is too slow for the oracle's literal algorithm in a unit
    test; check a mid-size shape end to end plus size-independent properties."""
    w, wo = make_wfns(60, 10, seed=2024)
    hist = []
    e = jb.RCCSD.do_rccsd(w, ctx=ctx, _e_hist=hist)
    ref = []
    orc.do_rccsd(wo, callback=lambda it, ee, T1, T2: ref.append(ee))
    assert np.abs(np.array(hist) - np.array(ref)).max() <= E_TOL
    # the sweep is a fixed-point iteration: late energies stop moving
    assert abs(hist[-1] - hist[-2]) < 1e-12


# ------------------------------------------------------------------------------------------
# device-resident tensors (DiskFourTensor replacement) and the device-resident entry points
# ------------------------------------------------------------------------------------------
def test_device_four_tensor_slices(ctx):
    rng = np.random.default_rng(0)
    a = rng.standard_normal((5, 4, 3, 7))
    t = jb.DeviceFourTensor.from_array(a, ctx=ctx)
    assert t.eltype is np.float64 and t.shape == (5, 4, 3, 7)
    assert np.array_equal(t[:, :, :, :], a)
    assert np.array_equal(t[1:4, :, 2, 3:7], a[1:4, :, 2, 3:7])
    assert t[4, 3, 2, 6] == a[4, 3, 2, 6]
    blk = rng.standard_normal((2, 4, 2))
    t[0:2, :, 1, 5:7] = blk
    a[0:2, :, 1, 5:7] = blk
    assert np.array_equal(t.to_array(), a)
    t[:, 1, :, 2] = 0.25
    a[:, 1, :, 2] = 0.25
    assert np.array_equal(t.to_array(), a)
    t.blockfill(1.5)                                   # blockfill! (DiskFourTensors.jl:88-95)
    assert np.array_equal(t.to_array(), np.full_like(a, 1.5))
    with pytest.raises(jb.JuesError):
        t[0:9, :, :, :]
    t.free()


def test_device_resident_entry_points_match_host_entry_points(ctx):
    w, wo = make_wfns(12, 3, seed=21)
    gdev = jb.DeviceFourTensor.from_array(w.ao_eri, ctx=ctx)
    wd = jb.Wfn(w.nalpha, w.nvira, w.epsa, w.Cao, w.Cav, gdev)
    assert jb.do_rmp2(wd, ctx=ctx) == jb.do_rmp2(w, ctx=ctx)
    assert jb.RCCD.do_rccd(wd, ctx=ctx) == jb.RCCD.do_rccd(w, ctx=ctx)
    assert jb.RCCSD.do_rccsd(wd, ctx=ctx) == jb.RCCSD.do_rccsd(w, ctx=ctx)
    out = jb.tei_transform(gdev, w.Cao, w.Cav, w.Cao, w.Cav, "oovv", ctx=ctx)
    assert isinstance(out, jb.DeviceFourTensor)
    ref = orc.tei_transform(w.ao_eri, w.Cao, w.Cav, w.Cao, w.Cav)
    assert np.abs(out.to_array() - ref).max() <= 1e-12 * np.abs(w.ao_eri).max()


@pytest.mark.parametrize("N", [6, 9])
def test_counter_eri_device_equals_host(ctx, N):
    """The counter-based generator is bit-identical on host and device and 8-fold symmetric."""
    t = jb.DeviceFourTensor.synth_eri(N, seed=77, scale=0.03, ctx=ctx)
    g = t.to_array()
    h = jb.synth.counter_eri(N, seed=77, scale=0.03)
    assert np.array_equal(g, h)
    for perm in [(1, 0, 2, 3), (0, 1, 3, 2), (2, 3, 0, 1), (3, 2, 1, 0)]:
        assert np.array_equal(g, g.transpose(perm))


def test_streamed_transform_equals_resident(ctx):
    """MP2 from a sigma-streamed AO tensor (the out-of-core path of Transformation.jl:94-192)."""
    import os
    w, wo = make_wfns(16, 4, seed=4)
    e1 = jb.do_rmp2(w, ctx=ctx)
    os.environ["JUES_B200_FORCE_STREAM"] = "1"
    try:
        e2 = jb.do_rmp2(w, ctx=ctx)
    finally:
        del os.environ["JUES_B200_FORCE_STREAM"]
    assert abs(e1 - e2) <= 1e-13
    assert abs(e1 - orc.do_rmp2(wo)) <= E_TOL


def test_virtual_synthetic_tensor_streams_like_the_dense_one(ctx):
    """A storage-less synthetic gao (sigma slabs generated on demand, the route for nbf whose
    N^4 does not fit) gives the same answers as the dense tensor, and matches the oracle."""
    N, o = 14, 4
    Cao, Cav, eps = jb.synth.orbitals(N, o, 5)
    sc = jb.synth.counter_scale(N)
    gv = jb.DeviceFourTensor.synth_eri(N, seed=5, scale=sc, ctx=ctx, virtual=True)
    gd = jb.DeviceFourTensor.synth_eri(N, seed=5, scale=sc, ctx=ctx)
    gh = jb.synth.counter_eri(N, 5, sc)
    assert np.array_equal(gv[2:9, :, 3, 1:12], gh[2:9, :, 3, 1:12])
    with pytest.raises(jb.JuesError):
        gv[0, 0, 0, 0] = 1.0
    wv = jb.Wfn(o, N - o, eps, Cao, Cav, gv)
    wd = jb.Wfn(o, N - o, eps, Cao, Cav, gd)
    wo = orc.Wfn(o, N - o, eps, Cao, Cav, gh)
    import os
    os.environ["JUES_B200_FORCE_STREAM"] = "1"       # several sigma slabs even at this size
    try:
        e_v = jb.RCCSD.do_rccsd(wv, ctx=ctx)
        m_v = jb.do_rmp2(wv, ctx=ctx)
        d_v = jb.RCCD.do_rccd(wv, ctx=ctx)
    finally:
        del os.environ["JUES_B200_FORCE_STREAM"]
    assert abs(e_v - jb.RCCSD.do_rccsd(wd, ctx=ctx)) <= 1e-12
    assert abs(e_v - orc.do_rccsd(wo)) <= E_TOL
    assert abs(m_v - orc.do_rmp2(wo)) <= E_TOL
    assert abs(d_v - orc.do_rccd(wo)) <= E_TOL


# ---- determinism of the transform at the shape that exposed the round-1 release race ----------------
def test_slab_transform_deterministic(ctx):
    """Every quarter of the <vv|vv> slab transform (dims v,v,v,v/2) repeated inside one call must
    reproduce the first run bit for bit (profiles/gemm_release_race_r02.md)."""
    import ctypes as C
    from jues.jl_b200 import _p, _f
    nbf, nocc = 72, 10
    v = nbf - nocc
    Cao, Cav, eps = jb.synth.orbitals(nbf, nocc, 5)
    g = jb.DeviceFourTensor.synth_eri(nbf, seed=5, scale=jb.synth.counter_scale(nbf), ctx=ctx)
    reps = 4
    try:
        for r in range(2):
            Cs = np.asfortranarray(Cav[:, r * (v // 2):(r + 1) * (v // 2)])
            st = (C.c_double * (16 * reps))()
            ctx._check(ctx._lib.jues_b200_transform_stress(ctx._h, g._h, _p(_f(Cav)), v, _p(_f(Cav)), v, _p(_f(Cav)), v,
                                                           _p(Cs), v // 2, reps, st))
            a = np.array(list(st)).reshape(reps, 4, 4)
            assert a[:, :, 0].max() == 0.0, a[:, :, 0]
    finally:
        g.free()


def test_gemm_repeatable(ctx):
    for (tA, tB, M, N, K, b) in [("N", "N", 40000, 62, 144, 1), ("T", "N", 124, 30000, 144, 1),
                                 ("N", "N", 124, 124, 144, 300), ("T", "N", 700, 700, 700, 1)]:
        nb, w = ctx.gemm_stress(tA, tB, M, N, K, b, reps=5)
        assert nb == 0 and w == 0.0, (tA, tB, M, N, K, b, nb, w)


# ---- parity at the BASELINE shapes --------------------------------------------------------------------
def test_c3_against_committed_oracle_trace(ctx):
    """BASELINE config 3 (nbf=120, nocc=20), the bench workload: every one of the 40 sweeps against the
    committed trace of the literal oracle (tests/golden/make_bench_golden.py): energy, ||T1||, ||T2|| per
    sweep and 64 sampled elements of the final T2."""
    import os
    p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "bench_ehist_nbf120_nocc20.npz")
    gold = np.load(p)
    nbf, nocc = int(gold["nbf"]), int(gold["nocc"])
    Cao, Cav, eps = jb.synth.orbitals(nbf, nocc, int(gold["seed"]))
    g = jb.DeviceFourTensor.synth_eri(nbf, seed=int(gold["seed"]), scale=jb.synth.counter_scale(nbf), ctx=ctx)
    w = jb.Wfn(nocc, nbf - nocc, eps, Cao, Cav, g)
    rec = []
    ctx.set_amplitude_callback(lambda it, e, T1, T2: rec.append((e, float(np.linalg.norm(T1)), float(np.linalg.norm(T2)),
                                                                  T2 if it == 40 else None, T1 if it == 40 else None)))
    try:
        e = jb.RCCSD.do_rccsd(w, ctx=ctx)
    finally:
        ctx.set_amplitude_callback(None)
        g.free()
    assert len(rec) == 41
    assert np.abs(np.array([r[0] for r in rec]) - gold["e_hist"]).max() <= E_TOL
    assert abs(e - gold["e_hist"][40]) <= E_TOL
    # norms of ~6e6 amplitudes: 1e-9 per element bounds the norm difference by 1e-9 * sqrt(n)
    assert np.abs(np.array([r[1] for r in rec]) - gold["t1_norm"]).max() <= AMP_TOL * np.sqrt(nocc * (nbf - nocc))
    assert np.abs(np.array([r[2] for r in rec]) - gold["t2_norm"]).max() <= AMP_TOL * nocc * (nbf - nocc)
    T2, T1 = rec[40][3], rec[40][4]
    idx = gold["t2_idx"]
    assert np.abs(T2[idx[:, 0], idx[:, 1], idx[:, 2], idx[:, 3]] - gold["t2_samples"]).max() <= AMP_TOL
    assert np.abs(T1 - gold["T1"]).max() <= AMP_TOL


def test_block_streamed_paths_at_n72(ctx):
    """The sub-block pipeline of the one-pass transform above toy sizes (nbf=72, several sub-blocks forced):
    host array, dense device tensor and the storage-less generated tensor give the same RMP2 energy as the
    oracle and the same RCCSD energies as each other."""
    import os
    N, o = 72, 10
    sc = jb.synth.counter_scale(N)
    Cao, Cav, eps = jb.synth.orbitals(N, o, 9)
    gh = jb.synth.counter_eri(N, 9, sc)
    wo = orc.Wfn(o, N - o, eps, Cao, Cav, gh)
    e_ref = orc.do_rmp2(wo)
    gd = jb.DeviceFourTensor.synth_eri(N, seed=9, scale=sc, ctx=ctx)
    gv = jb.DeviceFourTensor.synth_eri(N, seed=9, scale=sc, ctx=ctx, virtual=True)
    os.environ["JUES_B200_FORCE_STREAM"] = "1"
    try:
        res = {}
        for name, g in (("host", gh), ("dense", gd), ("virtual", gv)):
            w = jb.Wfn(o, N - o, eps, Cao, Cav, g)
            h = []
            res[name] = (jb.do_rmp2(w, ctx=ctx), jb.RCCSD.do_rccsd(w, ctx=ctx, _maxit=3, _e_hist=h), h)
    finally:
        del os.environ["JUES_B200_FORCE_STREAM"]
        gd.free(); gv.free()
    for name, (e2, ecc, h) in res.items():
        assert abs(e2 - e_ref) <= E_TOL, (name, e2, e_ref)
        assert np.abs(np.array(h) - np.array(res["host"][2])).max() <= 1e-12, name


def test_sampled_elements_of_a_large_generated_transform(ctx):
    """SURVEY 8d parity protocol for shapes the oracle cannot transform: <ij|ab> of the storage-less generated
    AO tensor at nbf=128 (2.7e8 AO elements never held anywhere), sampled (i, j) pairs checked against host
    quarter transforms of synth.counter_eri_element."""
    N, o = 128, 6
    v = N - o
    sc = jb.synth.counter_scale(N)
    Cao, Cav, eps = jb.synth.orbitals(N, o, 3)
    gv = jb.DeviceFourTensor.synth_eri(N, seed=3, scale=sc, ctx=ctx, virtual=True)
    try:
        w = jb.Wfn(o, v, eps, Cao, Cav, gv)
        ijab = jb.get_eri(w, "OOVV", ctx=ctx)
        ijab = ijab.to_array() if hasattr(ijab, "to_array") else np.asarray(ijab)
    finally:
        gv.free()
    pairs = [(0, 0), (2, 5), (5, 1)]
    # h[nu, sig] = sum_{mu lam} C[mu,i] C[lam,j] g[mu,nu,lam,sig], one sigma plane at a time
    h = np.zeros((len(pairs), N, N))
    for s_ in range(N):
        plane = jb.synth.counter_eri(N, 3, sc, sig_range=(s_, s_ + 1))[:, :, :, 0]
        for k, (i, j) in enumerate(pairs):
            h[k, :, s_] = np.einsum("m,mnl,l->n", Cao[:, i], plane, Cao[:, j], optimize=True)
    for k, (i, j) in enumerate(pairs):
        ref = Cav.T @ h[k] @ Cav                               # <ij|ab> = (ia|jb)
        assert np.abs(ijab[i, j] - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max()) + 1e-13, (i, j)


# ---- the streaming kernels for skinny products (csrc/skinny.cu) ---------------------------------------
@pytest.mark.parametrize("tA,tB,M,N,K", [
    ("T", "N", 3000, 1, 2000), ("T", "T", 2500, 1, 64), ("T", "N", 257, 1, 4097), ("T", "N", 10000, 1, 200),
    ("N", "T", 20, 100, 9000), ("N", "T", 20, 20, 10000), ("N", "T", 100, 100, 8200), ("N", "T", 6, 34, 8500),
    ("N", "T", 128, 128, 8192), ("N", "T", 2, 2, 20001),
    # shapes next to theirs that stay on the tile kernel
    ("N", "N", 4100, 20, 100), ("N", "N", 20, 4100, 100), ("T", "N", 3, 5000, 40), ("N", "T", 7, 33, 8500)])
@pytest.mark.parametrize("beta", [0.0, 0.3])
def test_gemm_skinny_streaming(ctx, tA, tB, M, N, K, beta):
    rng = np.random.default_rng(M + 7 * N + 13 * K)
    A = rng.standard_normal((K, M) if tA == "T" else (M, K))
    B = rng.standard_normal((N, K) if tB == "T" else (K, N))
    C0 = rng.standard_normal((M, N))
    ref = 1.7 * (A.T if tA == "T" else A) @ (B.T if tB == "T" else B) + beta * C0
    got = ctx.gemm(tA, tB, 1.7, A, B, beta, np.asfortranarray(C0.copy()))
    assert np.abs(got - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max()) * np.sqrt(K)
    assert ctx.gemm_stress(tA, tB, M, N, K, 1, reps=3) == (0, 0.0)
