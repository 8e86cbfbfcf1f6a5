"""The N>1 path on CPU: the sharded sweep (tests/sharded_model.py == the algorithm of
csrc/cc.cu for nranks > 1) run by 2 and 3 real processes over torch.distributed/gloo, checked
against the literal oracle; plus the host-side slab partition logic."""
import os
import socket

import numpy as np
import pytest

import jues.jl_b200 as jb
from oracle import jues_oracle as orc
import factorized_model as fm
import sharded_model as sm


def test_slab_bounds_cover_and_are_equal():
    for v in (2, 7, 19, 100, 400, 401):
        for P in (1, 2, 3, 4, 8):
            vp, _, _ = sm.slab_bounds(v, P, 0)
            assert vp >= v and vp % (2 * P) == 0 and vp - v < 2 * P
            edges = [sm.slab_bounds(v, P, r)[1:] for r in range(P)]
            assert edges[0][0] == 0 and edges[-1][1] == vp
            assert all(edges[r][1] == edges[r + 1][0] for r in range(P - 1))
            assert len({b - a for a, b in edges}) == 1 and (edges[0][1] - edges[0][0]) % 2 == 0


def test_c_library_uses_the_same_partition():
    """jues.jl_b200.slab_bounds is the host-side mirror used by bench.py / the Julia shim."""
    for v, P, r in [(19, 2, 1), (100, 8, 5), (400, 8, 7)]:
        assert jb.slab_bounds(v, P, r) == sm.slab_bounds(v, P, r)


def _setup(N, o, seed):
    g, Cao, Cav, eps = jb.synth.dense_inputs(N, o, seed=seed)
    return g, Cao, Cav, eps


@pytest.mark.parametrize("singles", [True, False])
def test_sharded_model_single_rank_equals_oracle(singles):
    N, o = 11, 3
    g, Cao, Cav, eps = _setup(N, o, 5)
    v = N - o
    I6 = fm.unique_integrals(g, Cao, Cav)
    vp, b0, b1 = sm.slab_bounds(v, 1, 0)
    R = sm.rank_integrals(sm.pad_virtuals(I6, v, vp), b0, b1)
    eo = eps[:o]
    ev = np.concatenate([eps[o:], np.full(vp - v, eps.max() + 1e3)])
    D = orc.form_Dijab(o, v, eps)
    t = np.zeros((o, vp))
    T = np.zeros((o, o, vp, vp))
    T[:, :, :v, :v] = I6["V"] / D
    if singles:
        I = orc.make_rccsd_integrals(g, Cao, Cav)
        tr, Tr = np.zeros((o, v)), I["oovv"] / D
    else:
        ints = orc.make_rccd_integrals(g, Cao, Cav)
        Tr = I6["V"] / D
    for _ in range(4):
        t, T = sm.sweep(R, t, T, eo, ev, b0, b1, singles=singles)
        if singles:
            tr, Tr = orc.rccsd_iteration(I, tr, Tr, orc.form_Dia(o, v, eps), D)
            assert np.abs(t[:, :v] - tr).max() < 1e-14
        else:
            Tr = orc.rccd_iteration(Tr, ints, D)
        assert np.abs(T[:, :, :v, :v] - Tr).max() < 1e-14
        if vp > v:
            assert np.abs(T[:, :, v:, :]).max() == 0.0 and np.abs(T[:, :, :, v:]).max() == 0.0


def _worker(rank, world, port, N, o, seed, singles, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        class Comm:
            def allreduce(self, x):
                t_ = torch.from_numpy(np.ascontiguousarray(x))
                dist.all_reduce(t_)
                return t_.numpy()

            def allgather_last(self, x):
                xs = np.ascontiguousarray(np.moveaxis(x, -1, 0))       # slab axis first: contiguous concat
                parts = [torch.empty(xs.shape, dtype=torch.float64) for _ in range(world)]
                dist.all_gather(parts, torch.from_numpy(xs))
                return np.moveaxis(np.concatenate([p.numpy() for p in parts], axis=0), 0, -1)

        g, Cao, Cav, eps = _setup(N, o, seed)
        v = N - o
        I6 = fm.unique_integrals(g, Cao, Cav)
        vp, b0, b1 = sm.slab_bounds(v, world, rank)
        R = sm.rank_integrals(sm.pad_virtuals(I6, v, vp), b0, b1)
        eo = eps[:o]
        ev = np.concatenate([eps[o:], np.full(vp - v, eps.max() + 1e3)])
        D = orc.form_Dijab(o, v, eps)
        t = np.zeros((o, vp))
        T = np.zeros((o, o, vp, vp))
        T[:, :, :v, :v] = I6["V"] / D
        for _ in range(3):
            t, T = sm.sweep(R, t, T, eo, ev, b0, b1, comm=Comm(), singles=singles)
        if rank == 0:
            if singles:
                I = orc.make_rccsd_integrals(g, Cao, Cav)
                tr, Tr = np.zeros((o, v)), I["oovv"] / D
                for _ in range(3):
                    tr, Tr = orc.rccsd_iteration(I, tr, Tr, orc.form_Dia(o, v, eps), D)
                err = max(np.abs(t[:, :v] - tr).max(), np.abs(T[:, :, :v, :v] - Tr).max())
            else:
                ints = orc.make_rccd_integrals(g, Cao, Cav)
                Tr = I6["V"] / D
                for _ in range(3):
                    Tr = orc.rccd_iteration(Tr, ints, D)
                err = np.abs(T[:, :, :v, :v] - Tr).max()
            q.put(float(err))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world,singles", [(2, True), (2, False), (3, True)])
def test_sharded_sweep_over_gloo(world, singles):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 11, 3, 7, singles, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    assert q.get(timeout=5) < 1e-13


def _auto_worker(rank, world, port, N, o, seed, q):
    """AutoRCCSD on `world` ranks: sharded sweeps with the replicated one-body Fock terms, then the (T)
    correction with the occupied pairs dealt to the ranks and a scalar all-reduce."""
    import torch
    import torch.distributed as dist
    import pt_model as pm
    from oracle import jues_oracle_auto as oa
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        class Comm:
            def allreduce(self, x):
                t_ = torch.from_numpy(np.ascontiguousarray(x))
                dist.all_reduce(t_)
                return t_.numpy()

            def allgather_last(self, x):
                xs = np.ascontiguousarray(np.moveaxis(x, -1, 0))
                parts = [torch.empty(xs.shape, dtype=torch.float64) for _ in range(world)]
                dist.all_gather(parts, torch.from_numpy(xs))
                return np.moveaxis(np.concatenate([p.numpy() for p in parts], axis=0), 0, -1)

        g, h, Ca, eps = jb.synth.noncanonical_inputs(N, o, seed=seed, ov_mix=0.02)
        v = N - o
        w = orc.Wfn(o, v, eps, Ca[:, :o].copy(), Ca[:, o:].copy(), g, hao=h, Ca=Ca)
        f, V, d, D, fo, fv = oa.auto_setup(w)
        I6 = fm.unique_integrals(g, w.Cao, w.Cav)
        vp, b0, b1 = sm.slab_bounds(v, world, rank)
        R = sm.rank_integrals(sm.pad_virtuals(I6, v, vp), b0, b1)
        pad = lambda x, ax: np.pad(x, [(0, vp - v) if k in ax else (0, 0) for k in range(x.ndim)])
        fp = (f[0], pad(f[1], (1,)), pad(f[2], (0, 1)))
        evp = np.concatenate([fv, np.full(vp - v, fv.max() + 1e3)])
        t, T = pad(f[1] / d, (1,)), pad(V[2] / D, (2, 3))
        T1, T2 = f[1] / d, V[2] / D
        for _ in range(3):
            t, T = sm.sweep(R, t, T, fo, evp, b0, b1, comm=Comm(), singles=True, fock=fp)
            T1, T2, _, _ = oa.auto_update_amp(T1, T2, f, V, d, D)
        err = max(np.abs(t[:, :v] - T1).max(), np.abs(T[:, :, :v, :v] - T2).max())
        # (T): every rank holds the gathered operands; pairs dealt round-robin, scalar all-reduce
        Dv = pm.device_tensors(t, T, pad(V[4], (1, 2, 3)), pad(V[1], (3,)), pad(V[2], (2, 3)))
        share = np.array([pm.pt_energy(Dv, fo, evp, rank=rank, nranks=world)])
        ept = float(Comm().allreduce(share)[0])
        ref = oa.compute_pT(T1=T1, T2=T2, Vvvvo=V[4].transpose(3, 1, 2, 0), Vvooo=V[1].transpose(3, 1, 0, 2),
                            Vvovo=V[2].transpose(2, 0, 3, 1), fo=fo, fv=fv)
        if rank == 0:
            q.put((float(err), abs(ept - ref), abs(ref)))
    finally:
        dist.destroy_process_group()


def test_auto_rccsd_and_triples_over_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    world = 2
    procs = [ctx.Process(target=_auto_worker, args=(r, world, port, 9, 3, 13, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0
    err, dpt, mag = q.get(timeout=5)
    assert err < 1e-13 and dpt < 1e-14 and mag > 1e-8


# ---- the sharded one-pass integral transformation (tests/transform_model.py == csrc/transform.cu) ---------
import transform_model as tm


def _classes_reference(g, Cao, Cavp, b0, b1):
    I6 = fm.unique_integrals(g, Cao, Cavp)
    R = sm.rank_integrals(I6, b0, b1)
    return R


@pytest.mark.parametrize("singles", [True, False])
def test_transform_model_single_rank(singles):
    N, o = 9, 3
    g, Cao, Cav, eps = _setup(N, o, 3)
    R = tm.cc_classes(g, Cao, Cav, tm.SoloComm(), singles=singles)
    ref = _classes_reference(g, Cao, Cav, 0, N - o)
    for k in R:
        assert np.abs(R[k] - ref[k]).max() < 1e-13, k
    w = orc.Wfn(o, N - o, eps, Cao, Cav, g)
    ijab = tm.mp2_slab(g, Cao, Cav, tm.SoloComm())
    assert np.abs(ijab - orc.get_eri(w, "OOVV")).max() < 1e-13


def _transform_worker(rank, world, port, N, o, seed, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        class Comm:
            size = world

            def alltoall(self, parts):
                shapes = [list(p.shape) for p in parts]
                # nu-block extents differ per source: exchange shapes first
                mine = torch.tensor([s[2] for s in shapes[rank:rank + 1]], dtype=torch.int64)
                allnb = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
                dist.all_gather(allnb, mine)
                out = []
                reqs = []
                for d in range(world):
                    shp = list(parts[rank].shape)
                    shp[2] = int(allnb[d][0])
                    out.append(torch.empty(shp, dtype=torch.float64))
                for d in range(world):
                    if d == rank:
                        out[d].copy_(torch.from_numpy(parts[d]))
                    else:
                        reqs.append(dist.isend(torch.from_numpy(parts[d]), d))
                        reqs.append(dist.irecv(out[d], d))
                for r_ in reqs:
                    r_.wait()
                return [x.numpy() for x in out]

            def allgather_last(self, x):
                xs = np.ascontiguousarray(np.moveaxis(x, -1, 0))
                parts = [torch.empty(xs.shape, dtype=torch.float64) for _ in range(world)]
                dist.all_gather(parts, torch.from_numpy(xs))
                return np.moveaxis(np.concatenate([p.numpy() for p in parts], axis=0), 0, -1)
        Comm.rank = rank
        g, Cao, Cav, eps = _setup(N, o, seed)
        v = N - o
        vp, b0, b1 = sm.slab_bounds(v, world, rank)
        Cavp = np.pad(Cav, ((0, 0), (0, vp - v)))
        R = tm.cc_classes(g, Cao, Cavp, Comm(), singles=True)
        ref = _classes_reference(g, Cao, Cavp, b0, b1)
        err = max(float(np.abs(R[k] - ref[k]).max()) for k in R)
        ijab = tm.mp2_slab(g, Cao, Cavp, Comm())
        w = orc.Wfn(o, v, eps, Cao, Cav, g)
        full = np.pad(orc.get_eri(w, "OOVV"), ((0, 0), (0, 0), (0, vp - v), (0, vp - v)))
        err = max(err, float(np.abs(ijab - full[:, :, :, b0:b1]).max()))
        q.put(err)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_transform_over_gloo(world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_transform_worker, args=(r, world, port, 10, 3, 21, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    for _ in range(world):
        assert q.get(timeout=5) < 1e-13
