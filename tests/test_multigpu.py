"""N>1 on real GPUs: one rank per GPU over NCCL, sharded by virtual slab (needs >= 2 devices;
run with `gpurun --gpus 2 -- python -m pytest tests/test_multigpu.py -m gpu`)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_two_rank_parity():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "tools", "dist_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "DIST_PARITY OK" in r.stdout


@pytest.mark.gpu
def test_single_process_group_drives_two_gpus():
    """jues_b200_init_multi: ONE handle, one host thread per GPU inside the library -- the reference-shaped
    entry points are called exactly as on one GPU (Input.jl:58-71 calls them once from one task)."""
    import numpy as np
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    import jues.jl_b200 as jb
    from oracle import jues_oracle as orc, jues_oracle_auto as oa
    ctx = jb.Context.multi(2)
    try:
        assert ctx.nranks == 2
        for (N, o, seed) in [(12, 3, 7), (31, 6, 11)]:
            g, Cao, Cav, eps = jb.synth.dense_inputs(N, o, seed=seed)
            w = jb.Wfn(o, N - o, eps, Cao, Cav, g)
            wo = orc.Wfn(o, N - o, eps, Cao, Cav, g)
            h = []
            e, T1, T2 = jb.RCCSD.do_rccsd(w, ctx=ctx, _return_T=True, _e_hist=h)
            ref = []
            er, T1r, T2r = orc.do_rccsd(wo, return_T=True, callback=lambda it, ee, a, b: ref.append(ee))
            assert np.abs(np.array(h) - np.array(ref)).max() <= 1e-10
            assert np.abs(T1 - T1r).max() <= 1e-9 and np.abs(T2 - T2r).max() <= 1e-9
            assert abs(jb.do_rmp2(w, ctx=ctx) - orc.do_rmp2(wo)) <= 1e-10
            assert abs(jb.RCCD.do_rccd(w, ctx=ctx) - orc.do_rccd(wo)) <= 1e-10
            assert ctx.comm_counters()["collectives"] > 0            # it really ran sharded
        # the callers either side of the path on the same handle
        g, hc, Ca, eps = jb.synth.noncanonical_inputs(14, 5, seed=17)
        w = jb.Wfn(5, 9, eps, Ca[:, :5].copy(), Ca[:, 5:].copy(), g, hao=hc, Ca=Ca)
        r = jb.AutoRCCSD.do_rccsd(w, ctx=ctx, do_pT=True, fcn=1, _return_all=True)
        wo = orc.Wfn(5, 9, eps, Ca[:, :5].copy(), Ca[:, 5:].copy(), g, hao=hc, Ca=Ca)
        ro = oa.do_auto_rccsd(wo, do_pT=True, fcn=1, return_all=True)
        assert abs(r["ecc"] - ro["ecc"]) <= 1e-10 and abs(r["ept"] - ro["ept"]) <= 1e-10
        assert r["iterations"] == ro["iterations"]
        # an entry point that does not shard runs on the leader's GPU alone
        Cfull = np.hstack([Ca[:, :5], Ca[:, 5:]])
        assert np.abs(jb.tei_transform(g, Cfull, ctx=ctx) - orc.tei_transform(g, Cfull)).max() <= 1e-12
        # generated (storage-less) tensors work on every member; dense device tensors belong to one GPU
        gv = jb.DeviceFourTensor.synth_eri(24, seed=3, ctx=ctx, virtual=True)
        Cao, Cav, eps = jb.synth.orbitals(24, 5, 3)
        e_v = jb.do_rmp2(jb.Wfn(5, 19, eps, Cao, Cav, gv), ctx=ctx)
        e_h = jb.do_rmp2(jb.Wfn(5, 19, eps, Cao, Cav, jb.synth.counter_eri(24, 3)), ctx=ctx)
        assert abs(e_v - e_h) <= 1e-12
        gd = jb.DeviceFourTensor.synth_eri(24, seed=3, ctx=ctx)
        with pytest.raises(jb.JuesError):
            jb.do_rmp2(jb.Wfn(5, 19, eps, Cao, Cav, gd), ctx=ctx)
        gd.free(); gv.free()
    finally:
        ctx.close()
