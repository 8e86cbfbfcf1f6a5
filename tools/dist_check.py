"""Multi-GPU parity check, launched with torchrun (one rank per GPU): the sharded RCCD / RCCSD /
RMP2 through the C ABI on every rank vs the CPU oracle (rank 0) and vs every other rank."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import jues.jl_b200 as jb

def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl")
    ctx = jb.Context(local)
    ctx.init_dist()
    worst = 0.0
    for (N, o, seed) in [(12, 3, 7), (24, 5, 2024), (31, 6, 11)]:
        g, Cao, Cav, eps = jb.synth.dense_inputs(N, o, seed=seed)
        w = jb.Wfn(o, N - o, eps, Cao, Cav, g)
        h1, h2 = [], []
        e_sd, T1, T2 = jb.RCCSD.do_rccsd(w, ctx=ctx, _return_T=True, _e_hist=h1)
        e_d, T2d = jb.RCCD.do_rccd(w, ctx=ctx, _return_T2=True, _e_hist=h2)
        e_mp2 = jb.do_rmp2(w, ctx=ctx)
        # sharded stand-alone transform (last MO index split over ranks, all-gathered)
        tei = jb.tei_transform(g, Cao, Cav, Cav, np.hstack([Cao, Cav]), "x", ctx=ctx)
        tei_p = jb.get_eri(w, "OVOV", ctx=ctx)
        # identical on every rank
        t = torch.tensor([e_sd, e_d, e_mp2, float(np.abs(T2).sum()), float(np.abs(T1).sum())], dtype=torch.float64, device="cuda")
        lo, hi = t.clone(), t.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        assert float((hi - lo).abs().max()) == 0.0, "ranks disagree"
        if rank == 0:
            from oracle import jues_oracle as orc
            wo = orc.Wfn(o, N - o, eps, Cao, Cav, g)
            ref = []
            er, T1r, T2r = orc.do_rccsd(wo, return_T=True, callback=lambda it, e, a, b: ref.append(e))
            errs = [abs(e_sd - er), np.abs(np.array(h1) - np.array(ref)).max(), np.abs(T1 - T1r).max(),
                    np.abs(T2 - T2r).max()]
            refd = []
            edr, T2dr = orc.do_rccd(wo, return_T2=True, callback=lambda it, e, b: refd.append(e))
            errs += [abs(e_d - edr), np.abs(np.array(h2) - np.array(refd)).max(), np.abs(T2d - T2dr).max(),
                     abs(e_mp2 - orc.do_rmp2(wo))]
            terr = max(np.abs(tei - orc.tei_transform(g, Cao, Cav, Cav, np.hstack([Cao, Cav]))).max(),
                       np.abs(tei_p - orc.get_eri(wo, "OVOV")).max()) / np.abs(g).max()
            assert terr < 1e-12, terr
            print(f"N={N} o={o} world={world} errs={['%.1e' % x for x in errs]} comm={ctx.comm_counters()}", flush=True)
            worst = max(worst, max(errs[0], errs[1], errs[4], errs[5], errs[7]) / 1e-10, max(errs[2], errs[3], errs[6]) / 1e-9)
    # section-8f entry points on the same communicator: AutoRCCSD (non-canonical Fock, frozen core,
    # convergence loop) with the (T) correction (pairs dealt to the ranks, scalar all-reduce) and mRCCD
    for (N, o, seed, fcn) in [(14, 5, 17, 1), (16, 4, 2, 0)]:
        g, h, Ca, eps = jb.synth.noncanonical_inputs(N, o, seed=seed)
        w = jb.Wfn(o, N - o, eps, Ca[:, :o].copy(), Ca[:, o:].copy(), g, hao=h, Ca=Ca)
        r = jb.AutoRCCSD.do_rccsd(w, ctx=ctx, do_pT=True, fcn=fcn, _return_all=True)
        gc, Cao, Cav, _ = jb.synth.dense_inputs(N, o, seed=seed)
        wc = jb.Wfn(o, N - o, eps, Cao, Cav, gc)
        m = jb.mRCCD.do_rccd(wc, ctx=ctx, _return_all=True)
        t = torch.tensor([r["ecc"], r["ept"], float(r["iterations"]), float(np.abs(r["T2"]).sum()), m["ecc"]],
                         dtype=torch.float64, device="cuda")
        lo, hi = t.clone(), t.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        assert float((hi - lo).abs().max()) == 0.0, "ranks disagree (auto)"
        if rank == 0:
            from oracle import jues_oracle as orc, jues_oracle_auto as oa
            wo = orc.Wfn(o, N - o, eps, Ca[:, :o].copy(), Ca[:, o:].copy(), g, hao=h, Ca=Ca)
            ro = oa.do_auto_rccsd(wo, do_pT=True, fcn=fcn, return_all=True)
            mo = oa.do_mrccd(orc.Wfn(o, N - o, eps, Cao, Cav, gc), return_all=True)
            errs = [abs(r["ecc"] - ro["ecc"]), abs(r["ept"] - ro["ept"]), np.abs(r["T2"] - ro["T2"]).max(),
                    float(abs(r["iterations"] - ro["iterations"])), abs(m["ecc"] - mo["ecc"])]
            print(f"AUTO N={N} o={o} fcn={fcn} world={world} errs={['%.1e' % x for x in errs]}", flush=True)
            worst = max(worst, errs[0] / 1e-10, errs[1] / 1e-10, errs[2] / 1e-9, errs[3] * 10, errs[4] / 1e-6)
    ok = torch.tensor([1.0 if worst <= 1.0 else 0.0], device="cuda")
    dist.broadcast(ok, src=0)
    if rank == 0:
        print("DIST_PARITY", "OK" if worst <= 1.0 else "FAIL", "worst/tol=%.3g" % worst, flush=True)
    dist.destroy_process_group()
    sys.exit(0 if float(ok[0]) == 1.0 else 1)

if __name__ == "__main__":
    main()
