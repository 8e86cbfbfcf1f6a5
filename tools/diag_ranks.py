"""Multi-rank diagnostic (torchrun, one rank per GPU): the sharded RCCSD sweep against the SAME library
run on one rank (a second, solo context on the same device) sweep by sweep -- amplitudes element by
element -- under several switches (plain / packed ladder, pool threshold, drained collectives).
Used to localise rank-count-dependent deviations; prints which (a, b) blocks of T2 differ first."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import jues.jl_b200 as jb


def run(ctx, w, maxit, capture):
    amps = []
    if capture:
        ctx.set_amplitude_callback(lambda it, e, T1, T2: amps.append((it, e, T1, T2)))
    hist = []
    e = jb.RCCSD.do_rccsd(w, ctx=ctx, _maxit=maxit, _e_hist=hist)
    ctx.set_amplitude_callback(None)
    return e, hist, amps


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl")
    shapes = [tuple(int(x) for x in s.split(",")) for s in os.environ.get("DIAG_SHAPES", "144,20").split(";")]
    nsw = int(os.environ.get("DIAG_SWEEPS", "4"))
    variants = [("default", {}), ("plain_ladder", {"JUES_B200_PLAIN_LADDER": "1"}),
                ("big_1MB", {"JUES_B200_BIG_MB": "1"}), ("sync_comm", {"JUES_B200_SYNC_COMM": "1"})]
    only = os.environ.get("DIAG_VARIANTS")
    if only:
        variants = [v for v in variants if v[0] in only.split(",")]
    for (nbf, nocc) in shapes:
        seed = 2024
        scale = jb.synth.counter_scale(nbf)
        Cao, Cav, eps = jb.synth.orbitals(nbf, nocc, seed)
        solo = jb.Context(local)
        gs = jb.DeviceFourTensor.synth_eri(nbf, seed=seed, scale=scale, ctx=solo)
        ws = jb.Wfn(nocc, nbf - nocc, eps, Cao, Cav, gs)
        e_s, h_s, a_s = run(solo, ws, nsw, True)
        e_s25, h_s25, _ = run(solo, ws, 26, False)
        gs.free(); solo.close()
        if rank == 0:
            print(f"== nbf={nbf} nocc={nocc} solo e_hist[:{nsw+1}]={['%.16f' % x for x in h_s]}", flush=True)
            print(f"   solo e_hist[22:27]={['%.16f' % x for x in h_s25[22:27]]}", flush=True)
        for name, env in variants:
            for k, v_ in env.items():
                os.environ[k] = v_
            ctx = jb.Context(local)
            ctx.init_dist(rank, world)
            gd = jb.DeviceFourTensor.synth_eri(nbf, seed=seed, scale=scale, ctx=ctx)
            wd = jb.Wfn(nocc, nbf - nocc, eps, Cao, Cav, gd)
            e_d, h_d, a_d = run(ctx, wd, nsw, True)
            e25, h25, _ = run(ctx, wd, 26, False)
            e25b, h25b, _ = run(ctx, wd, 26, False)
            for k in env:
                del os.environ[k]
            rep = {"variant": name, "rank": rank, "dE_per_sweep": [float(abs(a - b)) for a, b in zip(h_s, h_d)],
                   "dE_nocb_vs_solo": [float(a - b) for a, b in zip(h25[20:27], h_s25[20:27])],
                   "repeat_max_dE": float(np.abs(np.array(h25) - np.array(h25b)).max())}
            for (it, e, T1, T2), (_, es, T1s, T2s) in zip(a_d, a_s):
                d2 = np.abs(T2 - T2s)
                d1 = np.abs(T1 - T1s).max() if T1 is not None else 0.0
                bad = np.argwhere(d2 > 1e-13)
                info = {"it": it, "dT1": float(d1), "dT2": float(d2.max()), "nbad": int(len(bad))}
                if len(bad):
                    info["i"] = [int(bad[:, 0].min()), int(bad[:, 0].max())]
                    info["j"] = [int(bad[:, 1].min()), int(bad[:, 1].max())]
                    info["a"] = [int(bad[:, 2].min()), int(bad[:, 2].max())]
                    info["b"] = [int(bad[:, 3].min()), int(bad[:, 3].max())]
                    ab = set((int(r[2]), int(r[3])) for r in bad[:20000])
                    info["n_ab_blocks"] = len(ab)
                    info["ab_sample"] = sorted(ab)[:12]
                    # largest deviation and its position
                    k = np.unravel_index(np.argmax(d2), d2.shape)
                    info["argmax"] = [int(x) for x in k]
                rep.setdefault("sweeps", []).append(info)
            print(json.dumps(rep), flush=True)
            gd.free(); ctx.close()
            dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
