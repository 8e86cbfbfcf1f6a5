"""1-GPU diagnostic: the <vv|vv> slab transform (last MO index restricted to a column slab of Cav, as a
rank of a multi-GPU run does it) against the slab of the full transform, and against itself (determinism)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import jues.jl_b200 as jb

nbf, nocc, P = [int(x) for x in os.environ.get("DIAG_SHAPE", "144,20,2").split(",")]
seed = 2024
v = nbf - nocc
ctx = jb.Context(0)
Cao, Cav, eps = jb.synth.orbitals(nbf, nocc, seed)
g = jb.DeviceFourTensor.synth_eri(nbf, seed=seed, scale=jb.synth.counter_scale(nbf), ctx=ctx)
full = jb.tei_transform(g, Cav, Cav, Cav, Cav, "f", ctx=ctx)
full_h = full.to_array() if hasattr(full, "to_array") else np.asarray(full)
if hasattr(full, "free"): full.free()
vs = v // P
for r in range(P):
    Cs = np.asfortranarray(Cav[:, r * vs:(r + 1) * vs])
    outs = []
    for rep in range(3):
        t = jb.tei_transform(g, Cav, Cav, Cav, Cs, "s", ctx=ctx)
        h = t.to_array() if hasattr(t, "to_array") else np.asarray(t)
        if hasattr(t, "free"): t.free()
        outs.append(h)
    ref = full_h[:, :, :, r * vs:(r + 1) * vs]
    d = np.abs(outs[0] - ref)
    bad = np.argwhere(d > 1e-13 * np.abs(ref).max())
    rep = {"slab": r, "max_abs_ref": float(np.abs(ref).max()), "max_diff": float(d.max()), "nbad": int(len(bad)),
           "rep_diff": [float(np.abs(outs[0] - outs[k]).max()) for k in (1, 2)]}
    if len(bad):
        for ax in range(4):
            u = np.unique(bad[:, ax])
            rep[f"ax{ax}"] = [int(u.min()), int(u.max()), int(len(u))]
        rep["sample"] = bad[:10].tolist()
    print(json.dumps(rep), flush=True)
