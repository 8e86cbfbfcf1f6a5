"""The A/B switches of the sweep (environment variables read once per process) must all give the reference's
numbers: each is run in a subprocess at a small shape and compared with the CPU oracle, every sweep
(energies 1e-10, final amplitudes 1e-9)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import jues_oracle as orc

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r"""
import json, sys
sys.path.insert(0, %r)
import numpy as np
import jues.jl_b200 as jb
N, o, seed = 24, 5, 2024
g, Cao, Cav, eps = jb.synth.dense_inputs(N, o, seed=seed)
w = jb.Wfn(o, N - o, eps, Cao, Cav, g)
ctx = jb.Context(0)
h1, h2 = [], []
e1, T1, T2 = jb.RCCSD.do_rccsd(w, ctx=ctx, _return_T=True, _e_hist=h1)
e2, T2d = jb.RCCD.do_rccd(w, ctx=ctx, _return_T2=True, _e_hist=h2)
print("RESULT " + json.dumps({"h1": list(h1), "h2": list(h2), "T1": T1.ravel(order="F").tolist(),
                              "T2": T2.ravel(order="F").tolist(), "T2d": T2d.ravel(order="F").tolist()}))
""" % ROOT


@pytest.fixture(scope="module")
def reference():
    import jues.jl_b200 as jb
    N, o = 24, 5
    g, Cao, Cav, eps = jb.synth.dense_inputs(N, o, seed=2024)
    wo = orc.Wfn(o, N - o, eps, Cao, Cav, g)
    r1, r2 = [], []
    e1, T1, T2 = orc.do_rccsd(wo, return_T=True, callback=lambda it, e, a, b: r1.append(e))
    e2, T2d = orc.do_rccd(wo, return_T2=True, callback=lambda it, e, a: r2.append(e))
    return dict(h1=np.array(r1), h2=np.array(r2), T1=T1, T2=T2, T2d=T2d)


@pytest.mark.parametrize("switch", ["JUES_B200_PLAIN_SWEEP", "JUES_B200_NO_AMP_EXTRAS", "JUES_B200_NO_OVERLAP",
                                    "JUES_B200_NO_GRAPH", "JUES_B200_NO_ARENA", "JUES_B200_PLAIN_LADDER",
                                    "JUES_B200_NO_SKINNY"])
def test_switch_keeps_parity(switch, reference):
    env = dict(os.environ)
    env[switch] = "1"
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("RESULT ")][-1]
    d = json.loads(line[len("RESULT "):])
    ref = reference
    assert np.abs(np.array(d["h1"]) - ref["h1"]).max() <= 1e-10, switch
    assert np.abs(np.array(d["h2"]) - ref["h2"]).max() <= 1e-10, switch
    assert np.abs(np.array(d["T1"]) - ref["T1"].ravel(order="F")).max() <= 1e-9
    assert np.abs(np.array(d["T2"]) - ref["T2"].ravel(order="F")).max() <= 1e-9
    assert np.abs(np.array(d["T2d"]) - ref["T2d"].ravel(order="F")).max() <= 1e-9
