"""Generate tests/golden/synthetic_n8_o3.npz: per-sweep energies and amplitudes of the ORACLE
(oracle/jues_oracle.py) for a small seeded synthetic input.  The reference repository holds no
vectors for this path; these are regression fixtures of the oracle itself (so that a change to the
oracle, numpy, or the input generator is noticed), and a size-independent target for the GPU path.

    python tests/golden/make_synthetic_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import jues.jl_b200 as jb                      # noqa: E402  (input generator only)
from oracle import jues_oracle as orc          # noqa: E402

N, O, SEED = 8, 3, 2024
g, Cao, Cav, eps = jb.synth.dense_inputs(N, O, seed=SEED)
w = orc.Wfn(O, N - O, eps, Cao, Cav, g)
sd, d = [], []
e_sd, T1, T2 = orc.do_rccsd(w, return_T=True, callback=lambda it, e, a, b: sd.append(e))
e_d, T2d = orc.do_rccd(w, return_T2=True, callback=lambda it, e, b: d.append(e))
np.savez(os.path.join(os.path.dirname(os.path.abspath(__file__)), "synthetic_n8_o3.npz"),
         nbf=N, nocc=O, seed=SEED, gao=g, Cao=Cao, Cav=Cav, eps=eps,
         e_mp2=orc.do_rmp2(w), e_rccsd_hist=np.array(sd), e_rccd_hist=np.array(d),
         T1=T1, T2=T2, T2_rccd=T2d, oovv=orc.get_eri(w, "OOVV"),
         full_mo=orc.tei_transform(g, np.hstack([Cao, Cav])))
print("E_MP2 %.15f  E_CCSD %.15f  E_CCD %.15f" % (orc.do_rmp2(w), e_sd, e_d))
