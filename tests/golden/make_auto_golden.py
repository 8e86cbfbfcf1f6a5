"""Generate tests/golden/auto_n9_o3.npz: Fock matrix, per-sweep energies / residual norms, final
amplitudes, E(T) and the mRCCD result of the section-8f ORACLE (oracle/jues_oracle_auto.py) for a small
seeded non-canonical synthetic input with one frozen core orbital.  The reference repository holds no
vectors for these entry points; these are regression fixtures of the oracle itself and a
run-time-oracle-free target for the GPU path.

    python tests/golden/make_auto_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import jues.jl_b200 as jb                      # noqa: E402  (input generator only)
from oracle import jues_oracle as orc          # noqa: E402
from oracle import jues_oracle_auto as oa      # noqa: E402

N, O, SEED, FCN, MIX = 9, 3, 2024, 1, 0.02
g, h, Ca, eps = jb.synth.noncanonical_inputs(N, O, seed=SEED, ov_mix=MIX)
w = orc.Wfn(O, N - O, eps, Ca[:, :O].copy(), Ca[:, O:].copy(), g, hao=h, Ca=Ca)
r = oa.do_auto_rccsd(w, do_pT=True, fcn=FCN, return_all=True)
r0 = oa.do_auto_rccsd(w, return_all=True)
gc, Cao, Cav, _ = jb.synth.dense_inputs(N, O, seed=SEED)
m = oa.do_mrccd(orc.Wfn(O, N - O, eps, Cao, Cav, gc), return_all=True)
np.savez(os.path.join(os.path.dirname(os.path.abspath(__file__)), "auto_n9_o3.npz"),
         nbf=N, nocc=O, seed=SEED, fcn=FCN, ov_mix=MIX, gao=g, hao=h, Ca=Ca, eps=eps,
         fock=oa.get_fock(w), e_hist_fc=r["e_hist"], rms_hist_fc=r["rms_hist"], T1_fc=r["T1"], T2_fc=r["T2"],
         ept_fc=r["ept"], e_hist=r0["e_hist"], rms_hist=r0["rms_hist"], T1=r0["T1"], T2=r0["T2"],
         mrccd_ecc=m["ecc"], mrccd_iterations=m["iterations"], mrccd_rms0=m["rms_hist"][0])
print("fcn=%d: E_CCSD %.15f in %d sweeps, E(T) %.15f;  fcn=0: E_CCSD %.15f in %d sweeps;  mRCCD %.12f in %d sweeps"
      % (FCN, r["ecc"], r["iterations"], r["ept"], r0["ecc"], r0["iterations"], m["ecc"], m["iterations"]))
