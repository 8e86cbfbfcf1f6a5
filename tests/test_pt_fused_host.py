"""CPU check of the index arithmetic of the fused (T) kernel: jues.jl_b200/csrc/pt_fused.h is compiled
for the host (tests/native/pt_fused_host.cpp, g++) and summed over the kernel's thread space, on X
blocks laid out exactly as pt.cu's GEMMs write them, against the numpy model and the oracle."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import pt_model as pm
from jues.jl_b200 import synth
from oracle import jues_oracle as orc
from oracle import jues_oracle_auto as oa

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("native") / "libpt_fused_host.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", os.path.join(HERE, "native", "pt_fused_host.cpp"),
                    "-o", out], check=True)
    L = C.CDLL(out)
    dp = C.POINTER(C.c_double)
    L.pt_fused_host.restype = C.c_double
    L.pt_fused_host.argtypes = [dp] * 5 + [C.c_int] * 6
    return L


def P(a):
    return np.ascontiguousarray(a).ctypes.data_as(C.POINTER(C.c_double))


@pytest.mark.parametrize("N,o,seed", [(8, 3, 7), (10, 4, 5)])
def test_fused_triple_energy_on_device_layout(lib, N, o, seed):
    g, Cao, Cav, eps = synth.dense_inputs(N, o, seed=seed, scale=1.5 / N)
    w = orc.Wfn(o, N - o, eps, Cao, Cav, g)
    e, T1, T2 = orc.do_rccsd(w, maxit=5, return_T=True)
    ooov, oovv, ovvv = (orc.get_eri(w, s) for s in ("OOOV", "OOVV", "OVVV"))
    fo, fv = eps[:o].copy(), eps[o:].copy()
    v = N - o
    ref = oa.compute_pT(T1=T1, T2=T2, Vvvvo=ovvv.transpose(3, 1, 2, 0), Vvooo=ooov.transpose(3, 1, 0, 2),
                        Vvovo=oovv.transpose(2, 0, 3, 1), fo=fo, fv=fv)
    Dv = pm.device_tensors(T1, T2, ovvv, ooov, oovv)
    Vv = np.asfortranarray(Dv["Vv"]).ravel(order="F")
    t1 = np.asfortranarray(T1).ravel(order="F")
    total = 0.0
    for i in range(o):
        for j in range(i + 1):
            for k0, kb in ((0, j + 1),) if j < 2 else ((0, 2), (2, j - 1)):      # also a chunked batch
                if kb <= 0:
                    continue
                F = [pm.x_blocks(Dv, i, 0, j, 0, k0, 1, kb), pm.x_blocks(Dv, i, 0, k0, 1, j, 0, kb),
                     pm.x_blocks(Dv, k0, 1, i, 0, j, 0, kb), pm.x_blocks(Dv, k0, 1, j, 0, i, 0, kb),
                     pm.x_blocks(Dv, j, 0, k0, 1, i, 0, kb), pm.x_blocks(Dv, j, 0, i, 0, k0, 1, kb)]
                X = np.zeros(6 * kb * v ** 3)
                for s, Fs in enumerate(F):                     # Fs[n][a,b,c]
                    if s in (2, 3):                            # [p0,p1,kk,p2]
                        blk = np.asfortranarray(Fs.transpose(1, 2, 0, 3)).ravel(order="F")
                    else:                                      # [kk][p0,p1,p2]
                        blk = np.asfortranarray(Fs.transpose(1, 2, 3, 0)).ravel(order="F")
                    X[s * kb * v ** 3:(s + 1) * kb * v ** 3] = blk
                total += lib.pt_fused_host(P(X), P(Vv), P(t1), P(fo), P(fv), o, v, i, j, k0, kb)
    assert abs(total - ref) < 1e-14, (total, ref)
