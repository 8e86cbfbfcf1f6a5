"""
Offline STO-3G integral + RHF fixture generator  --  TEST INFRASTRUCTURE (oracle pinning).

The reference obtains its AO integrals from libint2 (via Lints, RHF.jl:43-57,82-88) or psi4
(stale tests); neither exists in this image.  This file is an independent s/p Gaussian integral
code (McMurchie-Davidson) plus a tightly converged RHF, enough to rebuild the inputs of the
reference's own known-answer tests so that the ORACLE can be pinned to the reference's constants:

    H2O, R(OH)=1.1 A, angle 104 deg, STO-3G    test/TestCoupledCluster.jl:11-16
        RMP2  -0.04914964480386458              test/TestMollerPlesset.jl:35
        RCCD  -0.07015050066089029              test/TestCoupledCluster.jl:41-42
        RCCSD -0.070680102078571                test/TestCoupledCluster.jl:44-45
    H2, R=1.0 A, STO-3G AO ERIs                 test/TestWavefunction.jl:33-35
        g[1,1,1,1]=0.7746059439198979  g[2,1,2,2]=0.3093089669634818  g[1,1,2,2]=0.4780413730018048

Agreement is limited by basis-set table digits / physical constants / SCF convergence of the
original psi4 run (~1e-8 Eh), not by the oracle: see tests/test_oracle_known_answers.py.

Run `python oracle/sto3g_fixture.py` to regenerate tests/golden/h2o_sto3g.npz and h2_sto3g.npz.
"""
from __future__ import annotations

import itertools
import math
import os

import numpy as np
from scipy.special import hyp1f1

ANGSTROM = 1.0 / 0.52917721067     # psi4 1.3-era CODATA 2014 bohr radius

# STO-3G (EMSL / psi4 sto-3g.gbs digits)
_S3 = (0.1543289673, 0.5353281423, 0.4446345422)
STO3G = {
    "H": [("s", (3.425250914, 0.6239137298, 0.1688554040), _S3)],
    "O": [("s", (130.7093214, 23.80886605, 6.443608313), _S3),
          ("s", (5.033151319, 1.169596125, 0.3803889600), (-0.09996722919, 0.3995128261, 0.7001154689)),
          ("p", (5.033151319, 1.169596125, 0.3803889600), (0.1559162750, 0.6076837186, 0.3919573931))],
}
Z = {"H": 1, "O": 8}


def _fact2(n):
    return 1 if n <= 0 else n * _fact2(n - 2)


class BasisFunction:
    def __init__(self, origin, shell, exps, coefs):
        self.origin = np.asarray(origin, float)
        self.shell = shell                      # (l, m, n)
        self.exps = exps
        l, m, n = shell
        L = l + m + n
        norm = [math.sqrt(2 ** (2 * L + 1.5) * a ** (L + 1.5)
                          / (_fact2(2 * l - 1) * _fact2(2 * m - 1) * _fact2(2 * n - 1) * math.pi ** 1.5))
                for a in exps]
        # normalise the contraction
        pref = math.pi ** 1.5 * _fact2(2 * l - 1) * _fact2(2 * m - 1) * _fact2(2 * n - 1) / 2.0 ** L
        N = 0.0
        for ia, ca in enumerate(coefs):
            for ib, cb in enumerate(coefs):
                N += norm[ia] * norm[ib] * ca * cb / (exps[ia] + exps[ib]) ** (L + 1.5)
        N = (N * pref) ** -0.5
        self.coefs = [N * c * nn for c, nn in zip(coefs, norm)]


def build_basis(atoms):
    bfs = []
    for sym, xyz in atoms:
        for kind, exps, coefs in STO3G[sym]:
            shells = [(0, 0, 0)] if kind == "s" else [(1, 0, 0), (0, 1, 0), (0, 0, 1)]
            for sh in shells:
                bfs.append(BasisFunction(xyz, sh, exps, coefs))
    return bfs


def E(i, j, t, Qx, a, b):
    """Hermite expansion coefficients."""
    p = a + b
    q = a * b / p
    if t < 0 or t > i + j:
        return 0.0
    if i == j == t == 0:
        return math.exp(-q * Qx * Qx)
    if j == 0:
        return (E(i - 1, j, t - 1, Qx, a, b) / (2 * p) - q * Qx / a * E(i - 1, j, t, Qx, a, b)
                + (t + 1) * E(i - 1, j, t + 1, Qx, a, b))
    return (E(i, j - 1, t - 1, Qx, a, b) / (2 * p) + q * Qx / b * E(i, j - 1, t, Qx, a, b)
            + (t + 1) * E(i, j - 1, t + 1, Qx, a, b))


def boys(n, x):
    return hyp1f1(n + 0.5, n + 1.5, -x) / (2.0 * n + 1.0)


def R(t, u, v, n, p, PC, RPC2):
    if t == u == v == 0:
        return (-2 * p) ** n * boys(n, p * RPC2)
    if t < 0 or u < 0 or v < 0:
        return 0.0
    if t > 0:
        return (t - 1) * R(t - 2, u, v, n + 1, p, PC, RPC2) + PC[0] * R(t - 1, u, v, n + 1, p, PC, RPC2)
    if u > 0:
        return (u - 1) * R(t, u - 2, v, n + 1, p, PC, RPC2) + PC[1] * R(t, u - 1, v, n + 1, p, PC, RPC2)
    return (v - 1) * R(t, u, v - 2, n + 1, p, PC, RPC2) + PC[2] * R(t, u, v - 1, n + 1, p, PC, RPC2)


def _overlap_prim(a, la, A, b, lb, B):
    p = a + b
    s = 1.0
    for k in range(3):
        s *= E(la[k], lb[k], 0, A[k] - B[k], a, b)
    return s * (math.pi / p) ** 1.5


def _kinetic_prim(a, la, A, b, lb, B):
    l2, m2, n2 = lb
    t0 = b * (2 * (l2 + m2 + n2) + 3) * _overlap_prim(a, la, A, b, lb, B)
    t1 = -2 * b * b * (_overlap_prim(a, la, A, b, (l2 + 2, m2, n2), B)
                       + _overlap_prim(a, la, A, b, (l2, m2 + 2, n2), B)
                       + _overlap_prim(a, la, A, b, (l2, m2, n2 + 2), B))
    t2 = -0.5 * (l2 * (l2 - 1) * _overlap_prim(a, la, A, b, (l2 - 2, m2, n2), B)
                 + m2 * (m2 - 1) * _overlap_prim(a, la, A, b, (l2, m2 - 2, n2), B)
                 + n2 * (n2 - 1) * _overlap_prim(a, la, A, b, (l2, m2, n2 - 2), B))
    return t0 + t1 + t2


def _nuclear_prim(a, la, A, b, lb, B, C):
    p = a + b
    P = (a * A + b * B) / p
    PC = P - C
    RPC2 = float(PC @ PC)
    val = 0.0
    for t in range(la[0] + lb[0] + 1):
        for u in range(la[1] + lb[1] + 1):
            for v in range(la[2] + lb[2] + 1):
                val += (E(la[0], lb[0], t, A[0] - B[0], a, b) * E(la[1], lb[1], u, A[1] - B[1], a, b)
                        * E(la[2], lb[2], v, A[2] - B[2], a, b) * R(t, u, v, 0, p, PC, RPC2))
    return val * 2 * math.pi / p


def _eri_prim(a, la, A, b, lb, B, c, lc, Cc, d, ld, D):
    p, q = a + b, c + d
    alpha = p * q / (p + q)
    P = (a * A + b * B) / p
    Q = (c * Cc + d * D) / q
    PQ = P - Q
    RPQ2 = float(PQ @ PQ)
    Eab = [[E(la[k], lb[k], t, A[k] - B[k], a, b) for t in range(la[k] + lb[k] + 1)] for k in range(3)]
    Ecd = [[E(lc[k], ld[k], t, Cc[k] - D[k], c, d) for t in range(lc[k] + ld[k] + 1)] for k in range(3)]
    val = 0.0
    for t, et in enumerate(Eab[0]):
        for u, eu in enumerate(Eab[1]):
            for v, ev in enumerate(Eab[2]):
                for tau, ft in enumerate(Ecd[0]):
                    for nu, fn in enumerate(Ecd[1]):
                        for phi, fp in enumerate(Ecd[2]):
                            val += (et * eu * ev * ft * fn * fp * (-1) ** (tau + nu + phi)
                                    * R(t + tau, u + nu, v + phi, 0, alpha, PQ, RPQ2))
    return val * 2 * math.pi ** 2.5 / (p * q * math.sqrt(p + q))


def _contract(fn, *bfs, extra=()):
    val = 0.0
    for idx in itertools.product(*[range(len(b.exps)) for b in bfs]):
        c = 1.0
        args = []
        for b, k in zip(bfs, idx):
            c *= b.coefs[k]
            args += [b.exps[k], b.shell, b.origin]
        val += c * fn(*args, *extra)
    return val


def integrals(atoms):
    bfs = build_basis(atoms)
    n = len(bfs)
    S = np.zeros((n, n)); T = np.zeros((n, n)); V = np.zeros((n, n))
    for i in range(n):
        for j in range(i + 1):
            S[i, j] = S[j, i] = _contract(_overlap_prim, bfs[i], bfs[j])
            T[i, j] = T[j, i] = _contract(_kinetic_prim, bfs[i], bfs[j])
            v = 0.0
            for sym, xyz in atoms:
                v -= Z[sym] * _contract(_nuclear_prim, bfs[i], bfs[j], extra=(np.asarray(xyz, float),))
            V[i, j] = V[j, i] = v
    g = np.zeros((n, n, n, n))
    for i in range(n):
        for j in range(i + 1):
            ij = i * (i + 1) // 2 + j
            for k in range(n):
                for l in range(k + 1):
                    kl = k * (k + 1) // 2 + l
                    if ij < kl:
                        continue
                    val = _contract(_eri_prim, bfs[i], bfs[j], bfs[k], bfs[l])
                    for (a, b, c, d) in ((i, j, k, l), (j, i, k, l), (i, j, l, k), (j, i, l, k),
                                         (k, l, i, j), (l, k, i, j), (k, l, j, i), (l, k, j, i)):
                        g[a, b, c, d] = val
    enuc = 0.0
    for (s1, x1), (s2, x2) in itertools.combinations(atoms, 2):
        enuc += Z[s1] * Z[s2] / np.linalg.norm(np.asarray(x1) - np.asarray(x2))
    return S, T, V, g, enuc


def rhf(S, H, g, nocc, maxit=200, tol=1e-13):
    """Closed-shell SCF with DIIS, converged far tighter than the reference's RHF.jl:77-138."""
    s, U = np.linalg.eigh(S)
    X = U @ np.diag(s ** -0.5) @ U.T
    def fock(Dm):
        return H + 2 * np.einsum("mnls,ls->mn", g, Dm) - np.einsum("mlns,ls->mn", g, Dm)
    e, C = np.linalg.eigh(X @ H @ X)
    C = X @ C
    Dm = C[:, :nocc] @ C[:, :nocc].T
    Fs, Es = [], []
    Eold = 0.0
    for it in range(maxit):
        F = fock(Dm)
        err = X @ (F @ Dm @ S - S @ Dm @ F) @ X
        Fs.append(F); Es.append(err)
        Fs, Es = Fs[-8:], Es[-8:]
        if len(Fs) > 1:
            m = len(Fs)
            B = -np.ones((m + 1, m + 1)); B[m, m] = 0
            for a in range(m):
                for b in range(m):
                    B[a, b] = np.vdot(Es[a], Es[b])
            rhs = np.zeros(m + 1); rhs[m] = -1
            c = np.linalg.solve(B, rhs)[:m]
            F = sum(ci * Fi for ci, Fi in zip(c, Fs))
        e, C = np.linalg.eigh(X @ F @ X)
        C = X @ C
        Dm = C[:, :nocc] @ C[:, :nocc].T
        Eel = np.sum(Dm * (H + fock(Dm)))
        if abs(Eel - Eold) < tol and np.abs(err).max() < 1e-11:
            break
        Eold = Eel
    return Eel, C, e


def h2o_geometry(r_ang=1.1, angle_deg=104.0):
    r = r_ang * ANGSTROM
    th = math.radians(angle_deg)
    return [("O", (0.0, 0.0, 0.0)), ("H", (r, 0.0, 0.0)), ("H", (r * math.cos(th), r * math.sin(th), 0.0))]


def generate(outdir):
    os.makedirs(outdir, exist_ok=True)
    atoms = h2o_geometry()
    S, T, V, g, enuc = integrals(atoms)
    Eel, C, eps = rhf(S, T + V, g, 5)
    np.savez(os.path.join(outdir, "h2o_sto3g.npz"), S=S, H=T + V, g=g, C=C, eps=eps, enuc=enuc,
             escf=Eel + enuc, nocc=5)
    print("H2O/STO-3G  E_SCF = %.12f  (Crawford's programming-project value: -74.942079928192)" % (Eel + enuc))
    h2 = [("H", (0.0, 0.0, 0.0)), ("H", (0.0, 0.0, 1.0 * ANGSTROM))]
    S2, T2, V2, g2, en2 = integrals(h2)
    np.savez(os.path.join(outdir, "h2_sto3g.npz"), S=S2, H=T2 + V2, g=g2, enuc=en2)
    print("H2 (11|11) = %.12f  (21|22) = %.12f  (11|22) = %.12f" % (g2[0, 0, 0, 0], g2[1, 0, 1, 1], g2[0, 0, 1, 1]))


if __name__ == "__main__":
    generate(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden"))
