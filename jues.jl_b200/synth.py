"""
Synthetic inputs of the shapes BASELINE.json names.

psi4 / libint2 / basis sets are not available offline, so benchmarks and parity tests use
8-fold-symmetric random "ERIs", an orthogonal MO coefficient matrix and gapped orbital
energies.  Two generators:

* ``dense_inputs``   -- numpy PCG64 normal deviates symmetrised over the 8 permutations;
                        for shapes whose N^4 fits on the host (N <~ 200).
* ``counter_eri``    -- counter-based (splitmix64 of the canonical pair-of-pairs index), so
                        any element / slab can be regenerated bit-identically on the host and
                        on the device (``jues_b200_synth_eri`` in the C ABI evaluates the
                        same function in a CUDA kernel) without ever holding N^4 anywhere.

This is host-side input plumbing (numpy); no arithmetic of the hot path happens here.
"""
from __future__ import annotations

import numpy as np

__all__ = ["dense_eri", "orbitals", "dense_inputs", "counter_eri", "counter_eri_element", "counter_scale",
           "default_scale", "core_hamiltonian", "noncanonical_inputs"]

_MASK = (1 << 64) - 1


def default_scale(nbf: int) -> float:
    """s = 0.4/N keeps the un-accelerated Jacobi CC iterations of the reference convergent
    (SURVEY.md section 8d)."""
    return 0.4 / nbf


def counter_scale(nbf: int) -> float:
    """Scale for the counter-based (uniform, un-averaged) generator that gives its elements the
    same standard deviation as `dense_eri` (normal deviates averaged over the 8 permutations):
    0.4/N * sqrt(3/8).  Larger values make the reference's DIIS-free Jacobi sweeps diverge."""
    return default_scale(nbf) * (3.0 / 8.0) ** 0.5


def dense_eri(nbf: int, seed: int = 2024, scale: float | None = None) -> np.ndarray:
    """(nbf,)*4 array with the 8-fold permutational symmetry of real ERIs, Fortran order."""
    s = default_scale(nbf) if scale is None else scale
    rng = np.random.default_rng(seed)
    g0 = rng.standard_normal((nbf, nbf, nbf, nbf)) * s
    g = g0 + g0.transpose(1, 0, 2, 3)
    g = g + g.transpose(0, 1, 3, 2)
    g = g + g.transpose(2, 3, 0, 1)
    g *= 0.125
    return np.asfortranarray(g)


def orbitals(nbf: int, nocc: int, seed: int = 2024):
    """Orthogonal C (QR of a normal matrix) split into Cao/Cav, and gapped orbital energies
    eps = [linspace(-2,-0.6,o), linspace(0.4,2.5,v)] (HOMO-LUMO gap 1.0)."""
    rng = np.random.default_rng(seed + 1_000_003)
    C, _ = np.linalg.qr(rng.standard_normal((nbf, nbf)))
    nvir = nbf - nocc
    eps = np.concatenate([np.linspace(-2.0, -0.6, nocc), np.linspace(0.4, 2.5, nvir)])
    return (np.asfortranarray(C[:, :nocc]), np.asfortranarray(C[:, nocc:]), eps)


def dense_inputs(nbf: int, nocc: int, seed: int = 2024, scale: float | None = None):
    """-> (gao, Cao, Cav, eps) for a shape (nbf, nocc)."""
    gao = dense_eri(nbf, seed, scale)
    Cao, Cav, eps = orbitals(nbf, nocc, seed)
    return gao, Cao, Cav, eps


def core_hamiltonian(gao, Cao, Cav, eps) -> np.ndarray:
    """The hao for which (Cao, Cav, eps) IS the converged RHF solution of the synthetic gao:
    hao = C diag(eps) C^T - G[D], G = 2J - K of the density of Cao (C is orthogonal), so that
    get_fock (IntegralTransformation.jl:119-141) returns diag(eps) in the basis of C.  Input
    plumbing for the AutoRCCSD shapes (N^4 numpy work, N <~ 200)."""
    C = np.concatenate([Cao, Cav], axis=1)
    D = Cao @ Cao.T
    G = 2.0 * np.einsum("mnls,ls->mn", gao, D, optimize=True) - np.einsum("mlns,ls->mn", gao, D, optimize=True)
    return np.asfortranarray(C @ np.diag(eps) @ C.T - G)


def noncanonical_inputs(nbf: int, nocc: int, seed: int = 2024, scale: float | None = None,
                        ov_mix: float = 0.0):
    """-> (gao, hao, Ca, eps_canonical): the `dense_inputs` problem with its occupied orbitals
    rotated among themselves and its virtuals among themselves by random orthogonal matrices
    (a non-canonical RHF reference: f_oo, f_vv off-diagonal != 0, f_ov = 0, same CCSD energy),
    plus an optional occupied-virtual rotation of angle ~ov_mix (f_ov != 0: not an RHF solution
    any more, but a valid input of AutoRCCSD's equations)."""
    gao, Cao, Cav, eps = dense_inputs(nbf, nocc, seed, scale)
    hao = core_hamiltonian(gao, Cao, Cav, eps)
    rng = np.random.default_rng(seed + 7_000_003)
    Uo, _ = np.linalg.qr(rng.standard_normal((nocc, nocc)))
    Uv, _ = np.linalg.qr(rng.standard_normal((nbf - nocc, nbf - nocc)))
    Ca = np.concatenate([Cao @ Uo, Cav @ Uv], axis=1)
    if ov_mix:
        K = rng.standard_normal((nbf, nbf)) * ov_mix
        K = K - K.T
        # orthogonal rotation exp(K) by its Cayley transform (orthogonal to rounding)
        I = np.eye(nbf)
        Ca = Ca @ np.linalg.solve(I - 0.5 * K, I + 0.5 * K)
    return gao, hao, np.asfortranarray(Ca), eps


# ---------------------------------------------------------------------------------------
# counter-based generator (bit-identical on host and device)
# ---------------------------------------------------------------------------------------
def _splitmix64(x: np.ndarray) -> np.ndarray:
    x = (x + np.uint64(0x9E3779B97F4A7C15))
    z = x
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def _canon(mu, nu, lam, sig):
    mu = np.asarray(mu, dtype=np.uint64)
    nu = np.asarray(nu, dtype=np.uint64)
    lam = np.asarray(lam, dtype=np.uint64)
    sig = np.asarray(sig, dtype=np.uint64)
    hi, lo = np.maximum(mu, nu), np.minimum(mu, nu)
    P = hi * (hi + np.uint64(1)) // np.uint64(2) + lo
    hi, lo = np.maximum(lam, sig), np.minimum(lam, sig)
    Q = hi * (hi + np.uint64(1)) // np.uint64(2) + lo
    hi, lo = np.maximum(P, Q), np.minimum(P, Q)
    return hi * (hi + np.uint64(1)) // np.uint64(2) + lo


def counter_eri_element(mu, nu, lam, sig, seed: int, scale: float):
    """g(mu nu|lam sig) = scale * (2u-1), u = (splitmix64(seed ^ K) >> 11) * 2^-53, K the
    canonical index of the symmetry-unique quadruple.  Vectorised over array arguments."""
    with np.errstate(over="ignore"):
        K = _canon(mu, nu, lam, sig)
        h = _splitmix64(np.uint64(seed & _MASK) ^ K)
    u = (h >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
    return scale * (2.0 * u - 1.0)


def counter_eri(nbf: int, seed: int = 2024, scale: float | None = None,
                sig_range: tuple[int, int] | None = None) -> np.ndarray:
    """Dense (Fortran-order) array of the counter-based ERIs, optionally only the slab
    sig in [lo, hi) (shape (nbf, nbf, nbf, hi-lo))."""
    s = counter_scale(nbf) if scale is None else scale
    lo, hi = (0, nbf) if sig_range is None else sig_range
    idx = np.arange(nbf, dtype=np.uint64)
    sig = np.arange(lo, hi, dtype=np.uint64)
    g = counter_eri_element(idx[:, None, None, None], idx[None, :, None, None],
                            idx[None, None, :, None], sig[None, None, None, :], seed, s)
    return np.asfortranarray(g)
