"""
jues.jl_b200 -- B200 (sm_100a) implementation of the JuES.jl hot path behind the reference's
own interface.

The reference is Julia; there is no Julia toolchain in the build image, so the executable
host-side mirror of its operator interface is this module (the Julia shim that `ccall`s the
same C ABI is ``jues.jl_b200/julia/JuESB200.jl``).  Names, argument meaning, return values and
error behaviour follow the reference:

===============================  =========================================================
here                             reference
===============================  =========================================================
``Wfn``                          JuES.Wavefunction.Wfn            (Wavefunction.jl:67-88)
``tei_transform``                JuES.Transformation.tei_transform (Transformation.jl:15-93)
``get_eri``                      JuES.IntegralTransformation.get_eri (IntegralTransformation.jl:38-101)
``do_rmp2``                      JuES.MollerPlesset.do_rmp2       (RMP2.jl:11-45)
``RCCD.do_rccd``                 JuES.CoupledCluster.RCCD.do_rccd  (RCCD.jl:33-83)
``RCCSD.do_rccsd``               JuES.CoupledCluster.RCCSD.do_rccsd (RCCSD.jl:33-116)
``AutoRCCSD.do_rccsd``           JuES.CoupledCluster.AutoRCCSD.do_rccsd (AutoRCCSD.jl:193-301)
``mRCCD.do_rccd``                JuES.CoupledCluster.mRCCD.do_rccd (mRCCD.jl:37-120; Float32 DIIS :143-207)
``get_fock``                     JuES.IntegralTransformation.get_fock (IntegralTransformation.jl:119-141)
``compute_pT``                   JuES.CoupledCluster.PerturbativeTriples.compute_pT (PerturbativeTriples.jl:35-138)
``DeviceFourTensor``             JuES.DiskTensors.DiskFourTensor  (DiskFourTensors.jl:5-95)
``gemm``                         LinearAlgebra.BLAS.gemm!         (mRCCD.jl:268-485 call sites)
===============================  =========================================================

All arithmetic happens in ``libjues_b200.so`` (hand-written CUDA, C ABI in
``include/jues_b200.h``).  There is no CPU fallback: without the shared library, or without
an sm_100 device, every compute entry point raises.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

from . import synth  # noqa: F401
from . import _lib
from ._lib import LibraryMissing  # noqa: F401

__all__ = ["Wfn", "Context", "DeviceFourTensor", "tei_transform", "get_eri", "do_rmp2", "RCCD",
           "RCCSD", "AutoRCCSD", "mRCCD", "DF", "DFRCCD", "do_df_rmp2", "get_fock", "compute_pT", "CC_DEFAULTS", "gemm", "JuesError", "LibraryMissing", "default_context", "synth"]

ERROR_NAMES = {-1: "EINVAL", -2: "ENOMEM", -3: "ECUDA", -4: "ENCCL", -5: "ESTATE"}


class JuesError(RuntimeError):
    """Raised where the reference would `error("...")` (IntegralTransformation.jl:41-43) or
    when the device library reports a failure."""

    def __init__(self, code: int, message: str):
        super().__init__(f"jues_b200 {ERROR_NAMES.get(code, code)}: {message}")
        self.code = code


def _f(a, shape=None) -> np.ndarray:
    """Column-major float64 view/copy: exactly what Julia hands to `ccall`."""
    a = np.asarray(a, dtype=np.float64)
    if not a.flags.f_contiguous:
        a = np.asfortranarray(a)
    if shape is not None and tuple(a.shape) != tuple(shape):
        raise JuesError(-1, f"array of shape {a.shape}, expected {tuple(shape)}")
    return a


def _p(a: Optional[np.ndarray]):
    return a.ctypes.data_as(_lib.c_double_p) if a is not None else None


# ------------------------------------------------------------------------------------------
# Wfn: the input record (Wavefunction.jl:67-88); only the fields the path reads are required.
# ------------------------------------------------------------------------------------------
@dataclass
class Wfn:
    nalpha: int
    nvira: int
    epsa: np.ndarray
    Cao: np.ndarray
    Cav: np.ndarray
    ao_eri: object            # ndarray (nbf,)*4 or DeviceFourTensor (Wavefunction.jl:87 Union)
    nbeta: int = field(default=-1)
    nvirb: int = field(default=-1)
    # read only by get_fock / AutoRCCSD (Wavefunction.jl:67,77,83)
    hao: Optional[np.ndarray] = None      # (nbf, nbf) core Hamiltonian
    Ca: Optional[np.ndarray] = None       # (nbf, nmo) all MO coefficients; default [Cao Cav]
    energy: float = 0.0                   # reference energy (only printed by the reference)
    # density fitting: (pqP, Jpqh) exactly as DF.setup_df returns them (DF.jl:30-51).  The reference asks psi4
    # for them; psi4 is not assumed here, so whoever builds the Wfn supplies them.
    df: Optional[tuple] = None

    def __post_init__(self):
        if self.nbeta < 0:
            self.nbeta = self.nalpha
        if self.nvirb < 0:
            self.nvirb = self.nvira
        if self.Ca is None:
            self.Ca = np.concatenate([np.asarray(self.Cao), np.asarray(self.Cav)], axis=1)

    @property
    def Cb(self):
        return self.Ca

    @property
    def Cbo(self):
        return self.Cao

    @property
    def Cbv(self):
        return self.Cav

    @property
    def nmo(self):
        return self.nalpha + self.nvira


# ------------------------------------------------------------------------------------------
# Context
# ------------------------------------------------------------------------------------------
class Context:
    """One GPU, one stream, one process rank (jues_ctx of the C ABI)."""

    def __init__(self, device: int = 0):
        self._lib = _lib.load()
        h = C.c_void_p()
        rc = self._lib.jues_b200_init(C.byref(h), int(device))
        if rc != 0:
            raise JuesError(rc, self._lib.jues_b200_last_error(None).decode())
        self._h = h
        self.device = device
        self._cb_keep = None

    @classmethod
    def multi(cls, ngpu: int = 0) -> "Context":
        """Single-process multi-GPU: one handle in front of `ngpu` devices (0 = every visible one); the
        sharding entry points called with it run on all of them (jues_b200_init_multi)."""
        self = cls.__new__(cls)
        self._lib = _lib.load()
        h = C.c_void_p()
        rc = self._lib.jues_b200_init_multi(C.byref(h), int(ngpu))
        if rc != 0:
            raise JuesError(rc, self._lib.jues_b200_last_error(None).decode())
        self._h = h
        self.device = 0
        self._cb_keep = None
        self.nranks = self._lib.jues_b200_group_size(h)
        return self

    # -- plumbing ---------------------------------------------------------------------------
    def _check(self, rc: int):
        if rc != 0:
            raise JuesError(rc, self._lib.jues_b200_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None):
            self._lib.jues_b200_finalize(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- instrumentation --------------------------------------------------------------------
    def phases(self):
        """[(name, ms)] CUDA-event timings of the last entry-point call."""
        n = self._lib.jues_b200_get_phases(self._h, None, 0)
        buf = (_lib.Phase * max(n, 1))()
        self._lib.jues_b200_get_phases(self._h, buf, n)
        return [(buf[i].name.decode(), buf[i].ms) for i in range(n)]

    def set_trace(self, level: int):
        """0 coarse phases, 1 regions (cc.part.*, cc.comm.*, tei.*), 2 every DGEMM launch ("gemm MxNxKxb")."""
        self._check(self._lib.jues_b200_set_trace(self._h, int(level)))

    def counters(self):
        fl = C.c_double()
        gl, al, bp = C.c_int64(), C.c_int64(), C.c_int64()
        self._check(self._lib.jues_b200_get_counters(self._h, C.byref(fl), C.byref(gl), C.byref(al), C.byref(bp)))
        return {"gemm_flops": fl.value, "gemm_launches": gl.value, "aux_launches": al.value,
                "bytes_peak": bp.value}

    def comm_counters(self):
        n, b = C.c_int64(), C.c_double()
        self._check(self._lib.jues_b200_get_comm_counters(self._h, C.byref(n), C.byref(b)))
        return {"collectives": n.value, "bytes_received": b.value}

    # -- multi-GPU: one process per GPU ---------------------------------------------------------
    def init_dist(self, rank: Optional[int] = None, nranks: Optional[int] = None):
        """Attach this context to the job's NCCL communicator.  The 128-byte unique id is made by
        rank 0 (jues_b200_nccl_unique_id) and broadcast with torch.distributed, which the host
        program must have initialised (any backend; gloo is enough)."""
        import torch
        import torch.distributed as dist
        rank = dist.get_rank() if rank is None else rank
        nranks = dist.get_world_size() if nranks is None else nranks
        idbuf = (C.c_ubyte * 128)()
        if nranks > 1:
            if rank == 0:
                rc = self._lib.jues_b200_nccl_unique_id(idbuf)
                if rc != 0:
                    raise JuesError(rc, self._lib.jues_b200_last_error(None).decode())
            t = torch.tensor(list(bytes(idbuf)), dtype=torch.uint8)
            if dist.get_backend() == "nccl":
                t = t.cuda()
            dist.broadcast(t, src=0)
            idbuf = (C.c_ubyte * 128)(*t.cpu().tolist())
        self._check(self._lib.jues_b200_init_dist(self._h, int(rank), int(nranks), idbuf))
        self.rank, self.nranks = rank, nranks

    def set_amplitude_callback(self, fn):
        """fn(it, energy, T1 or None, T2) after the guess and after every sweep (tests)."""
        if fn is None:
            self._cb_keep = None
            self._check(self._lib.jues_b200_set_amplitude_callback(self._h, _lib.AMP_CB(), None))
            return
        shapes = {}

        def tramp(user, it, e, p1, p2):
            o, v = shapes["o"], shapes["v"]
            T2 = np.ctypeslib.as_array(p2, shape=(o * o * v * v,)).reshape((o, o, v, v), order="F").copy()
            T1 = None
            if p1:
                T1 = np.ctypeslib.as_array(p1, shape=(o * v,)).reshape((o, v), order="F").copy()
            fn(it, e, T1, T2)

        cb = _lib.AMP_CB(tramp)
        self._cb_keep = (cb, shapes)
        self._check(self._lib.jues_b200_set_amplitude_callback(self._h, cb, None))

    def _cb_shapes(self, o, v):
        if self._cb_keep is not None:
            self._cb_keep[1]["o"] = o
            self._cb_keep[1]["v"] = v

    # -- BLAS.gemm! -------------------------------------------------------------------------
    def gemm(self, tA: str, tB: str, alpha: float, A, B, beta: float = 0.0, Cm=None):
        """C = alpha*op(A)*op(B) + beta*C (BLAS.gemm!(tA,tB,alpha,A,B,beta,C))."""
        A = _f(A)
        B = _f(B)
        M, K = (A.shape[1], A.shape[0]) if tA in "Tt" else A.shape
        K2, N = (B.shape[1], B.shape[0]) if tB in "Tt" else B.shape
        if K != K2:
            raise JuesError(-1, f"gemm: inner dimensions differ ({K} vs {K2})")
        if Cm is None:
            Cm = np.zeros((M, N), order="F")
            beta = 0.0
        else:
            Cm = _f(Cm, (M, N))
        self._check(self._lib.jues_b200_dgemm(self._h, tA.encode(), tB.encode(), M, N, K, alpha,
                                              _p(A), max(A.shape[0], 1), _p(B), max(B.shape[0], 1), beta,
                                              _p(Cm), max(M, 1)))
        return Cm

    def sa_ladder(self, tau, W4, nslabs: int = 1) -> np.ndarray:
        """sum_ef tau[i,j,e,f] W4[e,f,a,b] through the packed symmetric/antisymmetric pair space, evaluated
        as `nslabs` ranks of the coupled-cluster driver would (kernel-level check, jues_b200_sa_ladder)."""
        tau = _f(tau)
        o, _, v, _ = tau.shape
        tau = _f(tau, (o, o, v, v))
        W4 = _f(W4, (v, v, v, v))
        out = np.empty((o, o, v, v), order="F")
        self._check(self._lib.jues_b200_sa_ladder(self._h, _p(tau), _p(W4), o, v, int(nslabs), _p(out)))
        return out

    def gemm_bench(self, tA: str, tB: str, M: int, N: int, K: int, reps: int = 3) -> float:
        ms = C.c_double()
        self._check(self._lib.jues_b200_dgemm_bench(self._h, tA.encode(), tB.encode(), M, N, K, reps,
                                                    C.byref(ms)))
        return ms.value


    def gemm_stress(self, tA: str, tB: str, M: int, N: int, K: int, batch: int = 1, reps: int = 20):
        """(repetitions that differ from the first launch, worst sum of squared differences): 0, 0.0 expected."""
        nb, w = C.c_int(), C.c_double()
        self._check(self._lib.jues_b200_dgemm_stress(self._h, tA.encode(), tB.encode(), M, N, K, batch, reps,
                                                     C.byref(nb), C.byref(w)))
        return nb.value, w.value


def slab_bounds(nvir: int, nranks: int, rank: int):
    """Host-side mirror of the library's partition of the virtual index over ranks
    (csrc/cc.cu: setup_problem + dist.h: slab_of): nvir is padded to a multiple of 2*nranks and
    split into equal slabs.  Returns (padded nvir, first, last+1) of `rank`'s slab."""
    vp = -(-int(nvir) // (2 * nranks)) * (2 * nranks)
    vs = vp // nranks
    return vp, rank * vs, (rank + 1) * vs


_default_ctx: Optional[Context] = None


def default_context() -> Context:
    """Process-wide context on the device LOCAL_RANK (or 0) -- created on first use."""
    global _default_ctx
    if _default_ctx is None:
        import os
        _default_ctx = Context(int(os.environ.get("LOCAL_RANK", "0")))
    return _default_ctx


# ------------------------------------------------------------------------------------------
# DeviceFourTensor: the DiskFourTensor replacement (DiskFourTensors.jl:5-95)
# ------------------------------------------------------------------------------------------
def _ranger(idx, dim):
    """DiskTensors.ranger (DiskTensors.jl:28-34): Int / range / Colon -> half-open (lo, hi),
    plus whether the axis is dropped (integer index)."""
    if isinstance(idx, (int, np.integer)):
        i = int(idx)
        if i < 0:
            i += dim
        if not 0 <= i < dim:
            raise JuesError(-1, f"index {idx} out of range for extent {dim}")
        return i, i + 1, True
    if isinstance(idx, slice):
        if idx.step not in (None, 1):
            raise JuesError(-1, "DeviceFourTensor slices must have unit stride")
        for b in (idx.start, idx.stop):           # Julia raises BoundsError; do not clamp silently
            if b is not None and not (-dim <= b <= dim):
                raise JuesError(-1, f"slice {idx} out of range for extent {dim}")
        lo, hi, _ = idx.indices(dim)
        return lo, max(hi, lo), False
    raise JuesError(-1, f"unsupported index {idx!r}")


class DeviceFourTensor:
    """Rank-4 Float64 tensor resident in HBM with the DiskFourTensor surface: 4-index
    getindex / setindex! with integers, unit-stride ranges and `:`, `eltype`, `blockfill!`.
    Indices are 0-based (Python) where the reference's are 1-based (Julia)."""

    def __init__(self, d1: int, d2: int, d3: int, d4: int, ctx: Optional[Context] = None, _handle=None):
        self.ctx = ctx or default_context()
        self.shape = (int(d1), int(d2), int(d3), int(d4))
        if _handle is not None:
            self._h = _handle
        else:
            h = C.c_void_p()
            self.ctx._check(self.ctx._lib.jues_b200_t4_create(self.ctx._h, *self.shape, C.byref(h)))
            self._h = h

    @classmethod
    def from_array(cls, a, ctx: Optional[Context] = None) -> "DeviceFourTensor":
        a = _f(a)
        if a.ndim != 4:
            raise JuesError(-1, "DeviceFourTensor needs a rank-4 array")
        t = cls(*a.shape, ctx=ctx)
        t[:, :, :, :] = a
        return t

    @classmethod
    def synth_eri(cls, nbf: int, seed: int = 2024, scale: Optional[float] = None,
                  ctx: Optional[Context] = None, virtual: bool = False) -> "DeviceFourTensor":
        """Counter-based synthetic ERIs generated on the device (== synth.counter_eri).
        virtual=True: no storage; sigma slabs are generated on demand while a transformation
        streams through them (for nbf whose dense N^4 does not fit in HBM); read-only."""
        if virtual:
            ctx = ctx or default_context()
            s = synth.counter_scale(nbf) if scale is None else scale
            h = C.c_void_p()
            ctx._check(ctx._lib.jues_b200_t4_create_synth(ctx._h, nbf, C.c_uint64(seed), s, C.byref(h)))
            return cls(nbf, nbf, nbf, nbf, ctx=ctx, _handle=h)
        t = cls(nbf, nbf, nbf, nbf, ctx=ctx)
        s = synth.counter_scale(nbf) if scale is None else scale
        t.ctx._check(t.ctx._lib.jues_b200_t4_synth_eri(t._h, C.c_uint64(seed), s))
        return t

    @property
    def eltype(self):
        return np.float64

    dtype = eltype

    def blockfill(self, value: float):
        self.ctx._check(self.ctx._lib.jues_b200_t4_fill(self._h, float(value)))

    def _ranges(self, key):
        if not isinstance(key, tuple) or len(key) != 4:
            raise JuesError(-1, "DeviceFourTensor takes exactly four indices")
        r = [_ranger(k, d) for k, d in zip(key, self.shape)]
        lo = (C.c_int64 * 4)(*[x[0] for x in r])
        hi = (C.c_int64 * 4)(*[x[1] for x in r])
        ext = tuple(x[1] - x[0] for x in r)
        drop = tuple(x[2] for x in r)
        return lo, hi, ext, drop

    def __getitem__(self, key):
        lo, hi, ext, drop = self._ranges(key)
        out = np.empty(ext, order="F")
        self.ctx._check(self.ctx._lib.jues_b200_t4_get_slice(self._h, lo, hi, _p(out)))
        keep = tuple(e for e, d in zip(ext, drop) if not d)
        if not keep:
            return float(out.reshape(-1)[0])
        return out.reshape(keep, order="F")

    def __setitem__(self, key, value):
        lo, hi, ext, drop = self._ranges(key)
        v = np.asarray(value, dtype=np.float64)
        if v.ndim == 0:
            v = np.full(ext, float(v), order="F")
        else:
            keep = tuple(e for e, d in zip(ext, drop) if not d)
            if tuple(v.shape) != keep:
                raise JuesError(-1, f"cannot assign shape {v.shape} to a slice of shape {keep}")
            v = np.asfortranarray(v.reshape(ext, order="F"))
        self.ctx._check(self.ctx._lib.jues_b200_t4_set_slice(self._h, lo, hi, _p(v)))

    def to_array(self) -> np.ndarray:
        return self[:, :, :, :]

    def free(self):
        if getattr(self, "_h", None):
            self.ctx._lib.jues_b200_t4_destroy(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        try:
            if self.ctx._h:
                self.free()
        except Exception:
            pass


# ------------------------------------------------------------------------------------------
# Transformation.tei_transform / IntegralTransformation.get_eri
# ------------------------------------------------------------------------------------------
def tei_transform(gao, C1, C2=None, C3=None, C4=None, name: str = "default", *, phys: bool = False,
                  ctx: Optional[Context] = None):
    """`tei_transform(gao, C, name)` / `tei_transform(gao, C1, C2, C3, C4, name)`
    (Transformation.jl:15-26,39-93).  A DeviceFourTensor `gao` dispatches to the
    device-resident method (Transformation.jl:94-192) and returns a DeviceFourTensor.
    `name` is accepted and unused, as in the in-core reference method."""
    if isinstance(C2, str) and C3 is None:   # tei_transform(gao, C, "name")
        name, C2 = C2, None
    if C2 is None:
        C2 = C3 = C4 = C1
    ctx = ctx or (gao.ctx if isinstance(gao, DeviceFourTensor) else default_context())
    Cs = [_f(c) for c in (C1, C2, C3, C4)]
    nao = gao.shape[0]
    if tuple(gao.shape) != (nao,) * 4:
        raise JuesError(-1, f"gao must be (nao,nao,nao,nao), got {tuple(gao.shape)}")
    for c in Cs:
        if c.ndim != 2 or c.shape[0] != nao:
            raise JuesError(-1, f"coefficient matrix of shape {c.shape} does not match nao={nao}")
    d = [c.shape[1] for c in Cs]
    oshape = (d[0], d[2], d[1], d[3]) if phys else tuple(d)
    if isinstance(gao, DeviceFourTensor):
        h = C.c_void_p()
        ctx._check(ctx._lib.jues_b200_tei_transform_t4(
            ctx._h, gao._h, _p(Cs[0]), d[0], _p(Cs[1]), d[1], _p(Cs[2]), d[2], _p(Cs[3]), d[3],
            1 if phys else 0, C.byref(h)))
        return DeviceFourTensor(*oshape, ctx=ctx, _handle=h)
    g = _f(gao)
    out = np.empty(oshape, order="F")
    ctx._check(ctx._lib.jues_b200_tei_transform(
        ctx._h, _p(g), nao, _p(Cs[0]), d[0], _p(Cs[1]), d[1], _p(Cs[2]), d[2], _p(Cs[3]), d[3],
        1 if phys else 0, _p(out)))
    return out


def get_eri(wfn: Wfn, eri_string: str, notation: str = "phys", fcn: int = 0, ctx: Optional[Context] = None):
    """IntegralTransformation.get_eri (IntegralTransformation.jl:38-101): string-addressed
    transform (o/O/v/V), optional frozen core, physicists' order by default."""
    if len(eri_string.encode()) != 4:
        raise JuesError(-1, f"Invalid string given to JuES.IntegralTransformation.get_eri: {eri_string}")
    s = eri_string
    if notation == "phys":
        s = "".join(s[k] for k in (0, 2, 1, 3))                      # :46-48
    Cs = []
    for ch in s:                                                      # :52-68
        if ch == "o":
            Cs.append(np.asarray(wfn.Cbo)[:, fcn:])
        elif ch == "O":
            Cs.append(np.asarray(wfn.Cao)[:, fcn:])
        elif ch == "v":
            Cs.append(wfn.Cbv)
        elif ch == "V":
            Cs.append(wfn.Cav)
    if len(Cs) != 4:
        raise JuesError(-1, f"Invalid string given to JuES.IntegralTransformation.get_eri: {eri_string}")
    return tei_transform(wfn.ao_eri, *Cs, phys=(notation == "phys"), ctx=ctx)


# ------------------------------------------------------------------------------------------
# MollerPlesset.do_rmp2
# ------------------------------------------------------------------------------------------
def _wfn_args(wfn: Wfn):
    o, v = int(wfn.nalpha), int(wfn.nvira)
    nao = wfn.ao_eri.shape[0]
    Cao = _f(wfn.Cao, (nao, o))
    Cav = _f(wfn.Cav, (nao, v))
    eps = _f(wfn.epsa)
    if eps.shape[0] < o + v:
        raise JuesError(-1, "epsa shorter than nalpha + nvira")
    return nao, o, v, Cao, Cav, eps


def do_rmp2(refWfn: Wfn, ctx: Optional[Context] = None, **kwargs) -> float:
    """MollerPlesset.do_rmp2 (RMP2.jl:11-45).  kwargs are accepted and ignored like the
    reference's."""
    nao, o, v, Cao, Cav, eps = _wfn_args(refWfn)
    g = refWfn.ao_eri
    ctx = ctx or (g.ctx if isinstance(g, DeviceFourTensor) else default_context())
    e = C.c_double()
    if isinstance(g, DeviceFourTensor):
        ctx._check(ctx._lib.jues_b200_rmp2_t4(ctx._h, g._h, _p(Cao), o, _p(Cav), v, _p(eps), C.byref(e)))
    else:
        g = _f(g, (nao,) * 4)
        ctx._check(ctx._lib.jues_b200_rmp2(ctx._h, _p(g), nao, _p(Cao), o, _p(Cav), v, _p(eps), C.byref(e)))
    return e.value


# ------------------------------------------------------------------------------------------
# CoupledCluster.RCCD / RCCSD
# ------------------------------------------------------------------------------------------
class _RCCD:
    """JuES.CoupledCluster.RCCD"""
    MAXIT = 40            # RCCD.jl:34 (hard-wired; kwargs ignored)

    def do_rccd(self, refWfn: Wfn, ctx: Optional[Context] = None, *, _maxit: Optional[int] = None,
                _guess: str = "reference", _return_T2: bool = False, _e_hist: Optional[list] = None,
                **kwargs):
        """RCCD.do_rccd (RCCD.jl:33-83): 40 Jacobi sweeps from T2 = (ij|ab)/D, returns the
        correlation energy.  Public kwargs (doprint, maxit, return_T, ...) are accepted and
        ignored exactly like the reference; the underscore-prefixed keywords are test hooks."""
        nao, o, v, Cao, Cav, eps = _wfn_args(refWfn)
        g = refWfn.ao_eri
        ctx = ctx or (g.ctx if isinstance(g, DeviceFourTensor) else default_context())
        maxit = self.MAXIT if _maxit is None else int(_maxit)
        e = C.c_double()
        hist = np.zeros(maxit + 1)
        T2 = np.empty((o, o, v, v), order="F") if _return_T2 else None
        gm = {"reference": 0, "mp2": 1}[_guess]
        ctx._cb_shapes(o, v)
        if isinstance(g, DeviceFourTensor):
            ctx._check(ctx._lib.jues_b200_rccd_t4(ctx._h, g._h, _p(Cao), o, _p(Cav), v, _p(eps), maxit, gm,
                                                  C.byref(e), _p(hist), _p(T2)))
        else:
            g = _f(g, (nao,) * 4)
            ctx._check(ctx._lib.jues_b200_rccd(ctx._h, _p(g), nao, _p(Cao), o, _p(Cav), v, _p(eps), maxit,
                                               gm, C.byref(e), _p(hist), _p(T2)))
        if _e_hist is not None:
            _e_hist[:] = hist.tolist()
        return (e.value, T2) if _return_T2 else e.value


class _RCCSD:
    """JuES.CoupledCluster.RCCSD"""
    MAXIT = 40            # RCCSD.jl:36

    def do_rccsd(self, refWfn: Wfn, ctx: Optional[Context] = None, *, _maxit: Optional[int] = None,
                 _return_T: bool = False, _e_hist: Optional[list] = None, **kwargs):
        """RCCSD.do_rccsd (RCCSD.jl:33-116): T1 = 0, T2 = <ij|ab>/D, 40 Jacobi sweeps, returns
        the correlation energy.  Public kwargs are accepted and ignored like the reference."""
        nao, o, v, Cao, Cav, eps = _wfn_args(refWfn)
        g = refWfn.ao_eri
        ctx = ctx or (g.ctx if isinstance(g, DeviceFourTensor) else default_context())
        maxit = self.MAXIT if _maxit is None else int(_maxit)
        e = C.c_double()
        hist = np.zeros(maxit + 1)
        T1 = np.empty((o, v), order="F") if _return_T else None
        T2 = np.empty((o, o, v, v), order="F") if _return_T else None
        ctx._cb_shapes(o, v)
        if isinstance(g, DeviceFourTensor):
            ctx._check(ctx._lib.jues_b200_rccsd_t4(ctx._h, g._h, _p(Cao), o, _p(Cav), v, _p(eps), maxit,
                                                   C.byref(e), _p(hist), _p(T1), _p(T2)))
        else:
            g = _f(g, (nao,) * 4)
            ctx._check(ctx._lib.jues_b200_rccsd(ctx._h, _p(g), nao, _p(Cao), o, _p(Cav), v, _p(eps), maxit,
                                                C.byref(e), _p(hist), _p(T1), _p(T2)))
        if _e_hist is not None:
            _e_hist[:] = hist.tolist()
        return (e.value, T1, T2) if _return_T else e.value


RCCD = _RCCD()
RCCSD = _RCCSD()


# ------------------------------------------------------------------------------------------
# IntegralTransformation.get_fock, CoupledCluster.AutoRCCSD, PerturbativeTriples.compute_pT
# ------------------------------------------------------------------------------------------
CC_DEFAULTS = dict(cc_max_iter=50, cc_max_rms=1e-10, cc_e_conv=1e-10, diis=False, do_pT=False, fcn=0)
"""JuES.CoupledCluster.defaults (CoupledCluster.jl:36-43)."""


def get_fock(wfn: Wfn, spin: str = "alpha", ctx: Optional[Context] = None) -> np.ndarray:
    """IntegralTransformation.get_fock (IntegralTransformation.jl:119-141): MO-basis Fock matrix
    (nmo, nmo) from wfn.hao, wfn.ao_eri and the occupied orbitals."""
    if spin.lower() in ("alpha", "up", "a"):
        Cm, Co = wfn.Ca, wfn.Cao
    elif spin.lower() in ("beta", "down", "b"):
        Cm, Co = wfn.Cb, wfn.Cbo
    else:
        raise JuesError(-1, f"Invalid Spin option given to JuES.IntegralTransformation.get_fock: {spin}")
    if wfn.hao is None:
        raise JuesError(-1, "get_fock needs wfn.hao (core Hamiltonian)")
    g = wfn.ao_eri
    nao = g.shape[0]
    Cm = _f(Cm)
    nmo = Cm.shape[1]
    Cm = _f(Cm, (nao, nmo))
    Co = _f(Co)
    no = Co.shape[1]
    Co = _f(Co, (nao, no))
    h = _f(wfn.hao, (nao, nao))
    ctx = ctx or (g.ctx if isinstance(g, DeviceFourTensor) else default_context())
    out = np.empty((nmo, nmo), order="F")
    if isinstance(g, DeviceFourTensor):
        ctx._check(ctx._lib.jues_b200_get_fock_t4(ctx._h, g._h, _p(h), _p(Cm), nmo, _p(Co), no, _p(out)))
    else:
        g = _f(g, (nao,) * 4)
        ctx._check(ctx._lib.jues_b200_get_fock(ctx._h, _p(g), nao, _p(h), _p(Cm), nmo, _p(Co), no, _p(out)))
    return out


class _AutoRCCSD:
    """JuES.CoupledCluster.AutoRCCSD"""

    def do_rccsd(self, wfn: Wfn, ctx: Optional[Context] = None, *, _return_all: bool = False, **kwargs):
        """AutoRCCSD.do_rccsd (AutoRCCSD.jl:193-301).  Options are JuES.CoupledCluster.defaults
        (cc_max_iter, cc_e_conv, cc_max_rms, do_pT, fcn, diis); unknown kwargs are ignored like the
        reference's option loop (:199-205).  The reference returns nothing useful (its last
        expression is an @output); here the CCSD correlation energy is returned -- plus E(T) when
        do_pT -- or, with the test hook ``_return_all``, a dict with everything it prints."""
        opt = dict(CC_DEFAULTS)
        opt.update({k: v for k, v in kwargs.items() if k in CC_DEFAULTS})
        nelec = int(wfn.nalpha) + int(wfn.nbeta)
        if nelec % 2 != 0:                                                          # :209
            raise JuesError(-1, f"Number of electrons must be even for RHF. Given {nelec}")
        if wfn.hao is None:
            raise JuesError(-1, "AutoRCCSD needs wfn.hao (core Hamiltonian) to build the Fock matrix")
        ndocc = nelec // 2
        nmo = int(wfn.nmo)
        g = wfn.ao_eri
        nao = g.shape[0]
        Ca = _f(wfn.Ca, (nao, nmo))
        h = _f(wfn.hao, (nao, nao))
        ctx = ctx or (g.ctx if isinstance(g, DeviceFourTensor) else default_context())
        co = _lib.CCOptions(int(opt["cc_max_iter"]), float(opt["cc_e_conv"]), float(opt["cc_max_rms"]),
                            int(bool(opt["do_pT"])), int(opt["fcn"]), int(bool(opt["diis"])))
        no, nv = ndocc - co.fcn, nmo - ndocc
        e, ept = C.c_double(), C.c_double()
        its, conv = C.c_int(), C.c_int()
        eh = np.zeros(co.cc_max_iter + 1)
        rh = np.zeros(co.cc_max_iter + 1)
        T1 = np.empty((max(no, 0), nv), order="F") if _return_all else None
        T2 = np.empty((max(no, 0), max(no, 0), nv, nv), order="F") if _return_all else None
        ctx._cb_shapes(no, nv)
        tail = (nmo, ndocc, C.byref(co), C.byref(e), C.byref(ept), C.byref(its), C.byref(conv),
                _p(eh), _p(rh), _p(T1), _p(T2))
        if isinstance(g, DeviceFourTensor):
            ctx._check(ctx._lib.jues_b200_auto_rccsd_t4(ctx._h, g._h, _p(h), _p(Ca), *tail))
        else:
            g = _f(g, (nao,) * 4)
            ctx._check(ctx._lib.jues_b200_auto_rccsd(ctx._h, _p(g), nao, _p(h), _p(Ca), *tail))
        n = its.value
        if _return_all:
            return dict(ecc=e.value, ept=ept.value if opt["do_pT"] else None, iterations=n,
                        converged=bool(conv.value), e_hist=eh[:n + 1].copy(), rms_hist=rh[:n + 1].copy(),
                        T1=T1, T2=T2)
        return (e.value, ept.value) if opt["do_pT"] else e.value


AutoRCCSD = _AutoRCCSD()


class _mRCCD:
    """JuES.CoupledCluster.mRCCD"""

    def do_rccd(self, refWfn: Wfn, ctx: Optional[Context] = None, *, maxit: int = 40, doprint: bool = False,
                return_T2: bool = False, _return_all: bool = False):
        """mRCCD.do_rccd(refWfn; maxit=40, doprint=false, return_T2=false) (mRCCD.jl:37-120): RCCD from
        zero amplitudes with the reference's DIIS (Float32 vectors), stops at ||dT2|| < 1e-7.
        Returns the correlation energy, or (energy, T2) with return_T2."""
        nao, o, v, Cao, Cav, eps = _wfn_args(refWfn)
        g = refWfn.ao_eri
        ctx = ctx or (g.ctx if isinstance(g, DeviceFourTensor) else default_context())
        maxit = int(maxit)
        e, its = C.c_double(), C.c_int()
        rh, eh = np.zeros(max(maxit, 1)), np.zeros(max(maxit, 1))
        T2 = np.empty((o, o, v, v), order="F") if (return_T2 or _return_all) else None
        ctx._cb_shapes(o, v)
        tail = (_p(Cao), o, _p(Cav), v, _p(eps), maxit, C.byref(e), C.byref(its), _p(rh), _p(eh), _p(T2))
        if isinstance(g, DeviceFourTensor):
            ctx._check(ctx._lib.jues_b200_mrccd_t4(ctx._h, g._h, *tail))
        else:
            g = _f(g, (nao,) * 4)
            ctx._check(ctx._lib.jues_b200_mrccd(ctx._h, _p(g), nao, *tail))
        if _return_all:
            n = its.value
            return dict(ecc=e.value, iterations=n, rms_hist=rh[:n].copy(), e_hist=eh[:n].copy(), T2=T2)
        return (e.value, T2) if return_T2 else e.value


mRCCD = _mRCCD()


# ------------------------------------------------------------------------------------------
# Density fitting: DF.setup_df, MollerPlesset.do_df_rmp2, CoupledCluster.DFRCCD.do_df_rccd
# ------------------------------------------------------------------------------------------
class _DF:
    """JuES.DF"""

    @staticmethod
    def setup_df(refWfn: Wfn, dfbname: str = "default"):
        """DF.setup_df (DF.jl:30-51) returns (pqP, Jpqh) = ((pq|P), (P|Q)^(-1/2)) from psi4; here they are the
        `df` field of the Wfn (dfbname is accepted for the call signature and has nothing to select)."""
        if refWfn.df is None:
            raise JuesError(-1, "setup_df: this Wfn carries no density-fitting tensors (Wfn.df = (pqP, Jpqh))")
        pqP, Jpqh = refWfn.df
        return pqP, Jpqh


DF = _DF()


def _df_args(refWfn: Wfn, dfbname: str):
    pqP, Jpqh = DF.setup_df(refWfn, dfbname)
    o, v = int(refWfn.nalpha), int(refWfn.nvira)
    nao = np.asarray(refWfn.Cao).shape[0]
    pqP = _f(pqP)
    if pqP.ndim != 3 or pqP.shape[0] != nao or pqP.shape[1] != nao:
        raise JuesError(-1, f"pqP must be (nao, nao, naux) with nao = {nao}")
    naux = pqP.shape[2]
    Jpqh = _f(Jpqh, (naux, naux))
    Cao = _f(refWfn.Cao, (nao, o))
    Cav = _f(refWfn.Cav, (nao, v))
    eps = _f(refWfn.epsa)
    if eps.shape[0] < o + v:
        raise JuesError(-1, "epsa shorter than nalpha + nvira")
    return pqP, nao, naux, Jpqh, Cao, o, Cav, v, eps


def do_df_rmp2(refWfn: Wfn, ctx: Optional[Context] = None, *, dfbname: str = "default") -> float:
    """MollerPlesset.do_df_rmp2(refWfn) (DF-RMP2.jl:1-46)."""
    pqP, nao, naux, Jpqh, Cao, o, Cav, v, eps = _df_args(refWfn, dfbname)
    ctx = ctx or default_context()
    e = C.c_double()
    ctx._check(ctx._lib.jues_b200_df_rmp2(ctx._h, _p(pqP), nao, naux, _p(Jpqh), _p(Cao), o, _p(Cav), v, _p(eps),
                                          C.byref(e)))
    return e.value


class _DFRCCD:
    """JuES.CoupledCluster.DFRCCD"""

    def do_df_rccd(self, refWfn: Wfn, ctx: Optional[Context] = None, *, maxit: int = 40, doprint: bool = False,
                   return_T2: bool = False, dfbname: str = "default", _e_hist: Optional[list] = None):
        """DFRCCD.do_df_rccd(refWfn; maxit=40, doprint=false, return_T2=false, dfbname="default")
        (DF-RCCD.jl:11-54): `maxit` sweeps from the MP2 guess.  Returns the correlation energy, or
        (energy, T2) with return_T2."""
        pqP, nao, naux, Jpqh, Cao, o, Cav, v, eps = _df_args(refWfn, dfbname)
        ctx = ctx or default_context()
        maxit = int(maxit)
        e = C.c_double()
        hist = np.zeros(maxit + 1)
        T2 = np.empty((o, o, v, v), order="F") if return_T2 else None
        ctx._cb_shapes(o, v)
        ctx._check(ctx._lib.jues_b200_df_rccd(ctx._h, _p(pqP), nao, naux, _p(Jpqh), _p(Cao), o, _p(Cav), v, _p(eps),
                                              maxit, C.byref(e), _p(hist), _p(T2)))
        if _e_hist is not None:
            _e_hist[:] = list(hist)
        return (e.value, T2) if return_T2 else e.value


DFRCCD = _DFRCCD()


def compute_pT(*, T1, T2, Vvvvo, Vvooo, Vvovo, fo, fv, ctx: Optional[Context] = None) -> float:
    """PerturbativeTriples.compute_pT (PerturbativeTriples.jl:35-138), keyword arguments and array
    layouts exactly as the reference's."""
    T1 = _f(T1)
    o, v = T1.shape
    T2 = _f(T2, (o, o, v, v))
    Vvvvo = _f(Vvvvo, (v, v, v, o))
    Vvooo = _f(Vvooo, (v, o, o, o))
    Vvovo = _f(Vvovo, (v, o, v, o))
    fo = _f(fo, (o,))
    fv = _f(fv, (v,))
    ctx = ctx or default_context()
    e = C.c_double()
    ctx._check(ctx._lib.jues_b200_compute_pt(ctx._h, _p(T1), _p(T2), _p(Vvvvo), _p(Vvooo), _p(Vvovo),
                                             _p(fo), _p(fv), o, v, C.byref(e)))
    return e.value


def gemm(tA, tB, alpha, A, B, beta=0.0, Cm=None, ctx: Optional[Context] = None):
    return (ctx or default_context()).gemm(tA, tB, alpha, A, B, beta, Cm)
