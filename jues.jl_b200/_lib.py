"""ctypes loader for libjues_b200.so (the C ABI of include/jues_b200.h).

There is no CPU fallback: if the shared library is missing or cannot be loaded the import of
any compute entry point raises, and `Context()` raises when no sm_100 CUDA device exists."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("JUES_B200_LIB") or os.path.join(_HERE, "libjues_b200.so")

c_double_p = C.POINTER(C.c_double)
c_int64_p = C.POINTER(C.c_int64)


class Phase(C.Structure):
    _fields_ = [("name", C.c_char * 48), ("ms", C.c_double)]


class CCOptions(C.Structure):
    """jues_b200_cc_options (CoupledCluster.defaults, CoupledCluster.jl:36-43)."""
    _fields_ = [("cc_max_iter", C.c_int), ("cc_e_conv", C.c_double), ("cc_max_rms", C.c_double),
                ("do_pT", C.c_int), ("fcn", C.c_int), ("diis", C.c_int)]


c_int_p = C.POINTER(C.c_int)

AMP_CB = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_double, c_double_p, c_double_p)

# name -> (restype, argtypes); every symbol include/jues_b200.h declares
SIGNATURES = {
    "jues_b200_init": (C.c_int, [C.POINTER(C.c_void_p), C.c_int]),
    "jues_b200_finalize": (None, [C.c_void_p]),
    "jues_b200_last_error": (C.c_char_p, [C.c_void_p]),
    "jues_b200_version": (C.c_char_p, []),
    "jues_b200_nccl_unique_id": (C.c_int, [C.POINTER(C.c_ubyte)]),
    "jues_b200_init_dist": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_ubyte)]),
    "jues_b200_dgemm": (C.c_int, [C.c_void_p, C.c_char, C.c_char, C.c_int64, C.c_int64, C.c_int64,
                                  C.c_double, c_double_p, C.c_int64, c_double_p, C.c_int64,
                                  C.c_double, c_double_p, C.c_int64]),
    "jues_b200_tei_transform": (C.c_int, [C.c_void_p, c_double_p, C.c_int64,
                                          c_double_p, C.c_int64, c_double_p, C.c_int64,
                                          c_double_p, C.c_int64, c_double_p, C.c_int64,
                                          C.c_int, c_double_p]),
    "jues_b200_rmp2": (C.c_int, [C.c_void_p, c_double_p, C.c_int64, c_double_p, C.c_int64,
                                 c_double_p, C.c_int64, c_double_p, c_double_p]),
    "jues_b200_rccd": (C.c_int, [C.c_void_p, c_double_p, C.c_int64, c_double_p, C.c_int64,
                                 c_double_p, C.c_int64, c_double_p, C.c_int, C.c_int,
                                 c_double_p, c_double_p, c_double_p]),
    "jues_b200_rccsd": (C.c_int, [C.c_void_p, c_double_p, C.c_int64, c_double_p, C.c_int64,
                                  c_double_p, C.c_int64, c_double_p, C.c_int,
                                  c_double_p, c_double_p, c_double_p, c_double_p]),
    "jues_b200_set_amplitude_callback": (C.c_int, [C.c_void_p, AMP_CB, C.c_void_p]),
    "jues_b200_t4_create": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64,
                                      C.POINTER(C.c_void_p)]),
    "jues_b200_t4_destroy": (C.c_int, [C.c_void_p]),
    "jues_b200_t4_dims": (C.c_int, [C.c_void_p, c_int64_p]),
    "jues_b200_t4_fill": (C.c_int, [C.c_void_p, C.c_double]),
    "jues_b200_t4_set_slice": (C.c_int, [C.c_void_p, c_int64_p, c_int64_p, c_double_p]),
    "jues_b200_t4_get_slice": (C.c_int, [C.c_void_p, c_int64_p, c_int64_p, c_double_p]),
    "jues_b200_t4_synth_eri": (C.c_int, [C.c_void_p, C.c_uint64, C.c_double]),
    "jues_b200_t4_create_synth": (C.c_int, [C.c_void_p, C.c_int64, C.c_uint64, C.c_double,
                                            C.POINTER(C.c_void_p)]),
    "jues_b200_get_comm_counters": (C.c_int, [C.c_void_p, c_int64_p, c_double_p]),
    "jues_b200_tei_transform_t4": (C.c_int, [C.c_void_p, C.c_void_p,
                                             c_double_p, C.c_int64, c_double_p, C.c_int64,
                                             c_double_p, C.c_int64, c_double_p, C.c_int64,
                                             C.c_int, C.POINTER(C.c_void_p)]),
    "jues_b200_rmp2_t4": (C.c_int, [C.c_void_p, C.c_void_p, c_double_p, C.c_int64,
                                    c_double_p, C.c_int64, c_double_p, c_double_p]),
    "jues_b200_rccd_t4": (C.c_int, [C.c_void_p, C.c_void_p, c_double_p, C.c_int64,
                                    c_double_p, C.c_int64, c_double_p, C.c_int, C.c_int,
                                    c_double_p, c_double_p, c_double_p]),
    "jues_b200_rccsd_t4": (C.c_int, [C.c_void_p, C.c_void_p, c_double_p, C.c_int64,
                                     c_double_p, C.c_int64, c_double_p, C.c_int,
                                     c_double_p, c_double_p, c_double_p, c_double_p]),
    "jues_b200_get_fock": (C.c_int, [C.c_void_p, c_double_p, C.c_int64, c_double_p, c_double_p, C.c_int64,
                                     c_double_p, C.c_int64, c_double_p]),
    "jues_b200_get_fock_t4": (C.c_int, [C.c_void_p, C.c_void_p, c_double_p, c_double_p, C.c_int64,
                                        c_double_p, C.c_int64, c_double_p]),
    "jues_b200_cc_default_options": (C.c_int, [C.POINTER(CCOptions)]),
    "jues_b200_auto_rccsd": (C.c_int, [C.c_void_p, c_double_p, C.c_int64, c_double_p, c_double_p, C.c_int64,
                                       C.c_int64, C.POINTER(CCOptions), c_double_p, c_double_p, c_int_p,
                                       c_int_p, c_double_p, c_double_p, c_double_p, c_double_p]),
    "jues_b200_auto_rccsd_t4": (C.c_int, [C.c_void_p, C.c_void_p, c_double_p, c_double_p, C.c_int64,
                                          C.c_int64, C.POINTER(CCOptions), c_double_p, c_double_p, c_int_p,
                                          c_int_p, c_double_p, c_double_p, c_double_p, c_double_p]),
    "jues_b200_compute_pt": (C.c_int, [C.c_void_p, c_double_p, c_double_p, c_double_p, c_double_p,
                                       c_double_p, c_double_p, c_double_p, C.c_int64, C.c_int64, c_double_p]),
    "jues_b200_mrccd": (C.c_int, [C.c_void_p, c_double_p, C.c_int64, c_double_p, C.c_int64, c_double_p,
                                  C.c_int64, c_double_p, C.c_int, c_double_p, c_int_p, c_double_p,
                                  c_double_p, c_double_p]),
    "jues_b200_mrccd_t4": (C.c_int, [C.c_void_p, C.c_void_p, c_double_p, C.c_int64, c_double_p,
                                     C.c_int64, c_double_p, C.c_int, c_double_p, c_int_p, c_double_p,
                                     c_double_p, c_double_p]),
    "jues_b200_df_rmp2": (C.c_int, [C.c_void_p, c_double_p, C.c_int64, C.c_int64, c_double_p, c_double_p, C.c_int64,
                                    c_double_p, C.c_int64, c_double_p, c_double_p]),
    "jues_b200_df_rccd": (C.c_int, [C.c_void_p, c_double_p, C.c_int64, C.c_int64, c_double_p, c_double_p, C.c_int64,
                                    c_double_p, C.c_int64, c_double_p, C.c_int, c_double_p, c_double_p, c_double_p]),
    "jues_b200_sa_ladder": (C.c_int, [C.c_void_p, c_double_p, c_double_p, C.c_int64, C.c_int64, C.c_int,
                                      c_double_p]),
    "jues_b200_init_multi": (C.c_int, [C.POINTER(C.c_void_p), C.c_int]),
    "jues_b200_group_size": (C.c_int, [C.c_void_p]),
    "jues_b200_set_trace": (C.c_int, [C.c_void_p, C.c_int]),
    "jues_b200_get_phases": (C.c_int, [C.c_void_p, C.POINTER(Phase), C.c_int]),
    "jues_b200_get_counters": (C.c_int, [C.c_void_p, c_double_p, c_int64_p, c_int64_p, c_int64_p]),
    "jues_b200_dgemm_bench": (C.c_int, [C.c_void_p, C.c_char, C.c_char, C.c_int64, C.c_int64,
                                        C.c_int64, C.c_int, c_double_p]),
    "jues_b200_transform_stress": (C.c_int, [C.c_void_p, C.c_void_p, c_double_p, C.c_int64, c_double_p, C.c_int64,
                                             c_double_p, C.c_int64, c_double_p, C.c_int64, C.c_int, c_double_p]),
    "jues_b200_dgemm_stress": (C.c_int, [C.c_void_p, C.c_char, C.c_char, C.c_int64, C.c_int64, C.c_int64,
                                         C.c_int64, C.c_int, C.POINTER(C.c_int), c_double_p]),
}

_lib = None


class LibraryMissing(RuntimeError):
    pass


def load():
    """Load libjues_b200.so and attach the prototypes.  Raises LibraryMissing (never falls
    back to a CPU implementation)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LibraryMissing(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'`"
            " (or `make -C jues.jl_b200/csrc`).  jues.jl_b200 has no CPU fallback.")
    try:
        lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    except OSError as e:  # pragma: no cover
        raise LibraryMissing(f"cannot load {LIB_PATH}: {e}") from e
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the .so lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
