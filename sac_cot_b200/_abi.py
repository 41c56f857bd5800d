"""ctypes transcription of include/sac_cot.h.

`bind(cdll)` attaches argument/return types to every entry point the header declares and
returns the handle; it is used for the product library (sac_cot_b200/lib/libsaccot.so) and —
by tests/ and bench.py's CPU-baseline leg only — for the oracle, which exports the same ABI.
The reference itself has no FFI to mirror (/root/reference/README.md:1-2 is the whole repo).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

MAX_N = 65535
MAX_EDGES = 4096
MAX_APEX = 8
MAX_HYPOTHESES = 32768
MAX_DESC_DIM = 256
MAX_KEYPOINTS = 1048576

OK = 0
E_NULL, E_SIZE, E_PARAMS, E_NODEVICE, E_UNSUPPORTED, E_WHICH, E_CAPACITY, E_NOMEM, E_COMM = -1, -2, -3, -4, -5, -6, -7, -8, -9
COMM_ID_BYTES = 128

SCORE_INLIER_COUNT = 0
SCORE_TRUNCATED_RESIDUAL = 1
LOC_HOST, LOC_DEVICE = 0, 1

DBG_ADJ, DBG_T_NODE, DBG_NUM_EDGES, DBG_EDGE_KEYS, DBG_TOP_EDGES, DBG_TRIANGLES, DBG_HYP_RT, \
    DBG_HYP_SCORE, DBG_BEST_KEY, DBG_MASK, DBG_HIST, DBG_ADJ_FIRST = range(12)
COMPAT_FIRST_ORDER, COMPAT_SECOND_ORDER = 0, 1
PARAMS_SIZE_V1 = 32

DBG_DTYPES = {
    DBG_ADJ: np.uint32, DBG_T_NODE: np.uint32, DBG_NUM_EDGES: np.uint64, DBG_EDGE_KEYS: np.uint64,
    DBG_TOP_EDGES: np.uint64, DBG_TRIANGLES: np.int32, DBG_HYP_RT: np.float32,
    DBG_HYP_SCORE: np.uint64, DBG_BEST_KEY: np.uint64, DBG_MASK: np.uint32, DBG_HIST: np.uint32,
    DBG_ADJ_FIRST: np.uint32,
}


class Params(C.Structure):
    """struct sac_cot_params (include/sac_cot.h)."""
    _fields_ = [
        ("struct_size", C.c_uint32),
        ("tau_compat", C.c_float),
        ("tau_inlier", C.c_float),
        ("num_edges", C.c_int32),
        ("apex_per_edge", C.c_int32),
        ("score_mode", C.c_int32),
        ("refit", C.c_int32),
        ("compat_mode", C.c_int32),     # version 1 of the struct (32 bytes) ends here
        ("so_min_common", C.c_int32),
        ("reserved", C.c_int32),
    ]


_f32p = C.POINTER(C.c_float)
_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)
_u64p = C.POINTER(C.c_uint64)
_ctxp = C.c_void_p

# name -> (restype, argtypes): exactly the symbols include/sac_cot.h declares
SYMBOLS = {
    "sac_cot_params_default": (C.c_int, [C.POINTER(Params)]),
    "sac_cot_register": (C.c_int, [_f32p, _f32p, C.c_int32, C.POINTER(Params), _f32p, _f32p, _i32p]),
    "sac_cot_ctx_create": (C.c_int, [C.POINTER(_ctxp), C.c_int32, C.c_void_p]),
    "sac_cot_ctx_destroy": (C.c_int, [_ctxp]),
    "sac_cot_ctx_set": (C.c_int, [_ctxp, C.c_char_p, C.c_int64]),
    "sac_cot_ctx_get": (C.c_int, [_ctxp, C.c_char_p, _i64p]),
    "sac_cot_register_batch": (C.c_int, [_ctxp, C.POINTER(_f32p), C.POINTER(_f32p), _i32p, C.c_int32,
                                         C.POINTER(Params), _f32p, _f32p, _i32p]),
    "sac_cot_register_packed": (C.c_int, [_ctxp, C.c_void_p, C.c_void_p, _i64p, C.c_int32,
                                          C.POINTER(Params), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]),
    "sac_cot_match_packed": (C.c_int, [_ctxp, C.c_void_p, C.c_void_p, _i64p, C.c_void_p, C.c_void_p, _i64p, C.c_int32,
                                       C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]),
    "sac_cot_match_mutual": (C.c_int, [_ctxp, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, _i64p, _i64p, C.c_int32,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]),
    "sac_cot_match": (C.c_int, [_f32p, _f32p, C.c_int32, _f32p, _f32p, C.c_int32, C.c_int32, _i32p, _f32p, _f32p]),
    "sac_cot_group_create": (C.c_int, [C.POINTER(_ctxp), _i32p, C.c_int32]),
    "sac_cot_group_destroy": (C.c_int, [_ctxp]),
    "sac_cot_group_size": (C.c_int32, [_ctxp]),
    "sac_cot_group_ctx": (_ctxp, [_ctxp, C.c_int32]),
    "sac_cot_group_set": (C.c_int, [_ctxp, C.c_char_p, C.c_int64]),
    "sac_cot_group_register_packed": (C.c_int, [_ctxp, C.c_void_p, C.c_void_p, _i64p, C.c_int32,
                                                C.POINTER(Params), C.c_void_p, C.c_void_p, C.c_void_p]),
    "sac_cot_comm_unique_id": (C.c_int, [C.c_void_p]),
    "sac_cot_ctx_comm_init": (C.c_int, [_ctxp, C.c_void_p, C.c_int32, C.c_int32]),
    "sac_cot_ctx_set_comm": (C.c_int, [_ctxp, C.c_void_p, C.c_int32, C.c_int32]),
    "sac_cot_register_sharded": (C.c_int, [_ctxp, C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(Params),
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]),
    "sac_cot_sharded_phase1": (C.c_int, [_ctxp, _f32p, _f32p, C.c_int32, C.POINTER(Params), C.c_int32,
                                         C.c_int32, _u64p, _u64p]),
    "sac_cot_sharded_phase2": (C.c_int, [_ctxp, _u64p, _u64p, _u64p]),
    "sac_cot_sharded_phase3": (C.c_int, [_ctxp, C.c_uint64, _f32p, _f32p, _i32p]),
    "sac_cot_debug_get": (C.c_int, [_ctxp, C.c_int32, C.c_int32, C.c_void_p, C.c_size_t,
                                    C.POINTER(C.c_size_t)]),
    "sac_cot_strerror": (C.c_char_p, [C.c_int]),
    "sac_cot_version": (C.c_char_p, []),
}


# entry points that make the library bind NCCL (dlopen("libnccl.so.2") on first use)
_NCCL_BINDING = ("sac_cot_comm_unique_id", "sac_cot_ctx_comm_init", "sac_cot_ctx_set_comm")


def _prefer_torch_nccl():
    """torch ships its own libnccl.so.2 under the same soname as a system copy, and a process only ever holds ONE
    library per soname: whichever is loaded first serves everybody.  If this library's dlopen brought in an older system
    copy first, a later `import torch` dies on the symbols it misses (seen: system 2.27.3 against torch's 2.28.9,
    `undefined symbol: ncclDevCommCreate`).  So a Python process that has torch installed loads torch's copy before the
    library binds NCCL; a process without torch (or a C/C++ host) gets the system copy, or the file named by the
    environment variable SAC_COT_NCCL_LIB."""
    import sys

    if "torch" in sys.modules:
        return
    try:
        import torch  # noqa: F401
    except Exception:  # noqa: BLE001 - no torch: the system copy is the right one
        pass


def bind(lib: C.CDLL) -> C.CDLL:
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    for name in _NCCL_BINDING:
        raw = getattr(lib, name)
        if getattr(raw, "_sac_cot_guarded", False):
            continue

        def guarded(*a, _raw=raw):
            _prefer_torch_nccl()
            return _raw(*a)

        guarded._sac_cot_guarded = True
        guarded.restype, guarded.argtypes = raw.restype, raw.argtypes  # what tests/test_library_abi.py reads
        setattr(lib, name, guarded)
    return lib


def fptr(a: np.ndarray):
    return a.ctypes.data_as(_f32p)


def default_params(lib: C.CDLL, **overrides) -> Params:
    p = Params()
    rc = lib.sac_cot_params_default(C.byref(p))
    if rc != OK:
        raise RuntimeError(f"sac_cot_params_default -> {rc}")
    for k, v in overrides.items():
        if not hasattr(p, k):
            raise TypeError(f"unknown sac_cot_params field {k!r}")
        setattr(p, k, v)
    return p
