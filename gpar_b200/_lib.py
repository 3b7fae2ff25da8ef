"""ctypes binding of libgpar_b200.so (include/gpar_b200.h).

There is NO CPU fallback: if the shared object is missing or a call fails the
product path raises.  torch is imported first so that the CUDA runtime the
library links against (libcudart.so.12) is the one torch already loaded; torch
tensors are used for device memory and streams only.
"""
import ctypes as C
import os

import torch  # noqa: F401  (loads libcudart before the CDLL below)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GPAR_B200_LIB") or os.path.join(_HERE, "libgpar_b200.so")  # env override: kernel experiments only

TILE = 128
MAX_TERMS = 8
MAX_FEATS = 96
GRAD_NP = 2 * MAX_TERMS + 2 * MAX_FEATS + 1
TERM_EQ, TERM_RQ, TERM_LINEAR, TERM_CONST = 0, 1, 2, 3
FEAT_SCALE, FEAT_SIN, FEAT_COS = 0, 1, 2


class Term(C.Structure):
    _fields_ = [
        ("type", C.c_int32),
        ("f_begin", C.c_int32),
        ("f_end", C.c_int32),
        ("_pad", C.c_int32),
        ("variance", C.c_double),
        ("alpha", C.c_double),
    ]


class KernelSpec(C.Structure):
    _fields_ = [
        ("n_terms", C.c_int32),
        ("n_feats", C.c_int32),
        ("terms", Term * MAX_TERMS),
        ("feat_col", C.c_int32 * MAX_FEATS),
        ("feat_op", C.c_int32 * MAX_FEATS),
        ("feat_a", C.c_double * MAX_FEATS),
        ("feat_b", C.c_double * MAX_FEATS),
    ]


class GparError(RuntimeError):
    pass


_i64, _p, _d, _int = C.c_int64, C.c_void_p, C.c_double, C.c_int
_SPEC = C.POINTER(KernelSpec)

#: name -> (restype, argtypes); mirrors include/gpar_b200.h one to one.
SIGNATURES = {
    "gpar_abi_version": (_int, []),
    "gpar_last_error": (C.c_char_p, []),
    "gpar_gram": (_int, [_SPEC, _p, _i64, _i64, _p, _i64, _i64, _p, _d, _int, _p, _i64, _p]),
    "gpar_gram_batched": (_int, [_SPEC, _p, _i64, _i64, _i64, _p, _i64, _i64, _i64, _p, _i64, _d, _int, _p, _i64,
                                 _i64, _i64, _p]),
    "gpar_potrf_workspace_bytes": (C.c_size_t, [_i64, _i64, _i64]),
    "gpar_trsm_rows_scratch_bytes": (C.c_size_t, [_i64]),
    "gpar_potrf": (_int, [_p, _i64, _i64, _i64, _p, _i64, _i64, _i64, _i64, _p, _p, _p]),
    "gpar_potrf_multi_reset": (_int, [_p, _i64, _i64, _p, _p]),
    "gpar_potrf_multi": (_int, [_p, _i64, _i64, _p, _i64, _i64, _p, _p, _int, _int, _p, _p]),
    "gpar_ipc_alloc": (_int, [C.c_size_t, C.POINTER(C.c_void_p)]),
    "gpar_ipc_free": (_int, [_p]),
    "gpar_ipc_export": (_int, [_p, C.c_char_p]),
    "gpar_ipc_open": (_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "gpar_ipc_close": (_int, [_p]),
    "gpar_trsm_rows": (_int, [_p, _i64, _i64, _p, _p, _i64, _i64, _p, _p]),
    "gpar_syrk_sub": (_int, [_p, _i64, _i64, _i64, _p, _i64, _i64, _i64, _i64, _p]),
    "gpar_syrk_add": (_int, [_p, _i64, _i64, _i64, _p, _i64, _i64, _i64, _i64, _p]),
    "gpar_potri_scratch_bytes": (C.c_size_t, [_i64]),
    "gpar_potri": (_int, [_p, _i64, _i64, _p, _p, _i64, _p, _i64, _p, _p]),
    "gpar_gram_grad_workspace_bytes": (C.c_size_t, [_i64]),
    "gpar_gram_wgrad_workspace_bytes": (C.c_size_t, [_i64, _i64]),
    "gpar_gram_wgrad": (_int, [_SPEC, _p, _i64, _i64, _p, _i64, _i64, _p, _i64, _p, _p, _p, _p, _p, _p]),
    "gpar_row_sqnorm": (_int, [_p, _i64, _i64, _i64, _p, _p]),
    "gpar_gram_grad": (_int, [_SPEC, _p, _i64, _i64, _p, _p, _i64, _p, _p, _p, _p]),
    "gpar_transpose_scale": (_int, [_p, _i64, _i64, _i64, _p, _p, _i64, _p]),
    "gpar_vfe_rowterms": (_int, [_SPEC, _p, _i64, _i64, _p, _i64, _i64, _p, _p, _p, _p, _i64, _i64, _p, _p, _p]),
    "gpar_gemm_nt": (_int, [_p, _i64, _i64, _i64, _p, _i64, _p, _i64, _i64, _int, _p]),
    "gpar_axpy": (_int, [_i64, _d, _p, _p, _p]),
    "gpar_untransform": (_int, [_p, _i64, _i64, _p, _p, _int, _p]),
    "gpar_backsolve": (_int, [_p, _i64, _i64, _p, _p, _p, _p, _p]),
    "gpar_logdet_quad": (_int, [_p, _i64, _i64, _p, _p, _p]),
    "gpar_gemv": (_int, [_p, _i64, _i64, _i64, _p, _p, _p]),
    "gpar_gram_gemv": (_int, [_SPEC, _p, _i64, _i64, _p, _i64, _i64, _p, _p, _p]),
    "gpar_sample_affine": (_int, [_p, _i64, _i64, _i64, _p, _p, _i64, _p, _p, _i64, _i64, _p, _p]),
    "gpar_gather_rows": (_int, [_p, _i64, _p, _i64, _i64, _p, _i64, _p]),
    "gpar_scatter_col": (_int, [_p, _i64, _i64, _p, _p, _i64, _p]),
    "gpar_mean_identity": (_int, [_p, _p, _d, _p, _i64, _p, _p]),
    "gpar_mean_axis0": (_int, [_p, _i64, _i64, _p, _p]),
    "gpar_sum_axis0_add": (_int, [_p, _i64, _i64, _p, _p]),
    "gpar_percentile2_axis0": (_int, [_p, _i64, _i64, _i64, _d, _i64, _d, _p, _p, _p]),
}

#: diagnostics of include/gpar_b200_debug.h -- not part of the drop-in ABI.  Hooks into the dataflow kernel are
#: exported by the product library; the raw probes live in libgpar_b200_debug.so.
DEBUG_HOOKS = {
    "gpar_debug_set_dataflow_prof": (_int, [_p]),
    "gpar_debug_decode_ticket": (_int, [_i64, _i64, _i64, _i64, _i64, C.POINTER(C.c_int32)]),
    "gpar_debug_diag_profile": (_int, [_p, _i64, _i64, _p, _p, _p, _p]),
    "gpar_debug_trsm_row_plan": (_int, [_i64, _int, C.POINTER(C.c_int64)]),
}
DEBUG_PROBES = {
    "gpar_debug_latency_probe": (_int, [_p, _p]),
    "gpar_fp64_probe": (_int, [_int, _i64, _p, C.POINTER(C.c_double), _p]),
}
DEBUG_LIB_PATH = os.path.join(_HERE, "libgpar_b200_debug.so")
_dbg = None


def load_debug():
    """The probes library (bench.py's issue-rate probes, scripts/prof_latency.py)."""
    global _dbg
    if _dbg is None:
        if not os.path.exists(DEBUG_LIB_PATH):
            raise GparError(f"{DEBUG_LIB_PATH} is missing: build it with `python -m gpar_b200.build`")
        lib = C.CDLL(DEBUG_LIB_PATH)
        for name, (res, args) in DEBUG_PROBES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _dbg = lib
    return _dbg

_lib = None
#: number of kernel-launching C-ABI calls made through this module (bench.py's gpu_launches
#: is derived from per-call launch counts reported by the wrappers in engine.py).
call_count = 0


def load():
    """Load the shared library (once) and bind every symbol of the header."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GparError(
            f"{LIB_PATH} is missing: build it with `python -m gpar_b200.build` "
            "(there is no CPU fallback for the GPAR hot path)."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in list(SIGNATURES.items()) + list(DEBUG_HOOKS.items()):
        fn = getattr(lib, name)  # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    if lib.gpar_abi_version() != 1:
        raise GparError("libgpar_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().gpar_last_error().decode(errors="replace")
        raise GparError(f"{what} failed with code {rc}: {msg}")


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else C.c_void_p(t.data_ptr())
