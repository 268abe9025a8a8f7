"""ctypes binding of libmclcar_b200.so (the C ABI declared in include/mclcar_b200.h).

This is the Python stand-in for the `.Call` glue of r/mclcar_b200_glue.c: argument
unpacking only.  There is no fallback: if the shared library is missing the import
fails, and compute calls on a machine without a B200 return MCLCAR_ENODEV.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "lib" / "libmclcar_b200.so"

OK, EINVAL, ERHO, ECUDA, ENOMEM, ESTATE, ENODEV, ENCCL, ELIMIT = range(9)
COL_IS_SAMPLE, ROW_IS_SAMPLE = 0, 1
F64, F32 = 0, 1
SHIFT_MEAN, SHIFT_MAX = 0, 1
WHAT_LR, WHAT_GRAD, WHAT_HESS = 0, 1, 2
MAX_P, MAX_COV = 12, 8
FAMILY_DCAR, FAMILY_BINOM, FAMILY_POISSON, FAMILY_HCAR = 0, 1, 2, 3


class MclcarError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"mclcar_b200 error {code}: {msg}")
        self.code = code


class SamplerOpts(C.Structure):
    _fields_ = [("n_chains", C.c_int64), ("burn", C.c_int32), ("thin", C.c_int32), ("keep", C.c_int32),
                ("dtype", C.c_int32), ("reserved", C.c_int32 * 2), ("omega", C.c_double)]


def _load() -> C.CDLL:
    if not LIB_PATH.exists():
        # built in-tree by `python -m mclcar_b200.build` / __graft_entry__.build(); never silently absent
        if os.environ.get("MCLCAR_B200_AUTOBUILD", "1") == "1":
            from . import build as _build
            _build.build()
        if not LIB_PATH.exists():
            raise ImportError(f"{LIB_PATH} is missing: run `python -m mclcar_b200.build` (needs nvcc)")
    return C.CDLL(str(LIB_PATH))


lib = _load()

_vp, _i, _i64, _d = C.c_void_p, C.c_int, C.c_int64, C.c_double
_pd = C.POINTER(C.c_double)
_pi = C.POINTER(C.c_int)
_pi64 = C.POINTER(C.c_int64)

# name -> (restype, argtypes); mirrors include/mclcar_b200.h one to one
SIGNATURES = {
    "mclcar_ctx_create": (_i, [_i, C.c_uint64, C.POINTER(_vp)]),
    "mclcar_ctx_destroy": (None, [_vp]),
    "mclcar_last_error": (C.c_char_p, [_vp]),
    "mclcar_ctx_set_stream": (_i, [_vp, _vp]),
    "mclcar_ctx_sync": (_i, [_vp]),
    "mclcar_ctx_trim": (_i, [_vp]),
    "mclcar_ctx_set_shard": (_i, [_vp, _i, _i]),
    "mclcar_ctx_set_draw": (_i, [_vp, C.c_uint64]),
    "mclcar_ctx_get_draw": (_i, [_vp, C.POINTER(C.c_uint64)]),
    "mclcar_ctx_profile": (_i, [_vp, _i]),
    "mclcar_ctx_profile_read": (_i, [_vp, _pd, _pi64]),
    "mclcar_nccl_unique_id": (_i, [_vp, C.c_char_p]),
    "mclcar_nccl_comm_init": (_i, [_vp, C.c_char_p]),
    "mclcar_graph_torus": (_i, [_vp, _i, _i]),
    "mclcar_graph_csr": (_i, [_vp, _i, _pi, _pi, _pd]),
    "mclcar_graph_csr_opts": (_i, [_vp, _i, _pi, _pi, _pd, _i]),
    "mclcar_graph_kind": (_i, [_vp, _pi, _pi]),
    "mclcar_graph_set_eigenvalues": (_i, [_vp, _pd, _i]),
    "mclcar_graph_estimate_spectrum": (_i, [_vp, _i, C.c_uint64]),
    "mclcar_graph_spectrum_kind": (_i, [_vp, _pi]),
    "mclcar_set_design": (_i, [_vp, _pd, _i]),
    "mclcar_set_obs": (_i, [_vp, _pd, _pd]),
    "mclcar_graph_n": (_i, [_vp]),
    "mclcar_graph_ncolors": (_i, [_vp]),
    "mclcar_samples_from_host": (_i, [_vp, _vp, _i, _i, _i64, _i64, _i64, C.POINTER(_vp)]),
    "mclcar_samples_from_device": (_i, [_vp, _vp, _i, _i, _i64, _i64, _i64, C.POINTER(_vp)]),
    "mclcar_samples_free": (None, [_vp]),
    "mclcar_samples_count": (_i64, [_vp]),
    "mclcar_samples_fetch": (_i, [_vp, _vp, _i, _pd, _i64]),
    "mclcar_samples_diag": (_i, [_vp, _vp, _pd, _pi64]),
    "mclcar_samples_schedule": (_i, [_vp, _vp, _pi64, _pi, _pi, _pd]),
    "mclcar_ess_from_stats": (_i, [_pd, _i64, _i, _i64, _i64, _pd]),
    "mclcar_samples_ess": (_i, [_vp, _vp, _pd]),
    "mclcar_dcar_stats": (_i, [_vp, _vp, _pd]),
    "mclcar_dcar_prep": (_i, [_vp, _pd, _i, _i64, C.POINTER(SamplerOpts), C.POINTER(_vp)]),
    "mclcar_dcar_plan": (_i, [_vp, _d, _i64, C.POINTER(SamplerOpts), _pi, _pi64, _pi, _pi, _pd]),
    "mclcar_dcar_attach": (_i, [_vp, _vp, _pd, _i, _pd]),
    "mclcar_dcar_t10_t20": (_i, [_vp, _vp, _pd, _pd]),
    "mclcar_dcar_eval": (_i, [_vp, _vp, _pd, _i, _i, _pd, _pi64, _pi64, _i, _pd, _pd]),
    "mclcar_dcar_grad": (_i, [_vp, _vp, _pd, _i, _i, _i, _i, _pd, _pd]),
    "mclcar_dcar_hess": (_i, [_vp, _vp, _pd, _i, _i, _i, _pd]),
    "mclcar_dcar_mcmle_var": (_i, [_vp, _vp, _pd, _i, _i, _pd]),
    "mclcar_dcar_sigmabeta": (_i, [_vp, _pd, _i, _pd]),
    "mclcar_dcar_profile_eval": (_i, [_vp, _vp, _pd, _i, _pd, _i, _pd, _pd]),
    "mclcar_dcar_loglik": (_i, [_vp, _pd, _i, _i, _pd, _pd]),
    "mclcar_dcar_ploglik": (_i, [_vp, _pd, _i, _pd]),
    "mclcar_dcar_hessian_exact": (_i, [_vp, _pd, _i, _pd]),
    "mclcar_dcar_surface": (_i, [_vp, _vp, _pd, _i, _pd, _i, _pd, _i, _pd, _i, _pd, _pd]),
    "mclcar_glm_surface": (_i, [_vp, _vp, _pd, _i, _pd, _i, _pd, _i, _pd, _pd]),
    "mclcar_samples_stats_export": (_i, [_vp, _vp, _pi, _pd]),
    "mclcar_samples_from_stats": (_i, [_vp, _pd, _i64, _i, _i64, _i64, C.POINTER(_vp)]),
    "mclcar_hcar_stats": (_i, [_vp, _vp, _pd]),
    "mclcar_glm_prep": (_i, [_vp, _i, _pd, _i, _i64, C.POINTER(SamplerOpts), C.POINTER(_vp)]),
    "mclcar_glm_attach": (_i, [_vp, _vp, _i, _pd, _i, _pd]),
    "mclcar_glm_lhzy": (_i, [_vp, _vp, _pd]),
    "mclcar_glm_eval": (_i, [_vp, _vp, _pd, _i, _i, _pi64, _pi64, _i, _pd, _pd]),
    "mclcar_glm_grad": (_i, [_vp, _vp, _pd, _i, _i, _i, _i, _pd, _pd]),
    "mclcar_glm_hess": (_i, [_vp, _vp, _pd, _i, _i, _i, _pd]),
    "mclcar_glm_softplus_table_eval": (_i, [_pd, _i64, _pd, _pd]),
    "mclcar_glm_get_beta": (_i, [_vp, _vp, _pd, _pd, _i, C.c_double, _pd, _pd, _pi, _pi]),
    "mclcar_hcar_set_model": (_i, [_vp, _i, _pi, _pi, _pd, _pi, _pi, _pd, _pd, _i]),
    "mclcar_hcar_prep": (_i, [_vp, _pd, _i, _i64, C.POINTER(SamplerOpts), C.POINTER(_vp)]),
    "mclcar_hcar_attach": (_i, [_vp, _vp, _pd, _i, _pd]),
    "mclcar_hcar_lpsi": (_i, [_vp, _vp, _pd, _pd]),
    "mclcar_hcar_eval": (_i, [_vp, _vp, _pd, _i, _i, _pd, _pd, _pi]),
    "mclcar_hcar_grad": (_i, [_vp, _vp, _pd, _i, _i, _i, _i, _pd, _pd]),
    "mclcar_hcar_hess": (_i, [_vp, _vp, _pd, _i, _i, _i, _pd]),
    "mclcar_multi_dcar_prep": (_i, [C.POINTER(_vp), _i, _pd, _i, _i64, C.POINTER(SamplerOpts), C.POINTER(_vp)]),
    "mclcar_multi_dcar_run": (_i, [C.POINTER(_vp), C.POINTER(_vp), _i, _pd, _i, _i, _i, _i, _pd, _i, _pd, _pd]),
    "mclcar_multi_glm_prep": (_i, [C.POINTER(_vp), _i, _i, _pd, _i, _i64, C.POINTER(SamplerOpts), C.POINTER(_vp)]),
    "mclcar_multi_glm_run": (_i, [C.POINTER(_vp), C.POINTER(_vp), _i, _pd, _i, _i, _i, _i, _i, _pd, _pd]),
    "mclcar_multi_hcar_prep": (_i, [C.POINTER(_vp), _i, _pd, _i, _i64, C.POINTER(SamplerOpts), C.POINTER(_vp)]),
    "mclcar_multi_hcar_run": (_i, [C.POINTER(_vp), C.POINTER(_vp), _i, _pd, _i, _i, _i, _i, _i, _pd, _pd, _pi]),
    "mclcar_tuple_len": (_i64, [_i, _i, _i]),
    "mclcar_dcar_partial": (_i, [_vp, _vp, _pd, _i, _i, _i, _pi64, _pi64, _i, _pd]),
    "mclcar_tuples_merge": (_i, [_i, _i, _i, _i, _i, _pd, _pd]),
    "mclcar_dcar_finish": (_i, [_vp, _pd, _i, _pd, _i, _i, _i, _i, _pd, _pd, _pd, _pd]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)  # AttributeError here = header/library mismatch: fail loudly
    _fn.restype = _res
    _fn.argtypes = _args


def dptr(a):
    """double* of a C-contiguous float64 array (None -> NULL)."""
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_pd)


def i64ptr(a):
    if a is None:
        return None
    assert a.dtype == np.int64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_pi64)


def iptr(a):
    assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_pi)


def check(status: int, ctx=None):
    if status != OK:
        msg = lib.mclcar_last_error(ctx)
        raise MclcarError(status, msg.decode() if msg else "?")
