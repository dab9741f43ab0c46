"""ctypes binding of ``libhgef_b200.so`` (the C-ABI declared in ``include/hgef_b200.h``).

The library is the product: there is no CPU or eager fallback.  If it has not been
built (``make -C hypergef_b200/csrc`` or ``__graft_entry__.build()``), importing an op
raises :class:`HgefBuildError` instead of silently doing something slower.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HGEF_B200_LIB") or os.path.join(_HERE, "libhgef_b200.so")   # (override: A/B builds)

HG_OK, HG_EINVAL, HG_ECUDA, HG_ENOMEM, HG_EEMPTY, HG_EGRAPH = range(6)
HG_ACCUMULATE, HG_FORCE_SCALAR, HG_TWO_PASS, HG_FORCE_FUSED, HG_FORCE_PULL, HG_FORCE_STREAM = 1, 4, 8, 16, 32, 64
HG_FORCE_RING, HG_FORCE_FSTREAM = 128, 256


class HgefBuildError(ImportError):
    pass


class HgefGraphError(ValueError):
    """An index array names a vertex / segment / hyperedge that does not exist."""


_i64, _i32, _f32, _int, _vp = C.c_int64, C.c_int32, C.c_float, C.c_int, C.c_void_p
_pi64 = C.POINTER(C.c_int64)
_pi32 = C.POINTER(C.c_int32)

# name -> argtypes; every function returns int except hg_last_error.  Kept in the order of
# include/hgef_b200.h; tests/test_boundary.py checks the two lists against each other.
SIGNATURES = {
    "hg_abi_version": [],
    "hg_device_cc": [_int],
    "hg_tune_set": [C.c_char_p, _i32, _i32],
    "hg_balance_count_host": [_i64, _vp, _i32, _pi64, _pi64],
    "hg_balance_fill_host": [_i64, _vp, _i32, _vp, _vp, _vp, _vp],
    "hg_balance_count_dev": [_i64, _vp, _i32, _pi64, _pi64, _int, _vp],
    "hg_balance_fill_dev": [_i64, _vp, _i32, _i64, _i64, _vp, _vp, _vp, _vp, _int, _vp],
    "hg_csr_build_host": [_i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _pi64],
    "hg_csr_build_dev": [_i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _pi64, _int, _vp],
    "hg_degree_scale_dev": [_i64, _vp, _vp, _f32, _int, _vp, _int, _vp],
    "hg_mtx_open": [C.c_char_p, C.POINTER(_vp), _pi64, _pi64, _pi64],
    "hg_mtx_fill": [_vp, _vp, _vp],
    "hg_mtx_close": [_vp],
    "hg_plan_create": [C.POINTER(_vp), _i64, _i64, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _int, _vp],
    "hg_plan_destroy": [_vp],
    "hg_plan_info": [_vp, _pi64, _pi64, _pi64, _pi32],
    "hg_plan_reserve": [_vp, _i32, _vp],
    "hg_plan_debug": [_vp, _pi32, _vp],
    "hg_plan_check": [_vp, _vp],
    "hg_plan_launches": [_vp, _pi64],
    "hg_aggr_forward": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _vp],
    "hg_aggr_groups": [_i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32,
                       _int, _vp],
    "hg_copy_columns": [_vp, _vp, _i64, _i64, _i64, _i64, _i32, _int, _vp],
    "hg_edge_reduce": [_i64, _vp, _vp, _vp, _vp, _vp, _i32, _int, _vp],
    "hg_edge_scatter": [_i64, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _int, _vp],
    "hg_aggr_mean": [_i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _int, _vp],
    "hg_aggr_max_forward": [_i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _int, _vp],
    "hg_aggr_max_backward": [_i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _int, _vp],
    "hg_plan_max_forward": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp],
    "hg_plan_max_backward": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp],
    "hg_plan_edge_reduce": [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp],
    "hg_plan_edge_scatter": [_vp, _vp, _vp, _vp, _i32, _vp],
    "hg_weight_grad": [_i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _int, _vp],
}

_lib = None


def lib() -> C.CDLL:
    """The loaded library; raises HgefBuildError when it was never built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise HgefBuildError(
                f"{LIB_PATH} is missing: build it with `make -C hypergef_b200/csrc` "
                "(or `python -c 'import __graft_entry__ as g; g.build()'`). "
                "hypergef_b200 has no CPU fallback.")
        handle = C.CDLL(LIB_PATH)
        handle.hg_last_error.restype = C.c_char_p
        handle.hg_last_error.argtypes = []
        for name, argtypes in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = C.c_int
            fn.argtypes = argtypes
        _lib = handle
    return _lib


_EXC = {HG_EINVAL: ValueError, HG_ECUDA: RuntimeError, HG_ENOMEM: MemoryError,
        HG_EEMPTY: IndexError, HG_EGRAPH: HgefGraphError}


def check(rc: int) -> None:
    """Turn a non-zero return code into the Python exception the header documents."""
    if rc != HG_OK:
        msg = lib().hg_last_error().decode("utf-8", "replace")
        raise _EXC.get(rc, RuntimeError)(msg)


def call(name: str, *args) -> None:
    check(getattr(lib(), name)(*args))
