"""ctypes binding of libenf_b200.so (include/enf_b200.h).

The library is the product; this module only declares its C signatures.  If the
shared object is missing the import of any compute entry point fails loudly --
there is no Python/numpy fallback.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# ENF_B200_LIB: load an alternative build of the library (tuning experiments)
LIB_PATH = os.environ.get("ENF_B200_LIB") or os.path.join(_HERE, "libenf_b200.so")

ENF_F32, ENF_F64 = 0, 1
(ENF_CENTER_STRETCH, ENF_CENTER_CONTRACT, ENF_JOHNSON, ENF_JOHNSON_INV,
 ENF_SCALE_SHIFT, ENF_HOUSEHOLDER) = range(6)
ENF_NEGLL_ZYGOTE_PRIMAL = 1
ENF_UNIQUE_ID_BYTES = 128

# every symbol include/enf_b200.h declares: (name, restype, argtypes)
_vp, _i, _i64, _u64, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_size_t
_pvp = C.POINTER(C.c_void_p)


class enf_op(C.Structure):
    _fields_ = [("kind", C.c_int32), ("K", C.c_int32), ("params", C.c_void_p)]


class enf_target(C.Structure):
    _fields_ = [("kind", C.c_int32), ("K", C.c_int32), ("weights", C.POINTER(C.c_double)), ("means", C.POINTER(C.c_double)),
                ("sigmas", C.POINTER(C.c_double))]


ENF_TARGET_GAUSS_MIXTURE = 1
(ENF_JSU_PDF, ENF_JSU_LOGPDF, ENF_JSU_CDF, ENF_JSU_LOGCDF, ENF_JSU_CCDF, ENF_JSU_LOGCCDF, ENF_JSU_QUANTILE) = range(7)

SYMBOLS = [
    ("enf_init", _i, [_i, _pvp]),
    ("enf_destroy", _i, [_vp]),
    ("enf_last_error", C.c_char_p, [_vp]),
    ("enf_device_count", _i, [C.POINTER(_i)]),
    ("enf_sync", _i, [_vp]),
    ("enf_alloc", _i, [_vp, _sz, _pvp]),
    ("enf_free", _i, [_vp, _vp]),
    ("enf_host_alloc", _i, [_vp, _sz, _pvp]),
    ("enf_host_free", _i, [_vp, _vp]),
    ("enf_h2d", _i, [_vp, _vp, _vp, _sz]),
    ("enf_d2h", _i, [_vp, _vp, _vp, _sz]),
    ("enf_memset", _i, [_vp, _vp, _i, _sz]),
    ("enf_fill_normal", _i, [_vp, _i, _vp, _i, _i64, _i64, _u64]),
    ("enf_convert", _i, [_vp, _i, _vp, _i, _vp, _i64]),
    ("enf_chain_create", _i, [_vp, _i, _i, _i, C.POINTER(enf_op), _pvp]),
    ("enf_chain_set_params", _i, [_vp, _vp]),
    ("enf_chain_num_params", _i, [_vp, C.POINTER(_i64)]),
    ("enf_chain_destroy", _i, [_vp]),
    ("enf_forward", _i, [_vp, _vp, _i64, _vp]),
    ("enf_forward_ladj", _i, [_vp, _vp, _i64, _vp, _vp]),
    ("enf_forward_ladj_host", _i, [_vp, _vp, _i64, _vp, _vp]),
    ("enf_negll", _i, [_vp, _vp, _i64, C.POINTER(C.c_double)]),
    ("enf_negll_grad", _i, [_vp, _vp, _i64, _i, C.POINTER(C.c_double), _vp]),
    ("enf_negll_grad_partial", _i, [_vp, _vp, _i64, _pvp, C.POINTER(_i64)]),
    ("enf_negll_grad_finish", _i, [_vp, _vp, _i64, _i, C.POINTER(C.c_double), _vp]),
    ("enf_elbo_grad", _i, [_vp, C.POINTER(enf_target), _vp, _i64, _i, C.POINTER(C.c_double), _vp]),
    ("enf_johnsonsu", _i, [_vp, _i, _i, C.POINTER(C.c_double), _vp, _i64, _vp]),
    ("enf_group_unique_id", _i, [_vp]),
    ("enf_group_init", _i, [_vp, _i, _i, _vp]),
    ("enf_group_destroy", _i, [_vp]),
    ("enf_negll_grad_group", _i, [_vp, _vp, _i64, _i, C.POINTER(C.c_double), _vp]),
    ("enf_group_allreduce_sums", _i, [_vp, _i64]),
    ("enf_optimize_whitening", _i, [_vp, _vp, _i64, _i64, _i64, C.c_double, C.c_double, _i, _i, _i, _vp, _vp, _vp, C.POINTER(_i64)]),
    ("enf_optimize_whitening_batches", _i, [_vp, _vp, _i64, C.POINTER(_i64), _i64, C.c_double, C.c_double, _i, _i, _i, _vp, _vp, _vp, C.POINTER(_i64)]),
    ("enf_event_record", _i, [_vp, _i]),
    ("enf_event_elapsed_ms", _i, [_vp, _i, _i, C.POINTER(C.c_float)]),
    ("enf_launch_count", _i, [_vp, C.POINTER(_i64)]),
    ("enf_chain_describe", _i, [_vp, C.c_char_p, _sz]),
    ("enf_version", _i, []),
]

_lib = None


class EnfError(RuntimeError):
    """A non-zero status from libenf_b200 (the Julia shim raises the same way)."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"libenf_b200 error {code}: {msg}")
        self.code = code


def lib() -> C.CDLL:
    """Load libenf_b200.so and bind every declared symbol."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: build it with `make -C {os.path.join(_HERE, 'csrc')}` "
                "(or `python -c 'import __graft_entry__ as g; g.build()'`). "
                "There is no CPU fallback for the trafo-chain path.")
        l = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, res, args in SYMBOLS:
            fn = getattr(l, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc: int, ctx=None) -> None:
    if rc != 0:
        msg = lib().enf_last_error(ctx)
        raise EnfError(rc, msg.decode() if msg else "unknown error")
