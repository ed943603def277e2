"""ctypes binding of libqilcuda.so (include/qilcuda.h).  No CPU fallback: a missing library or a
missing GPU raises."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libqilcuda.so")


class ArgumentError(ValueError):
    """Julia ArgumentError (QIL_ERR_ARGUMENT)."""


class DomainError(ValueError):
    """Julia DomainError (QIL_ERR_DOMAIN)."""


class ErrorException(RuntimeError):
    """Julia ErrorException (QIL_ERR_RUNTIME)."""


class CudaError(RuntimeError):
    """CUDA runtime/driver failure (QIL_ERR_CUDA)."""


class UnsupportedError(NotImplementedError):
    """Shape outside what the library implements (QIL_ERR_UNSUPPORTED); never a silent fallback."""


_ERRMAP = {1: ArgumentError, 2: DomainError, 3: ErrorException, 4: AssertionError, 5: CudaError,
           6: UnsupportedError}

_lib = None

c_ctx = C.c_void_p
c_mps = C.c_void_p
c_mpo = C.c_void_p
i64p = C.POINTER(C.c_int64)

# name -> (argtypes, ) ; every function returns int status except the two string getters
_SIGS = {
    "qil_create": [C.c_int, C.POINTER(c_ctx)],
    "qil_create_on_stream": [C.c_int, C.c_void_p, C.POINTER(c_ctx)],
    "qil_destroy": [c_ctx],
    "qil_sync": [c_ctx],
    "qil_launch_count": [c_ctx, C.POINTER(C.c_uint64)],
    "qil_truncation_margin": [c_ctx, C.c_int, C.POINTER(C.c_double)],
    "qil_profile_enable": [c_ctx, C.c_int],
    "qil_profile_reset": [c_ctx],
    "qil_profile_read": [c_ctx, C.c_int, C.POINTER(C.c_double), i64p],
    "qil_profile_read_work": [c_ctx, C.c_int, C.POINTER(C.c_double), i64p, C.POINTER(C.c_double),
                              C.POINTER(C.c_double)],
    "qil_mps_from_host": [c_ctx, C.c_int, C.c_int, i64p, C.POINTER(C.c_void_p), C.c_double, C.POINTER(c_mps)],
    "qil_mps_info": [c_mps, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_double)],
    "qil_mps_dims": [c_mps, i64p],
    "qil_mps_get_core": [c_mps, C.c_int, C.c_void_p],
    "qil_mps_get_cores": [c_mps, C.c_void_p, C.c_int64],
    "qil_mps_set_amplitude": [c_mps, C.c_double],
    "qil_mps_clone": [c_mps, C.POINTER(c_mps)],
    "qil_mps_free": [c_mps],
    "qil_mpo_from_host": [c_ctx, C.c_int, C.c_int, i64p, C.POINTER(C.c_void_p), C.POINTER(c_mpo)],
    "qil_mpo_info": [c_mpo, C.POINTER(C.c_int), C.POINTER(C.c_int)],
    "qil_mpo_dims": [c_mpo, i64p],
    "qil_mpo_get_core": [c_mpo, C.c_int, C.c_void_p],
    "qil_mpo_free": [c_mpo],
    "qil_coefficient_batch": [c_ctx, c_mps, C.c_void_p, C.c_int64, C.c_void_p],
    "qil_coefficient_batch_dev": [c_ctx, c_mps, C.c_void_p, C.c_int64, C.c_void_p],
    "qil_coefficient_grid": [c_ctx, c_mps, C.c_void_p, C.c_void_p, C.c_void_p],
    "qil_coefficient_grid_dev": [c_ctx, c_mps, C.c_void_p, C.c_void_p, C.c_void_p],
    "qil_uploader_create": [c_ctx, C.c_int64, C.c_int, C.POINTER(C.c_void_p)],
    "qil_uploader_submit": [C.c_void_p, C.c_void_p, C.c_int64],
    "qil_uploader_acquire": [C.c_void_p, C.POINTER(C.c_void_p)],
    "qil_uploader_release": [C.c_void_p],
    "qil_uploader_destroy": [C.c_void_p],
    "qil_host_register": [C.c_void_p, C.c_int64],
    "qil_host_unregister": [C.c_void_p],
    "qil_mps_save": [c_mps, C.c_char_p],
    "qil_mps_load": [c_ctx, C.c_char_p, C.POINTER(c_mps)],
    "qil_mpo_save": [c_mpo, C.c_char_p],
    "qil_mpo_load": [c_ctx, C.c_char_p, C.POINTER(c_mpo)],
    "qil_argmax_abs_dev": [c_ctx, C.c_int, C.c_void_p, C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_double), C.c_void_p],
    "qil_coefficient_grid_argmax": [c_ctx, c_mps, C.c_void_p, C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_double),
                                    C.c_void_p],
    "qil_coefficient_batch_argmax": [c_ctx, c_mps, C.c_void_p, C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_double),
                                     C.c_void_p],
    "qil_mps_sum_sites": [c_ctx, c_mps, C.c_void_p, C.POINTER(c_mps)],
    "qil_mps_alloc": [c_ctx, C.c_int, C.c_int, C.c_void_p, C.c_double, C.POINTER(c_mps)],
    "qil_mps_core_ptr": [c_mps, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)],
    "qil_apply_mpo_mps": [c_ctx, c_mpo, c_mps, C.POINTER(c_mps)],
    "qil_apply_mpo_mps_zipup": [c_ctx, c_mpo, c_mps, C.c_double, C.c_int64, C.POINTER(c_mps)],
    "qil_apply_mpo_mps_batch": [c_ctx, c_mpo, C.c_void_p, C.c_int64, C.c_void_p],
    "qil_apply_mpo_mpo": [c_ctx, c_mpo, c_mpo, C.c_int, C.c_int, C.POINTER(c_mpo)],
    "qil_encode_svd": [c_ctx, C.c_int, C.c_void_p, C.c_int64, C.c_double, C.c_int64, C.POINTER(c_mps)],
    "qil_encode_svd_dev": [c_ctx, C.c_int, C.c_void_p, C.c_int64, C.c_double, C.c_int64, C.POINTER(c_mps)],
    "qil_encode_rsvd": [c_ctx, C.c_int, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_double,
                        C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.POINTER(c_mps)],
    "qil_encode_rsvd_dev": [c_ctx, C.c_int, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_double,
                            C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.POINTER(c_mps)],
    "qil_encode_rsvd_batch_dev": [c_ctx, C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int64,
                                  C.c_double, C.c_int64, C.c_int64, C.c_int, C.c_void_p, C.c_int64, C.c_int64,
                                  C.POINTER(c_mps)],
    "qil_get_stream": [c_ctx, C.POINTER(C.c_void_p)],
    "qil_encode_rsvd_sharded_dev": [c_ctx, C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int,
                                    C.c_int64, C.c_double, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_int64,
                                    C.POINTER(c_mps)],
    "qil_peer_create": [c_ctx, C.c_int, C.c_int, C.c_int64, C.POINTER(C.c_void_p), C.c_void_p],
    "qil_peer_connect": [C.c_void_p, C.c_void_p],
    "qil_peer_comm": [C.c_void_p, C.c_void_p],
    "qil_peer_destroy": [C.c_void_p],
    "qil_ztmps_split": [c_ctx, c_mps, C.c_double, C.c_int64, C.POINTER(c_mps)],
    "qil_canonicalize": [c_ctx, c_mps, C.c_int, C.c_int, C.c_double, C.c_int64],
    "qil_compress": [c_ctx, c_mps, C.c_int64, C.c_double, C.c_int],
    "qil_norm": [c_ctx, c_mps, C.POINTER(C.c_double)],
    "qil_build_qft_mpo": [c_ctx, C.c_int, C.c_double, C.c_int64, C.POINTER(c_mpo)],
    "qil_build_dt_mpo": [c_ctx, C.c_int, C.c_double, C.c_double, C.c_int64, C.POINTER(c_mpo)],
    "qil_build_zt_mpo": [c_ctx, C.c_int, C.c_double, C.c_double, C.c_int64, C.POINTER(c_mpo)],
    "qil_qr": [c_ctx, C.c_int, C.c_int64, C.c_int64, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p],
    "qil_rsvd": [c_ctx, C.c_int, C.c_int64, C.c_int64, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_double,
                 C.c_int64, C.c_int64, C.c_void_p, C.c_int64, i64p, C.c_void_p, C.c_void_p, C.c_void_p],
    "qil_svd_trunc": [c_ctx, C.c_int, C.c_int64, C.c_int64, C.c_void_p, C.c_double, C.c_int64, C.c_int64,
                      i64p, C.c_void_p, C.c_void_p, C.c_void_p],
}


class QilComm(C.Structure):
    """struct qil_comm (include/qilcuda.h): the two collectives of a row-sharded encode."""
    ALLREDUCE = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int64)
    ALLGATHER = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64)
    _fields_ = [("rank", C.c_int), ("world", C.c_int), ("user", C.c_void_p),
                ("allreduce_sum_f64", ALLREDUCE), ("allgather_f64", ALLGATHER)]


def declared_symbols():
    """Every entry point include/qilcuda.h declares (used by the CPU-side ABI test)."""
    return sorted(list(_SIGS) + ["qil_last_error", "qil_version"])


def load():
    """Load libqilcuda.so (building is the job of __graft_entry__.build / build.py)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python qilaplace.jl_b200/build.py` "
            "(there is no CPU fallback for the CUDA path)")
    lib = C.CDLL(LIB_PATH)
    for name, args in _SIGS.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = C.c_int
    lib.qil_last_error.restype = C.c_char_p
    lib.qil_last_error.argtypes = []
    lib.qil_version.restype = C.c_char_p
    lib.qil_version.argtypes = []
    _lib = lib
    return lib


def check(status):
    if status != 0:
        msg = load().qil_last_error().decode("utf-8", "replace")
        raise _ERRMAP.get(status, RuntimeError)(msg)


def call(name, *args):
    check(getattr(load(), name)(*args))
