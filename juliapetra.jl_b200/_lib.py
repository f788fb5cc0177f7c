"""ctypes binding of libjpetra_b200.so (include/jpetra_b200.h) -- the same entry points a Julia host
binds with `ccall`.  Loading fails loudly when the library is missing; every compute
call fails loudly (BackendError) when no sm_100 device is present.  There is no CPU fallback."""
from __future__ import annotations

import ctypes as C
import glob
import os
import sys

import numpy as np

from .errors import BackendError, InvalidArgumentError, InvalidStateError

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libjpetra_b200.so")

OK, INVALID_ARGUMENT, INVALID_STATE, CUDA_ERROR, NCCL_ERROR = 0, 1, 2, 3, 4
NO_TRANS, TRANS, CONJ_TRANS = 1, 2, 3
ADD, INSERT, REPLACE, ABSMAX, ZERO = 1, 2, 3, 4, 5
SUM, MAX, MIN = 0, 1, 2
F64, I64 = 0, 1

c_i32, c_i64, c_f64, c_vp = C.c_int32, C.c_int64, C.c_double, C.c_void_p
P = C.POINTER

# name -> argtypes; every function returns int32 status (jp_version returns the version)
PROTOTYPES = {
    "jp_init": [c_i32],
    "jp_finalize": [],
    "jp_last_error": [C.c_char_p, c_i64],
    "jp_device_info": [P(c_i32), P(c_i64), P(c_i32), P(c_i32)],
    "jp_synchronize": [],
    "jp_version": [],
    "jp_host_alloc": [c_i64, P(c_vp)],
    "jp_host_free": [c_vp],
    "jp_event_create": [P(c_i64)],
    "jp_event_record": [c_i64],
    "jp_event_elapsed_ms": [c_i64, c_i64, P(c_f64)],
    "jp_event_destroy": [c_i64],
    "jp_launch_count": [P(c_i64)],
    "jp_flush_l2": [c_i64],
    "jp_comm_unique_id": [c_vp],
    "jp_comm_create_serial": [P(c_i64)],
    "jp_comm_create_nccl": [c_vp, c_i32, c_i32, P(c_i64)],
    "jp_comm_destroy": [c_i64],
    "jp_comm_rank": [c_i64, P(c_i32), P(c_i32)],
    "jp_comm_barrier": [c_i64],
    "jp_comm_allreduce": [c_i64, c_i32, c_i32, c_vp, c_vp, c_i64],
    "jp_comm_scan_sum": [c_i64, c_i32, c_vp, c_vp, c_i64],
    "jp_comm_gather_all": [c_i64, c_i32, c_vp, c_i64, c_vp, c_vp],
    "jp_comm_broadcast": [c_i64, c_i32, c_vp, c_i64, c_i32],
    "jp_comm_alltoallv": [c_i64, c_i32, c_vp, c_vp, c_vp, c_vp],
    "jp_csr_create": [c_i32, c_i32, c_i32, c_i64, c_vp, c_vp, c_vp, c_i32, P(c_i64)],
    "jp_csr_destroy": [c_i64],
    "jp_csr_info": [c_i64, c_vp],
    "jp_csr_download": [c_i64, c_vp, c_vp, c_vp],
    "jp_csr_spmm_info": [c_i64, c_vp],
    "jp_csr_fill_complete": [c_i32, c_vp, c_vp, c_vp, c_i64, c_i64, c_i64, c_i32, P(c_i64), c_vp],
    "jp_csr_col_map": [c_i64, c_vp],
    "jp_mv_create": [c_i32, c_i32, P(c_i64)],
    "jp_mv_destroy": [c_i64],
    "jp_mv_shape": [c_i64, P(c_i32), P(c_i32)],
    "jp_mv_upload": [c_i64, c_vp],
    "jp_mv_download": [c_i64, c_vp],
    "jp_mv_copy": [c_i64, c_i64],
    "jp_mv_fill": [c_i64, c_f64],
    "jp_mv_scale": [c_i64, c_f64],
    "jp_mv_scale_vec": [c_i64, c_vp],
    "jp_mv_axpby": [c_i64, c_f64, c_i64, c_f64],
    "jp_mv_axpby_norm2": [c_i64, c_i64, c_f64, c_i64, c_f64, c_vp],
    "jp_mv_dot": [c_i64, c_i64, c_i64, c_vp],
    "jp_mv_norm2": [c_i64, c_i64, c_vp],
    "jp_mv_normp": [c_i64, c_i64, c_f64, c_vp],
    "jp_mv_comm_reduce": [c_i64, c_i64],
    "jp_scalars_create": [c_i32, P(c_i64)],
    "jp_scalars_destroy": [c_i64],
    "jp_scalars_upload": [c_i64, c_vp],
    "jp_scalars_download": [c_i64, c_vp],
    "jp_scalars_copy": [c_i64, c_i32, c_i32],
    "jp_mv_dot_dev": [c_i64, c_i64, c_i64, c_i64, c_i32],
    "jp_mv_axpby_dev": [c_i64, c_i64, c_i64, c_f64, c_i32, c_i32, c_f64, c_i32, c_i32],
    "jp_mv_axpby_norm2sq_dev": [c_i64, c_i64, c_i64, c_i64, c_f64, c_i32, c_i32, c_f64, c_i32, c_i32, c_i32],
    "jp_apply_dot_dev": [c_i64, c_i64, c_i64, c_i64, c_i64, c_f64, c_f64, c_i64, c_i64, c_i32],
    "jp_plan_create": [c_i64, c_i32, c_i32, c_i32, c_vp, c_vp, c_i32, c_vp, c_i32, c_vp, c_i32, c_vp, c_vp, c_i32, c_vp, c_vp,
                       c_i32, c_i32, P(c_i64)],
    "jp_plan_destroy": [c_i64],
    "jp_transfer": [c_i64, c_i64, c_i64, c_i32, c_i32],
    "jp_transfer_post": [c_i64, c_i64, c_i64, c_i32, c_i32],
    "jp_transfer_wait": [c_i64],
    "jp_local_apply": [c_i64, c_i64, c_i64, c_i32, c_f64, c_f64],
    "jp_apply": [c_i64, c_i64, c_i64, c_i64, c_i64, c_i32, c_f64, c_f64],
    "jp_apply_export": [c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, c_i32, c_f64, c_f64],
    "jp_apply_dot": [c_i64, c_i64, c_i64, c_i64, c_i64, c_i32, c_f64, c_f64, c_i64, c_vp],
    "jp_apply_host": [c_i64, c_i64, c_i64, c_vp, c_vp, c_i32, c_i32, c_f64, c_f64],
}

_lib = None
_initialised_device = None


def _preload_nccl():
    """Make one NCCL serve both torch and this library: prefer the copy torch bundles (loaded by SONAME
    libnccl.so.2), fall back to the system one."""
    for sp in sys.path:
        cand = glob.glob(os.path.join(sp, "nvidia", "nccl", "lib", "libnccl.so.2"))
        if cand:
            try:
                C.CDLL(cand[0], mode=C.RTLD_GLOBAL)
                return cand[0]
            except OSError:
                pass
    return None


def load():
    """Load libjpetra_b200.so (no device needed for this step)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        # not built yet (e.g. a fresh checkout: the .so is git-ignored): compile it; this needs nvcc, not a GPU
        import shutil
        import subprocess
        if shutil.which("nvcc") and shutil.which("make"):
            subprocess.run(["make", "-C", os.path.join(_HERE, "csrc"), "-j8"], check=False, capture_output=True)
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C juliapetra.jl_b200/csrc`). This backend has no CPU fallback.")
    _preload_nccl()
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError here = the .so does not export what the header declares
        fn.argtypes = argtypes
        fn.restype = c_i32
    _lib = lib
    return lib


def last_error() -> str:
    buf = C.create_string_buffer(2048)
    load().jp_last_error(buf, 2048)
    return buf.value.decode(errors="replace")


def check(rc: int):
    if rc == OK:
        return
    msg = last_error()
    if rc == INVALID_ARGUMENT:
        raise InvalidArgumentError(msg)
    if rc == INVALID_STATE:
        raise InvalidStateError(msg)
    raise BackendError(msg)


def init(device: int | None = None):
    """jp_init on LOCAL_RANK (torchrun) or device 0; idempotent."""
    global _initialised_device
    lib = load()
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    if _initialised_device is None:
        check(lib.jp_init(device))
        _initialised_device = device
    return lib


def lib():
    """The loaded library with the runtime initialised (fails loudly without a GPU)."""
    return init()


def ptr(a: np.ndarray | None):
    if a is None:
        return None
    return a.ctypes.data_as(c_vp)


def as_i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def as_i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def as_f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)
