"""ctypes binding of libcomo_b200.so (the C ABI declared in include/como_b200.h).

There is no CPU fallback: if the shared library is missing the import fails loudly, and every
wrapper raises RuntimeError on a non-zero status (mirroring the reference's TORCH_CHECK -> RuntimeError
convention, como/backend/src/cov.cpp:25-62).
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcomo_b200.so")
MAX_LEVELS = 8
TRACK_STAT_STRIDE = 8

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: build it with `python build.py` (nvcc, sm_100a). "
        "como_b200 has no CPU or PyTorch fallback path."
    )
lib = C.CDLL(LIB_PATH)


class TrackLevel(C.Structure):
    _fields_ = [
        ("vals", C.c_void_p), ("P", C.c_void_p), ("J", C.c_void_p), ("mask", C.c_void_p), ("img", C.c_void_p),
        ("n", C.c_int32), ("w", C.c_int32), ("h", C.c_int32), ("K", C.c_float * 9),
    ]


class TrackTerm(C.Structure):
    _fields_ = [("max_iter", C.c_int32), ("delta_norm", C.c_float), ("rel_tol", C.c_float), ("grad_norm", C.c_float)]


def _sig(name, restype, argtypes):
    f = getattr(lib, name)
    f.restype = restype
    f.argtypes = argtypes
    return f


abi_version = _sig("como_b200_abi_version", C.c_int, [])
last_error = _sig("como_b200_last_error", C.c_char_p, [])
track_workspace_bytes = _sig("como_b200_track_workspace_bytes", C.c_size_t, [C.c_int32, C.c_int32])
track_pyr = _sig(
    "como_b200_track_pyr", C.c_int,
    [C.POINTER(TrackLevel), C.c_int32, C.c_int32, C.POINTER(TrackTerm), C.c_void_p, C.c_void_p, C.c_void_p,
     C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p])
precalc_jacobians = _sig(
    "como_b200_precalc_jacobians", C.c_int,
    [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_float), C.c_int64, C.c_void_p, C.c_void_p])

# every symbol include/como_b200.h declares (checked by tests/test_abi.py without a GPU)
DECLARED_SYMBOLS = [
    "como_b200_abi_version", "como_b200_last_error", "como_b200_track_workspace_bytes", "como_b200_track_pyr",
    "como_b200_precalc_jacobians",
]


def check(status, what):
    if status != 0:
        raise RuntimeError(f"{what} failed (status {status}): {last_error().decode()}")


def stream_ptr(device=None):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def require_cuda(*tensors):
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("All variables must be on same device.  (como_b200 is CUDA-only: no CPU fallback)")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError("All variables must be on same device.")
    return dev
