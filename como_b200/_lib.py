"""ctypes binding of libcomo_b200.so (the C ABI declared in include/como_b200.h).

There is no CPU fallback: if the shared library is missing the import fails loudly, and every
wrapper raises RuntimeError on a non-zero status (mirroring the reference's TORCH_CHECK -> RuntimeError
convention, como/backend/src/cov.cpp:25-62).
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("COMO_B200_LIB") or os.path.join(_HERE, "libcomo_b200.so")   # override: tuning builds only
MAX_LEVELS = 8
TRACK_STAT_STRIDE = 32

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: build it with `python build.py` (nvcc, sm_100a). "
        "como_b200 has no CPU or PyTorch fallback path."
    )
lib = C.CDLL(LIB_PATH)


class TrackLevel(C.Structure):
    _fields_ = [
        ("vals", C.c_void_p), ("P", C.c_void_p), ("J", C.c_void_p), ("mask", C.c_void_p), ("img", C.c_void_p),
        ("pack", C.c_void_p), ("n", C.c_int32), ("w", C.c_int32), ("h", C.c_int32), ("c", C.c_int32), ("K", C.c_float * 9),
    ]


class TrackTerm(C.Structure):
    _fields_ = [("max_iter", C.c_int32), ("delta_norm", C.c_float), ("rel_tol", C.c_float), ("grad_norm", C.c_float)]


class PackItem(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst_offset_bytes", C.c_int64), ("count", C.c_int64), ("src_dtype", C.c_int32),
                ("dst_dtype", C.c_int32)]


PACK_MAX_ITEMS = 8
PACK_DTYPES = {torch.float32: 0, torch.float64: 1, torch.uint8: 2, torch.bool: 2, torch.int32: 3, torch.int64: 4}


def _sig(name, restype, argtypes):
    f = getattr(lib, name)
    f.restype = restype
    f.argtypes = argtypes
    return f


abi_version = _sig("como_b200_abi_version", C.c_int, [])
last_error = _sig("como_b200_last_error", C.c_char_p, [])
se3_exp = _sig("como_b200_se3_exp", None, [C.POINTER(C.c_double), C.POINTER(C.c_double)])
track_workspace_bytes = _sig("como_b200_track_workspace_bytes", C.c_size_t, [C.c_int32, C.c_int32])
track_pack_bytes = _sig("como_b200_track_pack_bytes", C.c_size_t, [C.c_int32, C.c_int32])
track_pack = _sig("como_b200_track_pack", C.c_int, [C.POINTER(TrackLevel), C.c_void_p])
track_pyr = _sig(
    "como_b200_track_pyr", C.c_int,
    [C.POINTER(TrackLevel), C.c_int32, C.c_int32, C.POINTER(TrackTerm), C.c_void_p, C.c_void_p, C.c_void_p,
     C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p])
track_debug_candidate_cap = _sig("como_b200_track_debug_candidate_cap", None, [C.c_int32])
precalc_jacobians = _sig(
    "como_b200_precalc_jacobians", C.c_int,
    [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_float), C.c_int64, C.c_int32, C.c_void_p, C.c_void_p])

VP, I32, I64, F64 = C.c_void_p, C.c_int32, C.c_int64, C.c_double
median_workspace_bytes = _sig("como_b200_median_workspace_bytes", C.c_size_t, [I32, I32])
median_f64 = _sig("como_b200_median_f64", C.c_int, [VP, VP, I32, I64, F64, VP, VP, VP, C.c_size_t, VP])
median_f32 = _sig("como_b200_median_f32", C.c_int, [VP, VP, I32, I64, C.c_float, VP, VP, VP, C.c_size_t, VP])
subselect_pixels = _sig("como_b200_subselect_pixels", C.c_int, [VP, I32, I32, I32, I32, VP, VP, VP])
ba_scaffold = _sig("como_b200_ba_scaffold", C.c_int,
                   [VP, VP, VP, VP, VP, VP, I32, I32, I32, C.POINTER(F64), VP, VP, VP])
predictor_apply = _sig("como_b200_predictor_apply", C.c_int, [VP, VP, I32, I64, I32, VP, VP])
predictor_stream_ctas = _sig("como_b200_predictor_stream_ctas", None, [I32])
predictor_colsum = _sig("como_b200_predictor_colsum", C.c_int, [VP, I64, I32, VP, VP])
ba_frames_bytes = _sig("como_b200_ba_frames_bytes", C.c_size_t, [I32])
ba_partial_doubles = _sig("como_b200_ba_partial_doubles", C.c_size_t, [I32])
ba_unit_ints = _sig("como_b200_ba_unit_ints", I32, [])
ba_target_group = _sig("como_b200_ba_target_group", I32, [])
ba_photo_residual = _sig("como_b200_ba_photo_residual", C.c_int,
                         [VP] * 13 + [I32] * 7 + [C.POINTER(F64), VP, VP, VP, VP, VP])
ba_photo_accum = _sig("como_b200_ba_photo_accum", C.c_int,
                      [VP] * 14 + [I32] * 9 + [C.POINTER(F64), I32] + [VP] * 11)
median_num_passes = _sig("como_b200_median_num_passes", I32, [I32])
median_pass_f64 = _sig("como_b200_median_pass_f64", C.c_int, [VP, VP, I32, I64, I32, VP, VP])
median_finish_f64 = _sig("como_b200_median_finish_f64", C.c_int, [I32, VP, F64, VP, VP, VP])
median_pack_words = _sig("como_b200_median_pack_words", I32, [])
median_dist_compact_f64 = _sig("como_b200_median_dist_compact_f64", C.c_int, [VP, VP, I32, I64, VP, VP, VP])
median_dist_finish_f64 = _sig("como_b200_median_dist_finish_f64", C.c_int, [VP, I32, I32, VP, F64, VP, VP, VP])
ba_priors = _sig("como_b200_ba_priors", C.c_int,
                 [VP] * 15 + [I32, I32, C.POINTER(F64), F64, I32, I32, I32, I32, C.POINTER(F64), I32, VP, VP, VP, VP])
ba_update = _sig("como_b200_ba_update", C.c_int, [VP, I32, I32, I32, VP, VP, VP, VP, VP, VP])

F32 = C.c_float
cross_covariance = _sig("como_b200_cross_covariance", C.c_int, [VP, VP, VP, VP, F64, I32, I32, I32, I32, VP, VP])
chol_append = _sig("como_b200_chol_append", C.c_int, [VP, VP, VP, VP, VP, F32, I32, I32, I32, I32, VP])
sampler_workspace_bytes = _sig("como_b200_sampler_workspace_bytes", C.c_size_t, [I32, I32])
sampler_greedy = _sig("como_b200_sampler_greedy", C.c_int,
                      [VP, VP, I32, I32, I32, I32, VP, VP, VP, VP, VP, VP, VP, F32, F32, I32, F32, F32, I32, VP, VP,
                       C.c_size_t, VP])
kmat_kmm = _sig("como_b200_kmat_kmm", C.c_int, [VP, I32, I32, I32, VP, I32, F64, F64, VP, VP, VP])
kmat_predictor = _sig("como_b200_kmat_predictor", C.c_int, [VP, I32, I32, I32, VP, VP, VP, I32, F64, VP, VP])
chol_solve_workspace_bytes = _sig("como_b200_chol_solve_workspace_bytes", C.c_size_t, [I32])
chol_ctas = _sig("como_b200_chol_ctas", None, [I32])
chol_schedule = _sig("como_b200_chol_schedule", None, [C.c_int32])
chol_solve = _sig("como_b200_chol_solve", C.c_int, [VP, VP, I32, VP, VP, C.c_size_t, VP])
kmat_rows = _sig("como_b200_kmat_rows", C.c_int, [VP, I32, I32, I32, VP, VP, VP, I32, F64, VP, VP, I64, VP, VP, VP, VP])
weighted_gram = _sig("como_b200_weighted_gram", C.c_int, [VP, VP, VP, VP, I64, I32, F64, F64, VP, VP, VP, VP])
rows_residual = _sig("como_b200_rows_residual", C.c_int, [VP, VP, VP, VP, I64, I32, VP, VP, VP])
sfm_linearize = _sig("como_b200_sfm_linearize", C.c_int,
                     [VP, VP, VP, VP, VP, I32, I32, I64, I32, C.POINTER(F64), C.POINTER(F64), VP, VP, VP, VP, VP])
sfm_accumulate = _sig("como_b200_sfm_accumulate", C.c_int, [VP, VP, VP, I64, I32, VP, VP, VP, VP])
reproject_dense = _sig("como_b200_reproject_dense", C.c_int,
                       [VP, I32, I32, C.POINTER(F64), C.POINTER(F64), F64, VP, VP, VP, VP, VP])
sample_depth_gradmag = _sig("como_b200_sample_depth_gradmag", C.c_int, [VP, I32, I32, VP, VP, I32, VP, VP, VP])

handoff_pack = _sig("como_b200_handoff_pack", C.c_int, [C.POINTER(PackItem), C.c_int32, C.c_void_p, C.c_void_p])
gray_pyramid = _sig("como_b200_gray_pyramid", C.c_int, [VP, I32, I32, I32, C.POINTER(C.c_void_p), VP])
image_pyramid_fused = _sig("como_b200_image_pyramid_fused", C.c_int,
                           [VP, I32, I32, I32, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), VP])
image_gradients = _sig("como_b200_image_gradients", C.c_int, [VP, I32, I32, VP, VP, VP])
img_and_grads_f64 = _sig("como_b200_img_and_grads_f64", C.c_int, [VP, I32, I32, VP, VP])
kf_reference_level = _sig("como_b200_kf_reference_level", C.c_int,
                          [VP, VP, VP, VP, I32, I32, I32, I32, I32, C.POINTER(F32), VP, F32, F32, VP, VP, VP, VP, VP, VP])
reproj_depth = _sig("como_b200_reproj_depth", C.c_int, [VP, I32, VP, C.POINTER(F32), I32, I32, VP, VP, VP])

# every symbol include/como_b200.h declares (checked by tests/test_abi.py without a GPU)
DECLARED_SYMBOLS = [
    "como_b200_abi_version", "como_b200_last_error", "como_b200_se3_exp", "como_b200_track_workspace_bytes", "como_b200_track_pack_bytes", "como_b200_track_pack", "como_b200_track_pyr",
    "como_b200_track_debug_candidate_cap",
    "como_b200_precalc_jacobians", "como_b200_median_workspace_bytes", "como_b200_median_f64", "como_b200_median_f32",
    "como_b200_subselect_pixels", "como_b200_ba_scaffold", "como_b200_predictor_apply", "como_b200_predictor_stream_ctas", "como_b200_predictor_colsum",
    "como_b200_ba_frames_bytes", "como_b200_ba_partial_doubles", "como_b200_ba_unit_ints", "como_b200_ba_target_group",
    "como_b200_ba_photo_residual", "como_b200_ba_photo_accum", "como_b200_median_num_passes",
    "como_b200_median_pass_f64", "como_b200_median_finish_f64", "como_b200_median_pack_words",
    "como_b200_median_dist_compact_f64", "como_b200_median_dist_finish_f64",
    "como_b200_ba_priors", "como_b200_ba_update", "como_b200_cross_covariance", "como_b200_chol_append",
    "como_b200_sampler_workspace_bytes", "como_b200_sampler_greedy", "como_b200_kmat_kmm", "como_b200_kmat_predictor",
    "como_b200_chol_solve_workspace_bytes", "como_b200_chol_solve", "como_b200_chol_ctas", "como_b200_chol_schedule",
    "como_b200_kmat_rows", "como_b200_weighted_gram", "como_b200_rows_residual",
    "como_b200_reproject_dense", "como_b200_sample_depth_gradmag", "como_b200_sfm_linearize", "como_b200_sfm_accumulate",
    "como_b200_handoff_pack", "como_b200_gray_pyramid", "como_b200_image_pyramid_fused", "como_b200_image_gradients", "como_b200_img_and_grads_f64", "como_b200_kf_reference_level", "como_b200_reproj_depth",
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
