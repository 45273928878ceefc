"""CPU-only: the C-ABI library loads and exports every symbol include/como_b200.h declares."""
import ctypes
import os
import re

from como_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "como_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(como_b200_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    syms = header_symbols()
    assert syms, "no symbols parsed from the header"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/como_b200.h but not exported"
    assert sorted(_lib.DECLARED_SYMBOLS) == syms


def test_abi_version_and_error_string():
    assert _lib.abi_version() == 3   # 3: packed keyframe tiles for the tracker (como_b200_track_pack)
    assert isinstance(_lib.last_error(), bytes)


def test_argument_errors_do_not_need_a_gpu():
    # null pointers are rejected before any CUDA call
    term = _lib.TrackTerm(1, 1e-3, 1e-3, 1.0)
    st = _lib.track_pyr(None, 1, 1, ctypes.byref(term), None, None, None, None, None, 0, None)
    assert st == -1
    assert b"null" in _lib.last_error()
    assert _lib.track_workspace_bytes(1000, 1) > 1000 * 4


def test_se3_exp_of_the_library_is_the_matrix_exponential():
    """The kernels' own se3_exp_tau_phi (common.cuh), evaluated on the host through the C ABI, against scipy's expm of
    the twist matrix: pins the restatement of the un-vendored lietorch exponential without a GPU."""
    import ctypes as C

    import numpy as np
    from scipy.linalg import expm

    from como_b200 import _lib

    rng = np.random.default_rng(0)
    for s in (1e-9, 1e-7, 1e-3, 0.3, 1.0, 3.0):
        x = rng.standard_normal(6) * s
        A = np.zeros((4, 4))
        A[:3, :3] = [[0, -x[5], x[4]], [x[5], 0, -x[3]], [-x[4], x[3], 0]]
        A[:3, 3] = x[:3]
        xin = (C.c_double * 6)(*x.tolist())
        out = (C.c_double * 16)()
        _lib.se3_exp(xin, out)
        np.testing.assert_allclose(np.array(out[:]).reshape(4, 4), expm(A), rtol=0, atol=1e-13)


def test_tracker_pack_and_workspace_sizes_without_a_gpu():
    """Host-side size functions of the tracker's packed keyframe tiles: 512-pixel tiles of 22 528 bytes
    ([P 12 | I_ref 4 | J 24 | residual 4] B per pixel), one set of tiles per image channel."""
    from como_b200 import _lib

    tile = 512 * 44
    assert _lib.track_pack_bytes(0, 1) == 0
    assert _lib.track_pack_bytes(1, 1) == tile and _lib.track_pack_bytes(512, 1) == tile
    assert _lib.track_pack_bytes(513, 1) == 2 * tile
    assert _lib.track_pack_bytes(513, 3) == 3 * 2 * tile
    assert _lib.track_pack_bytes(307200, 0) == 600 * tile           # c = 0 reads as one channel
    # the workspace no longer depends on the number of points (the residuals live in the packed tiles)
    assert _lib.track_workspace_bytes(1000, 4) == _lib.track_workspace_bytes(307200, 4)
    assert _lib.track_workspace_bytes(1000, 8) > _lib.track_workspace_bytes(1000, 4)
