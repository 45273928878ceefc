"""Drop-in for the reference's pybind module `como_backends` (como/backend/src/depth_cov_backends.cpp:3-6):
same two function names, argument order, return/mutation behaviour and RuntimeError messages, backed by
the sm_100a kernels in csrc/depthcov.cu.  CUDA tensors only (no CPU build of this backend).
"""
import torch

from como_b200 import _lib


def _same_device(*ts):
    devs = {t.device for t in ts}
    if len(devs) != 1 or not next(iter(devs)).type == "cuda":
        raise RuntimeError("All variables must be on same device.")
    return ts[0].device


def cross_covariance(x1, E1, x2, E2, scale):
    """x1 (B,N,2), E1 (B,N,2,2), x2 (B,M,2), E2 (B,M,2,2), float scale -> new (B,N,M) tensor.
    Inputs may be non-contiguous views (the reference does not check, cov.cpp:6-9)."""
    dev = _same_device(x1, E1, x2, E2)
    if x1.dtype not in (torch.float32, torch.float64):
        raise RuntimeError("cross_covariance: only float32/float64 are supported by como_b200")
    dt = x1.dtype
    B, n1, n2 = x1.shape[0], x1.shape[1], x2.shape[1]
    a, A, c, Cm = (t.to(dt).contiguous() for t in (x1, E1, x2, E2))
    out = torch.empty((B, n1, n2), dtype=dt, device=dev)
    with torch.cuda.device(dev):
        st = _lib.cross_covariance(_lib.ptr(a), _lib.ptr(A), _lib.ptr(c), _lib.ptr(Cm), float(scale), B, n1, n2,
                                   4 if dt == torch.float32 else 8, _lib.ptr(out), _lib.stream_ptr(dev))
    _lib.check(st, "cross_covariance")
    return out


def get_new_chol_obs_info(L, obs_info, var, k_ni, k_id, k_ii, N):
    """Mutates L (B,n,n), obs_info (B,n,d), var (B,d) in place (float32, contiguous)."""
    for name, t in (("L", L), ("obs_info", obs_info), ("var", var), ("k_ni", k_ni), ("k_id", k_id)):
        if not t.is_contiguous():
            raise RuntimeError(f"{name} must be contiguous")
    dev = _same_device(L, obs_info, k_ni, k_id)
    if L.dtype != torch.float32:
        raise RuntimeError("get_new_chol_obs_info: float32 only (as the reference kernels)")
    B, n, _ = L.shape
    d = obs_info.shape[2]
    with torch.cuda.device(dev):
        st = _lib.chol_append(_lib.ptr(L), _lib.ptr(obs_info), _lib.ptr(var), _lib.ptr(k_ni), _lib.ptr(k_id), float(k_ii),
                              B, n, d, int(N), _lib.stream_ptr(dev))
    _lib.check(st, "get_new_chol_obs_info")
