"""GP predictor of a keyframe: drop-in for `Mapping.prep_predictor` (como/odom/Mapping.py:430-468) built on
the fused K-matrix kernels (csrc/depthcov.cu): returns (K_mm_inv, L_mm, Knm_Kmminv) with the same shapes.
The 64 x 64 factorisation stays a library call (torch / cuSOLVER), as in the reference."""
import torch

from como_b200 import _lib


def prep_predictor(cov_params_img, coords_m, scale, photo_img_size=None, jitter=1e-6):
    """cov_params_img (B,4,H,W) float64, coords_m (B,M,2) [row,col] float64, scale = model.cov_modules[-1].get_scale()."""
    dev = _lib.require_cuda(cov_params_img, coords_m)
    B, _, H, W = cov_params_img.shape
    if photo_img_size is not None and tuple(photo_img_size) != (H, W):
        raise NotImplementedError("como_b200 prep_predictor expects the covariance image at the photometric resolution")
    M = coords_m.shape[1]
    cov = cov_params_img.to(torch.float64).contiguous()
    cm = coords_m.to(torch.float64).contiguous()
    E_m = torch.empty(B, M, 4, dtype=torch.float64, device=dev)
    K_mm = torch.empty(B, M, M, dtype=torch.float64, device=dev)
    sc = float(scale)
    with torch.cuda.device(dev):
        st = _lib.kmat_kmm(_lib.ptr(cov), B, H, W, _lib.ptr(cm), M, sc, float(jitter), _lib.ptr(E_m), _lib.ptr(K_mm),
                           _lib.stream_ptr(dev))
        _lib.check(st, "como_b200_kmat_kmm")
        L_mm, _ = torch.linalg.cholesky_ex(K_mm, upper=False)
        eye = torch.eye(M, device=dev, dtype=torch.float64).unsqueeze(0).repeat(B, 1, 1)
        K_mm_inv = torch.cholesky_solve(eye, L_mm, upper=False).contiguous()
        out = torch.empty(B, H, W, M, dtype=torch.float64, device=dev)
        st = _lib.kmat_predictor(_lib.ptr(cov), B, H, W, _lib.ptr(cm), _lib.ptr(E_m), _lib.ptr(K_mm_inv), M, sc,
                                 _lib.ptr(out), _lib.stream_ptr(dev))
        _lib.check(st, "como_b200_kmat_predictor")
    return K_mm_inv, L_mm, out
