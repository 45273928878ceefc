"""Depth distillation for keyframe creation: drop-ins for `distill_depth_from_scratch` and
`distill_conditional_depth_from_scratch` (como/depth_cov/core/distill_depth.py:85-175).

The reference materialises K_nm (n x m), the predictor K_nm K_mm^-1, the stacked (m+n) x m least-squares matrix
and its Gram through stock torch ops.  Here `kmat_rows` (csrc/kmat.cu) writes the predictor rows and the
predictive variance in one pass, `weighted_gram` forms the normal equations  A^T A = L^-T L^-1 + sum_n s_n^2 k_n k_n^T,
A^T b = sum_n s_n^2 log z_n k_n  on the FP64 tensor path, and only the m x m factorisations stay library calls
(torch / cuSOLVER), exactly as `lstsq_chol` (como/utils/lin_alg.py:82-87) does.

Test points are passed UNCOMPACTED with a validity mask (uint8): the reference's boolean-mask compaction only
changes the summation order.  There is no CPU path.
"""
import math

import torch

from como_b200 import _lib

F64 = torch.float64


def _scale_of(model):
    if hasattr(model, "get_scale"):
        return float(model.get_scale(-1))
    return float(model)


def predictor_rows(coords_m, coords_n, mask_n, cov_params_img, scale, want_var):
    """calc_kernel_matrices + get_predictor (distill_depth.py:8-48) at arbitrary test points.
    coords_m (1,m,2), coords_n (1,n,2) [row,col] float64, mask_n (n,) uint8 or None.
    Returns (rows (n,m), L_mm (1,m,m), var_n (n,) or None, var_min 0-d tensor or None)."""
    dev = _lib.require_cuda(cov_params_img, coords_m, coords_n)
    B, _, H, W = cov_params_img.shape
    if B != 1:
        raise RuntimeError("como_b200 distill_depth: batch size must be 1 (as the reference asserts)")
    cov = cov_params_img.to(F64).contiguous()
    cm = coords_m.to(F64).contiguous()
    cn = coords_n.to(F64).contiguous()
    m, n = cm.shape[1], cn.shape[1]
    E_m = torch.empty(1, m, 4, dtype=F64, device=dev)
    K_mm = torch.empty(1, m, m, dtype=F64, device=dev)
    rows = torch.empty(n, m, dtype=F64, device=dev)
    var_n = torch.empty(n, dtype=F64, device=dev) if want_var else None
    var_min = torch.empty(1, dtype=F64, device=dev) if want_var else None
    with torch.cuda.device(dev):
        stream = _lib.stream_ptr(dev)
        st = _lib.kmat_kmm(_lib.ptr(cov), 1, H, W, _lib.ptr(cm), m, float(scale), 0.0, _lib.ptr(E_m), _lib.ptr(K_mm), stream)
        _lib.check(st, "como_b200_kmat_kmm")
        L_mm, _ = torch.linalg.cholesky_ex(K_mm, upper=False)
        eye = torch.eye(m, device=dev, dtype=F64).unsqueeze(0)
        Kinv = torch.cholesky_solve(eye, L_mm, upper=False).contiguous()
        st = _lib.kmat_rows(_lib.ptr(cov), 1, H, W, _lib.ptr(cm), _lib.ptr(E_m), _lib.ptr(Kinv), m, float(scale), _lib.ptr(cn),
                            _lib.ptr(mask_n), n, _lib.ptr(rows), _lib.ptr(var_n), _lib.ptr(var_min), stream)
        _lib.check(st, "como_b200_kmat_rows")
    return rows, L_mm, var_n, var_min


def _gram(rows, y, var, mask, var_add, wscale):
    dev = rows.device
    n, m = rows.shape
    G = torch.empty(m, m, dtype=F64, device=dev)
    h = torch.empty(m, dtype=F64, device=dev)
    with torch.cuda.device(dev):
        st = _lib.weighted_gram(_lib.ptr(rows), _lib.ptr(y), _lib.ptr(var), _lib.ptr(mask), n, m, float(var_add),
                                float(wscale), _lib.ptr(G), _lib.ptr(h), None, _lib.stream_ptr(dev))
        _lib.check(st, "como_b200_weighted_gram")
    return G, h


def _chol_solve(ATA, ATb):
    L, _ = torch.linalg.cholesky_ex(ATA, upper=False)
    return torch.cholesky_solve(ATb, L, upper=False)


def distill_depth_masked(coords_m, coords_n, logz_obs, mask_n, cov_params_img, scale, with_prior, stdev_obs=None):
    """Core of distill_depth_from_scratch on uncompacted test points.  Returns (logz_m (1,m,1), residuals (n,),
    stats (3,) = [count, sum, centred sum of squares] of the residuals over valid points)."""
    dev = cov_params_img.device
    need_var = with_prior and stdev_obs is None
    rows, L_mm, var_n, var_min = predictor_rows(coords_m, coords_n, mask_n, cov_params_img, scale, want_var=need_var)
    m = rows.shape[1]
    if not with_prior:
        # plain least squares on the predictor rows (distill_depth.py:58-59): no observation weights
        G, h = _gram(rows, logz_obs, None, mask_n, 0.0, 1.0)
    elif stdev_obs is None:
        # var_n += min(var_n) + 1e-8; stdev_inv = 1/sqrt(var_n)  (distill_depth.py:44-46)
        G, h = _gram(rows, logz_obs, var_n, mask_n, float(var_min.item()) + 1e-8, 1.0)
    else:
        G, h = _gram(rows, logz_obs, None, mask_n, 0.0, 1.0 / (float(stdev_obs) ** 2))
    if with_prior:
        eye = torch.eye(m, device=dev, dtype=F64).unsqueeze(0)
        L_inv = torch.linalg.solve_triangular(L_mm, eye, upper=False)
        G = G + (L_inv.mT @ L_inv)[0]
    x = _chol_solve(G.unsqueeze(0), h.view(1, m, 1))
    n = rows.shape[0]
    res = torch.empty(n, dtype=F64, device=dev)
    stats = torch.empty(3, dtype=F64, device=dev)
    with torch.cuda.device(dev):
        st = _lib.rows_residual(_lib.ptr(rows), _lib.ptr(x.contiguous()), _lib.ptr(logz_obs), _lib.ptr(mask_n), n, m,
                                _lib.ptr(res), _lib.ptr(stats), _lib.stream_ptr(dev))
        _lib.check(st, "como_b200_rows_residual")
    return x, res, stats


def distill_conditional_masked(coords_m, z_m1, coords_n, logz_obs, log_median, mask_n, cov_params_img, scale, stdev_obs):
    """Core of distill_conditional_depth_from_scratch (distill_depth.py:126-175) on uncompacted test points."""
    dev = cov_params_img.device
    rows, _, _, _ = predictor_rows(coords_m, coords_n, mask_n, cov_params_img, scale, want_var=False)
    m = rows.shape[1]
    m1 = z_m1.shape[1]
    w = 1.0 / (float(stdev_obs) ** 2)
    G, h = _gram(rows, logz_obs, None, mask_n, 0.0, w)
    sp2 = (1.0 / 5e-2) ** 2
    ATA = G[m1:, m1:] + sp2 * torch.eye(m - m1, device=dev, dtype=F64)
    ATb = h[m1:] - G[m1:, :m1] @ torch.log(z_m1.to(F64)).reshape(m1) + sp2 * log_median
    return _chol_solve(ATA.unsqueeze(0), ATb.view(1, m - m1, 1))


def _compact_inputs(coords_n, z_obs, min_depth):
    """Reference-shaped inputs (already filtered coords (1,n,2), z_obs (1,n,1)) -> (logz (n,), mask (n,) uint8)."""
    z = z_obs.reshape(-1).to(F64)
    mask = (z > min_depth)
    logz = torch.where(mask, torch.log(torch.where(mask, z, torch.ones_like(z))), torch.zeros_like(z)).contiguous()
    return logz, mask.to(torch.uint8).contiguous()


def distill_depth_from_scratch(coords_m, coords_n, z_obs, cov_params_img, model, distill_with_prior, min_depth,
                               stdev_obs=None):
    """Drop-in for distill_depth.py:85-122.  Returns (logz_m (1,m,1), logz_residuals (1,n_valid,1))."""
    assert coords_m.shape[0] == 1
    logz, mask = _compact_inputs(coords_n, z_obs, min_depth)
    x, res, _ = distill_depth_masked(coords_m, coords_n, logz, mask, cov_params_img, _scale_of(model), distill_with_prior,
                                     stdev_obs)
    return x, res[mask.bool()].view(1, -1, 1)


def distill_conditional_depth_from_scratch(coords_m, z_m1, coords_n, cov_params_img, z_obs, model, min_depth, stdev_obs):
    """Drop-in for distill_depth.py:156-175.  Returns logz_m2 (1,m2,1)."""
    assert coords_m.shape[0] == 1
    logz, mask = _compact_inputs(coords_n, z_obs, min_depth)
    zv = z_obs.reshape(-1)[mask.bool()]
    log_median = torch.log(torch.median(zv))
    return distill_conditional_masked(coords_m, z_m1, coords_n, logz, log_median, mask, cov_params_img, _scale_of(model),
                                      stdev_obs)
