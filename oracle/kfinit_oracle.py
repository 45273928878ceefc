"""ORACLE (test infrastructure, never on the product path): keyframe-creation path on the CPU.

Restates, in plain torch-CPU float64,
* calc_kernel_matrices / get_predictor / distill_depth / distill_conditional_depth_with_scale_prior
  (como/depth_cov/core/distill_depth.py:8-175) with lstsq_chol (como/utils/lin_alg.py:82-87),
* track_and_init (como/odom/frontend/corr.py:60-242): reprojection of sparse and dense points, depth
  distillation in the new frame, two-way correspondence check, anchor re-selection / new anchors through the
  greedy sampler (oracle/depthcov_oracle.sample_sparse_coords) and conditional depth initialisation.
Pinned by tests/golden/kfinit_64x48.npz, generated from the unmodified reference (oracle/gen_golden.py kfinit).
"""
import math

import numpy as np
import torch

from oracle import depthcov_oracle as DO

F64 = torch.float64


def _swap(c):
    return torch.stack((c[..., 1], c[..., 0]), -1)


def backproject(K, p_xy, z):
    """camera.py:45-56: p (B,N,2) [x,y], z (B,N,1) -> P (B,N,3)."""
    rx = (p_xy[..., 0] - K[0, 2]) / K[0, 0]
    ry = (p_xy[..., 1] - K[1, 2]) / K[1, 1]
    return torch.stack((rx, ry, torch.ones_like(rx)), -1) * z


def project(K, P):
    """camera.py:20-40."""
    return torch.stack((K[0, 0] * P[..., 0] / P[..., 2] + K[0, 2], K[1, 1] * P[..., 1] / P[..., 2] + K[1, 2]), -1)


def reproject(coords_rc, z, T, K):
    """corr.py:37-43: pixel (row,col) + depth in frame i -> (row,col) and 3-D point in frame j."""
    Pi = backproject(K, _swap(coords_rc).to(F64), z)
    Pj = Pi @ T[:, :3, :3].transpose(1, 2) + T[:, None, :3, 3]
    return _swap(project(K, Pj)), Pj


def inv_se3(T):
    R = T[:, :3, :3]
    Ti = torch.eye(4, dtype=T.dtype).repeat(T.shape[0], 1, 1)
    Ti[:, :3, :3] = R.transpose(1, 2)
    Ti[:, :3, 3] = -(R.transpose(1, 2) @ T[:, :3, 3:4])[..., 0]
    return Ti


def in_bounds(coords, P, img_size, min_depth):
    """corr.py:17-29."""
    ok = (coords[0, :, 1] >= 1) & (coords[0, :, 1] < img_size[1] - 1) & (coords[0, :, 0] >= 1) & (coords[0, :, 0] < img_size[0] - 1)
    return ok & (P[0, :, 2] > min_depth)


def kernel_matrices(coords_m, coords_n, cov_img, scale):
    """distill_depth.py:8-27 (no jitter on K_mm here)."""
    size = cov_img.shape[-2:]
    cm = DO._normalize(coords_m.to(F64), size)
    cn = DO._normalize(coords_n.to(F64), size)
    E_m, E_n = DO._interp_cov(cov_img, cm), DO._interp_cov(cov_img, cn)
    K_mm = DO.cov_python(cm, E_m, cm, E_m, scale)
    K_nm = DO.cov_python(cn, E_n, cm, E_m, scale)
    det = E_n[..., 0, 0] * E_n[..., 1, 1] - E_n[..., 0, 1] * E_n[..., 1, 0]
    det2 = (2 * E_n[..., 0, 0]) * (2 * E_n[..., 1, 1]) - (2 * E_n[..., 0, 1]) * (2 * E_n[..., 1, 0])
    t0 = math.sqrt(3.0) * math.sqrt(1e-8)
    K_nn = 2.0 * torch.sqrt(det) / torch.sqrt(det2 + 1e-8) * ((1 + t0) * math.exp(-t0)) * scale
    return K_mm, K_nm, K_nn


def predictor(K_mm, K_nm, K_nn):
    """distill_depth.py:30-48."""
    L, _ = torch.linalg.cholesky_ex(K_mm)
    M = L.shape[-1]
    Kinv = torch.cholesky_solve(torch.eye(M, dtype=F64)[None], L)
    KK = K_nm @ Kinv
    var = K_nn - torch.sum(K_nm * KK, dim=2)
    var = var + (torch.min(var) + 1e-8)
    return KK, L, (1.0 / torch.sqrt(var)).unsqueeze(-1)


def lstsq_chol(A, b):
    ATA, ATb = A.mT @ A, A.mT @ b
    L, _ = torch.linalg.cholesky_ex(ATA)
    return torch.cholesky_solve(ATb, L)


def distill_depth_from_scratch(coords_m, coords_n, z_obs, cov_img, scale, with_prior, min_depth):
    """distill_depth.py:85-122 (+51-82)."""
    KK, L, sinv = predictor(*kernel_matrices(coords_m, coords_n, cov_img, scale))
    ok = z_obs[0, :, 0] > min_depth
    KK, z, sinv = KK[:, ok], z_obs[:, ok], sinv[:, ok]
    logz = torch.log(z)
    m = KK.shape[2]
    if not with_prior:
        x = lstsq_chol(KK, logz)
    else:
        Linv = torch.linalg.solve_triangular(L, torch.eye(m, dtype=F64)[None], upper=False)
        A = torch.cat((Linv, sinv * KK), 1)
        b = torch.cat((torch.zeros(1, m, 1, dtype=F64), sinv * logz), 1)
        x = lstsq_chol(A, b)
    return x, KK @ x - logz


def distill_conditional_from_scratch(coords_m, z_m1, coords_n, cov_img, z_obs, scale, min_depth, stdev_obs):
    """distill_depth.py:156-175 (+126-153)."""
    KK, L, _ = predictor(*kernel_matrices(coords_m, coords_n, cov_img, scale))
    ok = z_obs[0, :, 0] > min_depth
    KK, z = KK[:, ok], z_obs[:, ok]
    sinv = 1.0 / stdev_obs
    m1 = z_m1.shape[1]
    m2 = KK.shape[2] - m1
    s = torch.log(torch.median(z))
    sp = 1.0 / 5e-2
    A = torch.cat((sp * torch.eye(m2, dtype=F64)[None], sinv * KK[:, :, m1:]), 1)
    b = torch.cat((sp * s * torch.ones(1, m2, 1, dtype=F64), sinv * (torch.log(z) - KK[:, :, :m1] @ torch.log(z_m1))), 1)
    return lstsq_chol(A, b)


def _grid_sample_zeros(img, coords_rc):
    """F.grid_sample(bilinear, zeros padding, align_corners=False) at pixel (row,col) coords; img (1,1,H,W)."""
    H, W = img.shape[-2:]
    cn = DO._normalize(coords_rc, (H, W))
    grid = torch.stack((cn[..., 1], cn[..., 0]), -1).unsqueeze(1)
    s = torch.nn.functional.grid_sample(img, grid, mode="bilinear", padding_mode="zeros", align_corners=False)
    return s.reshape(1, 1, -1).permute(0, 2, 1)


def scharr_mag(x):
    """ImageGradientModule (utils/image_processing.py:8-45) on a (1,1,H,W) image -> sqrt(gx^2 + gy^2)."""
    kx = torch.tensor([[-3.0, 0, 3], [-10, 0, 10], [-3, 0, 3]], dtype=x.dtype) / 32.0
    ky = kx.t().contiguous()
    xp = torch.nn.functional.pad(x, (1, 1, 1, 1), mode="reflect")
    gx = torch.nn.functional.conv2d(xp, kx[None, None])
    gy = torch.nn.functional.conv2d(xp, ky[None, None])
    return torch.sqrt(gx * gx + gy * gy)


def track_and_init(pose1, pose2, coords_m1, z_m1, z_img1, cov_img2, K, scale, corr, samp, debug=None):
    """corr.py:60-242.  Returns (coords_2, z2, corr_mask, coords_all, z_all)."""
    K = K[0]
    H, W = cov_img2.shape[-2:]
    N = H * W
    Tji = inv_se3(pose2) @ pose1
    rr, cc = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    coords_n1 = torch.stack((rr.reshape(-1), cc.reshape(-1)), 1)[None]
    z_n1 = z_img1.reshape(1, 1, N).permute(0, 2, 1)
    cj_m, Pj_m = reproject(coords_m1, z_m1, Tji, K)
    cj_n, Pj_n = reproject(coords_n1, z_n1, Tji, K)
    ok_m = in_bounds(cj_m, Pj_m, (H, W), corr["min_obs_depth"])
    ok_n = in_bounds(cj_n, Pj_n, (H, W), corr["min_obs_depth"])
    cj_m, Pj_m = cj_m[:, ok_m], Pj_m[:, ok_m]
    cj_n, zj_n = cj_n[:, ok_n], Pj_n[:, ok_n, 2:3]

    logz_m, res = distill_depth_from_scratch(cj_m, cj_n, zj_n, cov_img2, scale, corr["distill_with_prior"], corr["min_obs_depth"])
    if debug is not None:
        debug.update(dd_logz_m=logz_m, dd_res_std=float(torch.std(res)), dd_n=int(res.shape[1]), dd_coords_m=cj_m)
    z_m = torch.exp(logz_m)
    P_m = backproject(K, _swap(cj_m), z_m)
    ci_m, Pi_m = reproject(cj_m, z_m, inv_se3(Tji), K)
    z_proj = _grid_sample_zeros(z_img1, ci_m)
    P_proj = backproject(K, _swap(ci_m), z_proj)
    gmag = scharr_mag(torch.log(z_img1))
    g_ref = _grid_sample_zeros(gmag, coords_m1[:, ok_m])

    def err(Pa, Pb):
        mode = corr["corr_mode"]
        if mode == "z":
            return torch.abs(Pa[..., 2:3] - Pb[..., 2:3])
        if mode in ("logz", "logr"):
            return torch.abs(torch.log(Pa[..., 2:3]) - torch.log(Pb[..., 2:3]))
        return torch.linalg.norm(Pa - Pb, dim=-1, keepdim=True)

    e = torch.maximum(err(P_proj, Pi_m), err(Pj_m, P_m))
    keep = ((e < corr["corr_thresh"]) & (g_ref < corr["logz_grad_mag_thresh"]))[0, :, 0]
    coords_1, z1 = cj_m[:, keep], Pj_m[:, keep, 2:3]
    if coords_1.shape[1] > 0:
        _, inds = DO.sample_sparse_coords(cov_img2, samp["max_num_coords"], samp["max_stdev_thresh"], border=samp["border"],
                                          terminate_early=True, dist_thresh=samp["dist_thresh"], signal_var=scale,
                                          fixed_var=samp["fixed_var"], coords_domain=coords_1)
        if debug is not None:
            debug.update(ss0_inds=inds)
        sampled = torch.zeros(coords_1.shape[1], dtype=torch.bool)
        sampled[inds[0]] = True
        coords_1, z1 = coords_1[:, sampled], z1[:, sampled]
        keep = keep.clone()
        keep[keep.clone()] = sampled
    corr_mask = ok_m.clone()
    corr_mask[ok_m] = keep
    if coords_1.shape[1] < samp["max_num_coords"]:
        coords_2, inds2 = DO.sample_sparse_coords(cov_img2, samp["max_num_coords"], samp["max_stdev_thresh"], border=samp["border"],
                                                  terminate_early=False, dist_thresh=samp["dist_thresh"], signal_var=scale,
                                                  fixed_var=samp["fixed_var"], curr_coords=coords_1)
        if debug is not None:
            debug.update(ss1_inds=inds2)
        coords_2 = coords_2.to(F64)
        sigma_r = torch.std(res)
        coords_all = torch.cat((coords_1, coords_2), 1)
        logz_2 = distill_conditional_from_scratch(coords_all, z1, cj_n, cov_img2, zj_n, scale, 0.0, sigma_r)
        if debug is not None:
            debug.update(dc_logz_2=logz_2)
        z2 = torch.exp(logz_2)
        z_all = torch.cat((z1, z2), 1)
    else:
        coords_all, z_all = coords_1.clone(), z1.clone()
        coords_2, z2 = torch.empty(1, 0, 2, dtype=F64), torch.empty(1, 0, 1, dtype=F64)
    return coords_2, z2, corr_mask, coords_all, z_all
