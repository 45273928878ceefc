"""Keyframe creation: drop-in for `track_and_init` (como/odom/frontend/corr.py:60-242) on the sm_100a kernels.

Dense part (every pixel of the last keyframe): one reprojection kernel, the fused K-matrix/predictor kernel with
variance, the DMMA normal-equation kernel and a streaming residual kernel (csrc/kfinit.cu, csrc/kmat.cu) -- the
n x m matrices of the reference are written once (predictor rows) and read twice.  Sparse part (<= 64 anchors):
a few small torch tensor ops (device plumbing) around the greedy sampler (csrc/depthcov.cu).  No CPU path.
"""
import ctypes as C

import torch

from como_b200 import _lib
from como_b200.depth_cov.core import distill_depth as DD
from como_b200.depth_cov.core.samplers import sample_sparse_coords

F64 = torch.float64


def _swap(c):
    return torch.stack((c[..., 1], c[..., 0]), -1)


def _backproject(K, p_xy, z):
    rx = (p_xy[..., 0] - K[0, 2]) / K[0, 0]
    ry = (p_xy[..., 1] - K[1, 2]) / K[1, 1]
    return torch.stack((rx, ry, torch.ones_like(rx)), -1) * z


def _project(K, P):
    return torch.stack((K[0, 0] * P[..., 0] / P[..., 2] + K[0, 2], K[1, 1] * P[..., 1] / P[..., 2] + K[1, 2]), -1)


def reproject_points(coords_i, zi, Tji, K):
    """corr.py:37-43 for a handful of sparse points (row, col)."""
    Pi = _backproject(K[0], _swap(coords_i).to(F64), zi)
    Pj = Pi @ Tji[:, :3, :3].transpose(1, 2) + Tji[:, None, :3, 3]
    return _swap(_project(K[0], Pj)), Pj


def _inv_se3(T):
    R = T[:, :3, :3]
    Ti = torch.eye(4, dtype=T.dtype, device=T.device).repeat(T.shape[0], 1, 1)
    Ti[:, :3, :3] = R.transpose(1, 2)
    Ti[:, :3, 3] = -(R.transpose(1, 2) @ T[:, :3, 3:4])[..., 0]
    return Ti


def _in_bounds(coords, P, img_size, min_depth):
    ok = (coords[0, :, 1] >= 1) & (coords[0, :, 1] < img_size[1] - 1) & (coords[0, :, 0] >= 1) & (coords[0, :, 0] < img_size[0] - 1)
    return ok & (P[0, :, 2] > min_depth)


def _corr_errors(Pa, Pb, mode):
    if mode == "z":
        return torch.abs(Pa[..., 2:3] - Pb[..., 2:3])
    if mode in ("logz", "logr"):
        return torch.abs(torch.log(Pa[..., 2:3]) - torch.log(Pb[..., 2:3]))
    if mode == "3d":
        return torch.linalg.norm(Pa - Pb, dim=-1, keepdim=True)
    raise ValueError("corr_mode: " + str(mode))


def _median_lower(values):
    """Exact lower median (torch.median semantics) of a 1-D float64 CUDA tensor with the radix-select kernel."""
    dev = values.device
    n = values.numel()
    seg = torch.tensor([0, n], dtype=torch.int64, device=dev)
    out = torch.empty(1, dtype=F64, device=dev)
    with torch.cuda.device(dev):
        ws = torch.empty(int(_lib.median_workspace_bytes(1, 8)), dtype=torch.uint8, device=dev)
        st = _lib.median_f64(_lib.ptr(values), _lib.ptr(seg), 1, n, 1.0, _lib.ptr(out), None, _lib.ptr(ws), ws.numel(),
                             _lib.stream_ptr(dev))
        _lib.check(st, "como_b200_median_f64")
    return out[0]


def track_and_init(pose1, pose2, coords_m1, z_m1, z_img1, cov_params_img2, K, model, corr_params, sampling_params,
                   rgb_img_size, rgb1=None, rgb2=None, debug=None):
    """Same arguments and returns as the reference: (coords_2, z2, corr_mask, coords_all, z_all).
    `model` is the DepthCov module (only `get_scale(-1)` is used) or the scale as a float."""
    dev = _lib.require_cuda(pose1, pose2, coords_m1, z_m1, z_img1, cov_params_img2, K)
    b, _, H, W = cov_params_img2.shape
    if b != 1:
        raise RuntimeError("como_b200 track_and_init: batch size must be 1 (as the reference asserts)")
    if tuple(z_img1.shape[-2:]) != (H, W):
        raise NotImplementedError("como_b200 track_and_init expects depth and covariance images of the same size")
    scale = DD._scale_of(model)
    N = H * W
    min_depth = float(corr_params["min_obs_depth"])
    pose1, pose2, K = pose1.to(F64), pose2.to(F64), K.to(F64)
    z_img = z_img1.to(F64).contiguous()
    cov = cov_params_img2.to(F64).contiguous()
    Tji = _inv_se3(pose2) @ pose1

    # ---- dense reprojection of the last keyframe's depth image
    coords_j_n = torch.empty(1, N, 2, dtype=F64, device=dev)
    logz_n = torch.empty(N, dtype=F64, device=dev)
    zj_n = torch.empty(N, dtype=F64, device=dev)
    mask_n = torch.empty(N, dtype=torch.uint8, device=dev)
    host = torch.cat((Tji[0, :3, :].reshape(-1), torch.stack((K[0, 0, 0], K[0, 1, 1], K[0, 0, 2], K[0, 1, 2])))).cpu()
    T12 = (C.c_double * 12)(*host[:12].tolist())
    intr4 = (C.c_double * 4)(*host[12:].tolist())
    with torch.cuda.device(dev):
        st = _lib.reproject_dense(_lib.ptr(z_img), H, W, T12, intr4, min_depth, _lib.ptr(coords_j_n), _lib.ptr(logz_n),
                                  _lib.ptr(zj_n), _lib.ptr(mask_n), _lib.stream_ptr(dev))
        _lib.check(st, "como_b200_reproject_dense")

    # ---- sparse anchors of the last keyframe in the new frame
    coords_j_m1, Pj_m1 = reproject_points(coords_m1, z_m1.to(F64), Tji, K)
    mask_m1 = _in_bounds(coords_j_m1, Pj_m1, (H, W), min_depth)
    coords_j_m1_f, Pj_m1_f = coords_j_m1[:, mask_m1], Pj_m1[:, mask_m1]

    # ---- latent depths of the reprojected anchors under the new frame's covariance
    logz_m, res, stats = DD.distill_depth_masked(coords_j_m1_f, coords_j_n, logz_n, mask_n, cov, scale,
                                                 bool(corr_params["distill_with_prior"]))
    z_m = torch.exp(logz_m)
    P_m = _backproject(K[0], _swap(coords_j_m1_f), z_m)

    # ---- two-way check: back into the last keyframe, compare with its dense depth; reject depth edges
    coords_i_m1, Pi_m1 = reproject_points(coords_j_m1_f, z_m, _inv_se3(Tji), K)
    nm = coords_j_m1_f.shape[1]
    z_proj = torch.empty(nm, dtype=F64, device=dev)
    grad_ref = torch.empty(nm, dtype=F64, device=dev)
    coords_m1_f = coords_m1[:, mask_m1].to(F64)
    with torch.cuda.device(dev):
        st = _lib.sample_depth_gradmag(_lib.ptr(z_img), H, W, _lib.ptr(coords_i_m1.contiguous()), _lib.ptr(coords_m1_f.contiguous()),
                                       nm, _lib.ptr(z_proj), _lib.ptr(grad_ref), _lib.stream_ptr(dev))
        _lib.check(st, "como_b200_sample_depth_gradmag")
    P_proj = _backproject(K[0], _swap(coords_i_m1), z_proj.view(1, nm, 1))
    mode = corr_params["corr_mode"]
    corr_errors = torch.maximum(_corr_errors(P_proj, Pi_m1, mode), _corr_errors(Pj_m1_f, P_m, mode))
    keep = ((corr_errors < corr_params["corr_thresh"])[0, :, 0]) & (grad_ref < corr_params["logz_grad_mag_thresh"])

    coords_1 = coords_j_m1_f[:, keep]
    z1 = Pj_m1_f[:, keep, 2:3]
    signal_var = scale
    if coords_1.shape[1] > 0:
        _, inds = sample_sparse_coords(cov, sampling_params["max_num_coords"], "greedy_conditional_entropy",
                                       sampling_params["max_stdev_thresh"], border=sampling_params["border"],
                                       terminate_early=True, dist_thresh=sampling_params["dist_thresh"],
                                       signal_var=signal_var, fixed_var=sampling_params["fixed_var"], coords_domain=coords_1)
        sampled = torch.zeros(coords_1.shape[1], device=dev, dtype=torch.bool)
        sampled[inds[0, :]] = True
        coords_1, z1 = coords_1[:, sampled], z1[:, sampled]
        keep = keep.clone()
        keep[keep.clone()] = sampled
        if debug is not None:
            debug["ss0_inds"] = inds
    corr_mask = mask_m1.clone()
    corr_mask[mask_m1] = keep

    cnt = stats[0]
    sigma_r = torch.sqrt(stats[2] / (cnt - 1.0))
    if debug is not None:
        debug.update(dd_logz_m=logz_m, dd_res_std=float(sigma_r), dd_n=int(cnt.item()), dd_coords_m=coords_j_m1_f)

    if coords_1.shape[1] < sampling_params["max_num_coords"]:
        coords_2, inds2 = sample_sparse_coords(cov, sampling_params["max_num_coords"], sampling_params["mode"],
                                               sampling_params["max_stdev_thresh"], border=sampling_params["border"],
                                               terminate_early=False, dist_thresh=sampling_params["dist_thresh"],
                                               signal_var=signal_var, fixed_var=sampling_params["fixed_var"],
                                               curr_coords=coords_1)
        coords_2 = coords_2.to(dtype=coords_1.dtype)
        coords_all = torch.cat((coords_1, coords_2), dim=1)
        # conditional distillation uses min_depth = 0.0 on the same (already filtered) observations
        mask2 = mask_n if min_depth >= 0.0 else (mask_n.bool() & (zj_n > 0.0)).to(torch.uint8)
        log_median = torch.log(_median_lower(zj_n[mask2.bool()].contiguous()))
        logz_2 = DD.distill_conditional_masked(coords_all, z1, coords_j_n, logz_n, log_median, mask2, cov, scale, float(sigma_r))
        z2 = torch.exp(logz_2)
        z_all = torch.cat((z1, z2), dim=1)
        if debug is not None:
            debug.update(ss1_inds=inds2, dc_logz_2=logz_2)
    else:
        coords_all, z_all = coords_1.clone(), z1.clone()
        coords_2 = torch.empty((1, 0, 2), device=dev, dtype=coords_1.dtype)
        z2 = torch.empty((1, 0, 1), device=dev, dtype=z1.dtype)
    return coords_2, z2, corr_mask, coords_all, z_all
