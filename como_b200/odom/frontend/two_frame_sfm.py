"""Two-frame SfM bootstrap: drop-ins for `two_frame_sfm_pyr`, `two_frame_sfm` and `setup_reference`
(como/odom/frontend/two_frame_sfm.py:15-392) on the sm_100a kernels (csrc/sfm.cu, kmat.cu, select.cu, chol.cu).

Per Gauss-Newton iteration: one linearisation kernel (predictor dot products on the FP64 tensor path, projection,
bilinear lookup, rank-one depth Jacobian), the exact median of |r|, one accumulation pass (Gram + 7 stack rows on
DMMA), the (6+M) x (6+M) system assembled from M x M pieces and solved by the tiled Cholesky.  The pose update and
the termination test need `delta` on the host once per iteration -- the reference synchronises there as well.
No CPU path.
"""
import ctypes as C
import math

import numpy as np
import torch

from como_b200 import _lib
from como_b200.odom.mapping_core import solve_system

F64 = torch.float64


def _se3_exp_wv(xi):
    """SE(3) exponential, COMO order [omega, v] (lie_algebra.py:52-56 feeds lietorch [v, omega]); numpy float64."""
    w, v = np.asarray(xi[:3], dtype=np.float64), np.asarray(xi[3:], dtype=np.float64)
    th2 = float(w @ w)
    if th2 < 1e-12:
        A, B, Cc = 1.0 - th2 / 6.0, 0.5 - th2 / 24.0, 1.0 / 6.0 - th2 / 120.0
    else:
        th = math.sqrt(th2)
        A, B, Cc = math.sin(th) / th, (1.0 - math.cos(th)) / th2, (th - math.sin(th)) / (th2 * th)
    Wm = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]], dtype=np.float64)
    WW = Wm @ Wm
    T = np.eye(4)
    T[:3, :3] = np.eye(3) + A * Wm + B * WW
    T[:3, 3] = (np.eye(3) + B * Wm + Cc * WW) @ v
    return T


def two_frame_sfm(Tji_init, sparse_log_depth_init, aff_init, test_coords_i, vals_i, Knm_Kmminv, img_and_grads_j,
                  dr_prior_dd, H_prior_d_d, intrinsics, sigmas, term_criteria, init_cfg):
    """One pyramid level (two_frame_sfm.py:306-392).  Same arguments / returns as the reference (batch size 1)."""
    dev = _lib.require_cuda(Tji_init, sparse_log_depth_init, test_coords_i, vals_i, Knm_Kmminv, img_and_grads_j)
    if Knm_Kmminv.shape[0] != 1 or vals_i.shape[1] != 1:
        raise RuntimeError("como_b200 two_frame_sfm: batch size 1 and gray images only")
    N, M = int(Knm_Kmminv.shape[1]), int(Knm_Kmminv.shape[2])
    H, W = int(img_and_grads_j.shape[-2]), int(img_and_grads_j.shape[-1])
    Knm = Knm_Kmminv[0].to(F64).contiguous()
    coords = test_coords_i[0].to(torch.int64).contiguous()
    vals = vals_i.reshape(-1).to(F64).contiguous()
    img = img_and_grads_j[0].to(F64).contiguous()
    Kh = intrinsics.detach().to("cpu", F64)
    intr4 = (C.c_double * 4)(float(Kh[0, 0]), float(Kh[1, 1]), float(Kh[0, 2]), float(Kh[1, 2]))
    dr_prior = dr_prior_dd[0].to(F64)
    H_prior = H_prior_d_d[0].to(F64)
    T = Tji_init[0].detach().to("cpu", F64).numpy().copy()
    d = sparse_log_depth_init[0].to(F64).reshape(M).clone().contiguous()
    aff = aff_init.clone()

    rec = torch.empty(N, 8, dtype=F64, device=dev)
    absr = torch.empty(N, dtype=F64, device=dev)
    proj = torch.empty(N, 3, dtype=F64, device=dev)
    stats = torch.empty(2, dtype=F64, device=dev)
    sigma = torch.empty(1, dtype=F64, device=dev)
    G = torch.empty(M, M, dtype=F64, device=dev)
    St = torch.empty(7, M, dtype=F64, device=dev)
    small = torch.empty(28, dtype=F64, device=dev)
    seg = torch.tensor([0, N], dtype=torch.int64, device=dev)
    iu = torch.triu_indices(6, 6, device=dev)
    with torch.cuda.device(dev):
        stream = _lib.stream_ptr(dev)
        mws = torch.empty(int(_lib.median_workspace_bytes(1, 8)), dtype=torch.uint8, device=dev)
        dr_mean = (Knm.sum(0) / N).reshape(1, M)                   # linearize_mean_log_depth_prior_system (:127-133)
        H_mean = dr_mean.T @ dr_mean
        it, prev = 0, float("inf")
        while True:
            T12 = (C.c_double * 12)(*T[:3, :].reshape(-1).tolist())
            st = _lib.sfm_linearize(_lib.ptr(Knm), _lib.ptr(d), _lib.ptr(coords), _lib.ptr(vals), _lib.ptr(img), H, W, N, M, T12,
                                    intr4, _lib.ptr(rec), _lib.ptr(absr), _lib.ptr(proj), _lib.ptr(stats), stream)
            _lib.check(st, "como_b200_sfm_linearize")
            st = _lib.median_f64(_lib.ptr(absr), _lib.ptr(seg), 1, N, 1.4826, _lib.ptr(sigma), None, _lib.ptr(mws), mws.numel(),
                                 stream)
            _lib.check(st, "como_b200_median_f64")
            st = _lib.sfm_accumulate(_lib.ptr(Knm), _lib.ptr(rec), _lib.ptr(sigma), N, M, _lib.ptr(G), _lib.ptr(St),
                                     _lib.ptr(small), stream)
            _lib.check(st, "como_b200_sfm_accumulate")
            # assemble the (6 + M) system: photometric blocks + sparse-depth prior (:136-146) + mean-log-depth prior (:149-163)
            Hm = torch.zeros(6 + M, 6 + M, dtype=F64, device=dev)
            g = torch.zeros(6 + M, dtype=F64, device=dev)
            HTT = torch.zeros(6, 6, dtype=F64, device=dev)
            HTT[iu[0], iu[1]] = small[:21]
            HTT = HTT + torch.triu(HTT, 1).T
            rp = dr_prior @ d
            mean_logz = stats[0] / N
            Hm[:6, :6] = HTT
            Hm[:6, 6:] = St[:6]
            Hm[6:, :6] = St[:6].T
            Hm[6:, 6:] = G + H_prior + H_mean
            g[:6] = -small[21:27]
            g[6:] = -St[6] - dr_prior.T @ rp - dr_mean[0] * mean_logz
            total_err_t = small[27] + torch.sum(rp * rp) + mean_logz * mean_logz
            delta = solve_system(Hm, g)
            host = torch.cat((delta.reshape(-1), total_err_t.reshape(1))).cpu().numpy()   # the one sync of the iteration
            dh, total_err = host[:-1], float(host[-1])
            T = T @ _se3_exp_wv(dh[:6])                             # update_vars / batch_se3
            d = d + delta[6:, 0]
            it += 1
            delta_norm = float(np.linalg.norm(dh[:6]))
            abs_decrease = prev - total_err
            rel_decrease = abs(abs_decrease) / prev if prev != 0.0 else float("inf")
            if math.isinf(prev):
                rel_decrease = float("nan")                         # inf / inf in the reference's tensor arithmetic
            done = (it >= init_cfg["max_iter"] or delta_norm < init_cfg["delta_norm"]
                    or (rel_decrease < init_cfg["rel_tol"] and abs_decrease > 0))
            prev = total_err
            if done:
                break
    valid = ~torch.isnan(rec[:, 0])
    coords_j = torch.stack((proj[valid, 1], proj[valid, 0]), -1).unsqueeze(0)      # swap_coords_xy(pj): (row, col)
    depths_j = proj[valid, 2].reshape(1, -1, 1)
    Tji = torch.from_numpy(T).to(dev).unsqueeze(0)
    mean_log_depth = (stats[0] / N).reshape(1, 1, 1)
    two_frame_sfm.last_iters = it
    return Tji, d.reshape(1, M, 1), aff, coords_j, depths_j, mean_log_depth


def two_frame_sfm_pyr(Tji_init, sparse_log_depth_init, aff_init, test_coords_i, vals_i, Knm_Kmminv, img_and_grads_j,
                      dr_prior_dd, H_prior_d_d, intrinsics, sigmas, term_criteria, init_cfg):
    """Coarse to fine (two_frame_sfm.py:15-52); the list arguments are per level, coarsest first."""
    Tji, d, aff = Tji_init.clone(), sparse_log_depth_init.clone(), aff_init.clone()
    out = None
    two_frame_sfm_pyr.last_iters = []
    for l in range(len(vals_i)):
        out = two_frame_sfm(Tji, d, aff, test_coords_i[l], vals_i[l], Knm_Kmminv[l], img_and_grads_j[l], dr_prior_dd,
                            H_prior_d_d, intrinsics[l], sigmas, term_criteria, init_cfg)
        Tji, d, aff = out[0], out[1], out[2]
        two_frame_sfm_pyr.last_iters.append(two_frame_sfm.last_iters)
    return out


def setup_reference(img_and_grads, sparse_coords_norm, model, cov_params_img, intrinsics):
    """Drop-in for setup_reference (two_frame_sfm.py:55-112): per level the reference intensities, the pixel list and
    the predictor K_nm K_mm^-1 (fused K-matrix kernel at the level's pixel centres mapped into the covariance image),
    plus the sparse-depth prior linearisation.  Pixels are listed in natural order (the reference draws a random
    permutation of ALL pixels with torch.multinomial; every consumer sums over pixels)."""
    from como_b200.depth_cov.core import distill_depth as DD
    dev = _lib.require_cuda(img_and_grads[-1], sparse_coords_norm, cov_params_img, intrinsics)
    if img_and_grads[-1].shape[-3] != 3:
        raise NotImplementedError("como_b200 setup_reference: gray images only (3 channels = I, gx, gy)")
    dtype = img_and_grads[-1].dtype
    Hc, Wc = int(cov_params_img.shape[-2]), int(cov_params_img.shape[-1])
    scale = DD._scale_of(model)
    dims = torch.tensor([Hc, Wc], dtype=F64, device=dev)
    coords_m = ((sparse_coords_norm.to(F64) + 1.0) * dims - 1.0) / 2.0          # unnormalize_coordinates (utils/coords.py:23-26)
    M = coords_m.shape[1]
    cov = cov_params_img.to(F64).contiguous()
    E_m = torch.empty(1, M, 4, dtype=F64, device=dev)
    K_mm = torch.empty(1, M, M, dtype=F64, device=dev)
    with torch.cuda.device(dev):
        st = _lib.kmat_kmm(_lib.ptr(cov), 1, Hc, Wc, _lib.ptr(coords_m.contiguous()), M, float(scale), 0.0, _lib.ptr(E_m),
                           _lib.ptr(K_mm), _lib.stream_ptr(dev))
        _lib.check(st, "como_b200_kmat_kmm")
    L_mm, _ = torch.linalg.cholesky_ex(K_mm, upper=False)
    eye = torch.eye(M, device=dev, dtype=F64).unsqueeze(0)
    dr_prior_dd = torch.linalg.solve_triangular(L_mm, eye, upper=False)         # linearize_sparse_depth_prior (:115-124)
    H_prior_d_d = dr_prior_dd.mT @ dr_prior_dd
    Kinv = torch.cholesky_solve(eye, L_mm, upper=False).contiguous()
    # IntrinsicsPyramidModule(0, levels)(K, [1, 1]) with the reference's resize_intrinsics (geometry/camera.py:4-16):
    # level i uses T_i @ K with T_i = [[s,0,s],[0,s,s],[0,0,1]], s = 2^-i -- the scale is ADDED to the principal point
    levels = len(img_and_grads)
    Kp = []
    for i in range(levels):
        sc = 2.0 ** (-i)
        Tm = torch.tensor([[sc, 0, sc], [0, sc, sc], [0, 0, 1.0]], device=dev, dtype=intrinsics.dtype)
        Kp.insert(0, Tm @ intrinsics)
    vals_pyr, coords_pyr, Knm_pyr, sizes = [], [], [], []
    for l in range(levels):
        img = img_and_grads[l][:, :1]
        h, w = int(img.shape[-2]), int(img.shape[-1])
        rr, cc = torch.meshgrid(torch.arange(h, device=dev), torch.arange(w, device=dev), indexing="ij")
        tc = torch.stack((rr.reshape(-1), cc.reshape(-1)), -1).unsqueeze(0)
        coords_pyr.append(tc)
        vals_pyr.append(img.reshape(1, 1, h * w))
        # pixel centre (r, c) of the level in covariance-image pixel units: normalise with (h, w), unnormalise with (Hc, Wc)
        full = torch.stack(((rr.reshape(-1).to(F64) + 0.5) * (Hc / h) - 0.5, (cc.reshape(-1).to(F64) + 0.5) * (Wc / w) - 0.5), -1)
        rows = torch.empty(h * w, M, dtype=F64, device=dev)
        with torch.cuda.device(dev):
            st = _lib.kmat_rows(_lib.ptr(cov), 1, Hc, Wc, _lib.ptr(coords_m.contiguous()), _lib.ptr(E_m), _lib.ptr(Kinv), M,
                                float(scale), _lib.ptr(full.contiguous()), None, h * w, _lib.ptr(rows), None, None,
                                _lib.stream_ptr(dev))
            _lib.check(st, "como_b200_kmat_rows")
        Knm_pyr.append(rows.unsqueeze(0).to(dtype))
        sizes.append(img.shape[-2:])
    return vals_pyr, coords_pyr, Knm_pyr, sizes, Kp, dr_prior_dd.to(dtype), H_prior_d_d.to(dtype)
