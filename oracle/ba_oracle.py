"""ORACLE (test infrastructure, never on the product path): CPU restatement (torch, fp64) of one
window bundle-adjustment Gauss-Newton iteration of COMO's mapper, `Mapping.iterate`.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this.
Pinned against the live reference by tests/golden/ba_*.npz (oracle/gen_golden.py drives the unmodified
reference Mapping, real DepthCov UNet + sampler included, and records H, g, delta and the updated state);
see tests/test_oracle_ba.py.

The restatement is NOT a transcription: it uses the rank-1 structure of the anchor-depth Jacobian
(d r / d z_m = alpha_n * K[n,m] / z_m) so the reference's (b,N,3,M,1) tensor is never formed -- the
same algebra the CUDA kernels implement.  Follows (reference file:line):
  Mapping.iterate / store_vars / setup_system       como/odom/Mapping.py:701-968
  prep_geometry_scaffold / prep_dense_ref           como/odom/Mapping.py:603-699
  get_batch_remap_function, project_landmarks,
  subselect_pixels, backproject_cloud, setup_test_points   como/odom/backend/sparse_map.py:18-230
  setup_photometric_pairs (+ temporal neighbours)   como/odom/backend/graph_pair_construction.py:5-182
  create_photo_system / batch_photo_cost / interp_img / robustify   como/odom/backend/photo.py:24-353
  gradient / block reductions / scatter / solve / update  como/odom/backend/linear_system.py:6-152
  gp_ml_cost, mean_log_depth_cost                   como/odom/factors/gp_priors.py:7-150
  log_depth_prior (mode first_mean)                 como/odom/factors/depth_prior.py:7-141
  pixel_prior_cost (mode first)                     como/odom/factors/pixel_prior.py:6-130
  linearize_pose_prior, SE3_logmap quirk            como/odom/factors/pose_prior_factors.py:5-19, geometry/lie_algebra.py:117-176
  linearize_scalar_prior / multi                    como/odom/factors/scalar_prior_factors.py:4-34
"""
import math

import torch

from oracle.track_oracle import lower_median

HUBER_K = 1.345
F64 = torch.float64


def skew(v):
    z = torch.zeros_like(v[..., 0])
    return torch.stack(
        (torch.stack((z, -v[..., 2], v[..., 1]), -1), torch.stack((v[..., 2], z, -v[..., 0]), -1),
         torch.stack((-v[..., 1], v[..., 0], z), -1)), -2)


def se3_exp_batch(delta_wv):
    """delta (B,6) in COMO order [omega, v] -> (B,4,4); lietorch convention restated (see track_oracle)."""
    w, v = delta_wv[:, :3], delta_wv[:, 3:]
    th2 = (w * w).sum(-1)
    th = th2.sqrt()
    small = th2 < 1e-12
    ths = torch.where(small, torch.ones_like(th), th)
    A = torch.where(small, 1 - th2 / 6, torch.sin(ths) / ths)
    B = torch.where(small, 0.5 - th2 / 24, (1 - torch.cos(ths)) / (ths * ths))
    C = torch.where(small, 1.0 / 6 - th2 / 120, (ths - torch.sin(ths)) / ths ** 3)
    W = skew(w)
    WW = W @ W
    I = torch.eye(3, dtype=delta_wv.dtype).expand_as(W)
    T = torch.zeros(w.shape[0], 4, 4, dtype=delta_wv.dtype)
    T[:, :3, :3] = I + A[:, None, None] * W + B[:, None, None] * WW
    T[:, :3, 3] = ((I + B[:, None, None] * W + C[:, None, None] * WW) @ v[..., None])[..., 0]
    T[:, 3, 3] = 1
    return T


def inv_se3(T):
    Ti = torch.zeros_like(T)
    Rt = T[..., :3, :3].transpose(-1, -2)
    Ti[..., :3, :3] = Rt
    Ti[..., :3, 3] = -(Rt @ T[..., :3, 3:4])[..., 0]
    Ti[..., 3, 3] = 1
    return Ti


def se3_log_quirk(T):
    """Reference SE3_logmap incl. its elementwise `(0.5*t)*(w^ x t)` term (lie_algebra.py:159-176). T (4,4)."""
    R, t = T[:3, :3], T[:3, 3]
    tr = R[0, 0] + R[1, 1] + R[2, 2]
    tr3 = tr - 3.0
    theta = torch.acos(0.5 * (tr - 1))
    mag = torch.where(tr3 < -1e-6, theta / (2.0 * torch.sin(theta)), 0.5 - tr3 / 12.0 + tr3 * tr3 / 60.0)
    w = mag * torch.stack((R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]))
    th = torch.clamp(torch.linalg.norm(w), min=1e-6)
    wn = w / th
    tan = torch.tan(0.5 * th)
    wxt = torch.linalg.cross(wn, t)
    Vinv_t = t - (0.5 * t) * wxt + (1.0 - th / (2.0 * tan)) * torch.linalg.cross(wn, wxt)
    return torch.cat((w, Vinv_t))


# ------------------------------------------------------------------------------------------ integer parts
def photometric_pairs(num_kf, kf_ts, recent_ts):
    """fwd (i,i+1), bwd (i+1,i); each one-way frame to its bracketing keyframes by timestamp
    (graph_pair_construction.py:5-17,97-133,155-182; radius mode off as in config/como.yml:40-41)."""
    ref = list(range(0, num_kf - 1)) + list(range(1, num_kf))
    tgt = list(range(1, num_kf)) + list(range(0, num_kf - 1))
    ow_kf, ow_id = [], []
    nr = len(recent_ts)
    if nr > 0:
        kf_ind = -1
        while recent_ts[0] > kf_ts[kf_ind + 1]:
            kf_ind += 1
            if kf_ind == num_kf - 1:
                break
        r = 0
        if kf_ind < num_kf - 1:
            while r < nr:
                if recent_ts[r] > kf_ts[kf_ind + 1]:
                    kf_ind += 1
                if kf_ind >= num_kf - 1:
                    break
                ow_kf += [kf_ind, kf_ind + 1]
                ow_id += [r, r]
                r += 1
        while r < nr:
            ow_kf.append(kf_ind)
            ow_id.append(r)
            r += 1
    return ref, tgt, ow_kf, ow_id


def batched_landmark_ids(corr_mask):
    """(K,L) bool -> (K,M) landmark ids per keyframe in increasing order (every row must hold M trues)."""
    K = corr_mask.shape[0]
    rows = [torch.nonzero(corr_mask[k])[:, 0] for k in range(K)]
    M = max(int(r.numel()) for r in rows)
    out = torch.full((K, M), -1, dtype=torch.long)
    for k, r in enumerate(rows):
        out[k, : r.numel()] = r
    return out


def subselect_pixels(img_and_grads, win):
    """argmax of |grad I| in each win x win cell, first max wins (max_pool2d semantics). -> (K,N,2) [row,col]."""
    K, c3, H, W = img_and_grads.shape
    c = c3 // 3
    gn = torch.sqrt((img_and_grads[:, c:2 * c] ** 2 + img_and_grads[:, 2 * c:] ** 2).sum(1))
    hc, wc = H // win, W // win
    cells = gn[:, : hc * win, : wc * win].reshape(K, hc, win, wc, win).permute(0, 1, 3, 2, 4).reshape(K, hc, wc, win * win)
    # first index of the maximum in row-major order inside the cell
    mx = cells.max(-1, keepdim=True).values
    first = torch.argmax((cells == mx).to(torch.int8), dim=-1)
    rr = torch.arange(hc).view(1, hc, 1) * win + first // win
    cc = torch.arange(wc).view(1, 1, wc) * win + first % win
    return torch.stack((rr, cc), -1).reshape(K, hc * wc, 2)


def bilinear3(img3, u, v):
    """img3 (3,H,W); sample at x=u, y=v (pixel centres at integers), zero padding. -> (3,N)"""
    _, H, W = img3.shape
    x0 = torch.floor(u)
    y0 = torch.floor(v)
    fx, fy = u - x0, v - y0
    x0, y0 = x0.long(), y0.long()

    def tap(xx, yy):
        ok = (xx >= 0) & (xx < W) & (yy >= 0) & (yy < H)
        val = img3[:, yy.clamp(0, H - 1), xx.clamp(0, W - 1)]
        return val * ok.to(val.dtype)

    return (tap(x0, y0) * ((1 - fx) * (1 - fy)) + tap(x0 + 1, y0) * (fx * (1 - fy))
            + tap(x0, y0 + 1) * ((1 - fx) * fy) + tap(x0 + 1, y0 + 1) * (fx * fy))


# ------------------------------------------------------------------------------------------ iterate
def iterate(s, cfg, rank=0, world=1, photo_only=False):
    """One BA GN iteration on a state dict `s` (keys = the reference Mapping attribute names).
    Mutates s like Mapping.iterate and returns a dict of intermediates (H, g, delta, errors, pairs ...).
    rank/world model the multi-GPU sharding: the robust scale uses ALL pairs of a batch, but only the pairs
    whose reference keyframe k has k % world == rank are accumulated; with photo_only the function returns
    after the photometric blocks (the quantity the ranks all-reduce)."""
    Kmat = s["intrinsics"][0]
    fx, fy, cx, cy = Kmat[0, 0], Kmat[1, 1], Kmat[0, 2], Kmat[1, 2]
    kf_poses = s["kf_poses"]
    nk = kf_poses.shape[0]
    nr = s["recent_poses"].shape[0] if s["recent_poses"].numel() > 0 else 0
    corr = s["correspondence_mask"]
    L = s["P_m"].shape[0]
    lm_ids = batched_landmark_ids(corr)  # (K,M)
    M = lm_ids.shape[1]
    med = s["median_depths"]

    # ---- scaffold (a13, a14)
    first_kf = torch.argmax(corr.int(), dim=0)
    first_full = torch.zeros_like(corr)
    first_full[first_kf, torch.arange(L)] = True
    first_b = torch.gather(first_full, 1, lm_ids)  # (K,M)
    R_wc, t_wc = kf_poses[:, :3, :3], kf_poses[:, :3, 3]
    pfo = s["pm_first_obs"]
    ray_fo = torch.stack(((pfo[..., 0] - cx) / fx, (pfo[..., 1] - cy) / fy, torch.ones_like(pfo[..., 0])), -1)
    init_Pw = ((med[:, None, None] * ray_fo) @ R_wc.transpose(1, 2)) + t_wc[:, None, :]
    init_Pm = init_Pw[first_b]  # (L,3) in (kf, slot) order -- indexed below as if by landmark id (reference quirk)
    Pwm = s["P_m"][lm_ids]
    rPwm = init_Pm[lm_ids]
    R_cw = R_wc.transpose(1, 2)
    t_cw = -(R_cw @ t_wc[..., None])[..., 0]
    Pc = Pwm @ R_cw.transpose(1, 2) + t_cw[:, None, :]
    z_mask = Pc[..., 2] < 0.1 * med[:, None]
    if z_mask.any():
        rPc = rPwm @ R_cw.transpose(1, 2) + t_cw[:, None, :]
        Pc = torch.where(z_mask[..., None], rPc, Pc)
    z = Pc[..., 2]
    logzm = torch.log(z)[..., None]  # (K,M,1)
    pm = torch.stack((fx * Pc[..., 0] / z + cx, fy * Pc[..., 1] / z + cy), -1)
    # d Pc / d Twc = [Pc^ | -I]   (== dPc_dTcw @ -Adj(Twc), sparse_map.py:46)
    dPc_dT = torch.cat((skew(Pc), -torch.eye(3, dtype=F64).expand(nk, M, 3, 3)), -1)  # (K,M,3,6)
    dz_dP = R_cw[:, 2, :]  # (K,3)   d z / d Pw, constant per keyframe
    dz_dT = dPc_dT[:, :, 2, :]  # (K,M,6)
    dpi = torch.zeros(nk, M, 2, 3, dtype=F64)
    dpi[..., 0, 0] = fx / z
    dpi[..., 0, 2] = -fx * Pc[..., 0] / z / z
    dpi[..., 1, 1] = fy / z
    dpi[..., 1, 2] = -fy * Pc[..., 1] / z / z
    dp_dP = dpi @ R_cw[:, None]  # (K,M,2,3)
    dp_dT = dpi @ dPc_dT  # (K,M,2,6)
    reinit = z_mask[first_b]
    s["P_m"] = s["P_m"].clone()
    s["P_m"][reinit] = init_Pm[reinit]
    u = 1.0 / z  # dlogz/dz
    dlogz_dT = u[..., None] * dz_dT  # (K,M,6)
    dlogz_dP = u[..., None] * dz_dP[:, None, :]  # (K,M,3)

    # ---- dense reference points (a15, a16)
    win = cfg["photo_construction"]["nonmax_suppression_window"]
    coords_n = subselect_pixels(s["kf_img_and_grads"], win)
    N = coords_n.shape[1]
    kidx = torch.arange(nk)[:, None].expand(nk, N)
    rows, cols = coords_n[..., 0], coords_n[..., 1]
    vals_n = s["kf_img_and_grads"][kidx, 0, rows, cols]  # (K,N)
    Kt = s["Knm_Kmminv"][kidx, rows, cols, :]  # (K,N,M)
    logz_n = (Kt @ logzm)[..., 0]
    z_n = torch.exp(logz_n)
    ray = torch.stack(((cols.to(F64) - cx) / fx, (rows.to(F64) - cy) / fy, torch.ones(nk, N, dtype=F64)), -1)
    Pc_n = z_n[..., None] * ray
    q_n = Kt @ dlogz_dT  # (K,N,6)
    RPc = Pc_n @ R_wc.transpose(1, 2)  # R_wc Pc
    Pw_n = RPc + t_wc[:, None, :]
    med_n = torch.stack([lower_median(Pc_n[k, :, 2]) for k in range(nk)])

    # ---- store_vars (a25)
    s["pm"], s["logzm"] = pm, logzm
    depth = torch.exp((s["Knm_Kmminv"] @ logzm[:, None, :, :])[..., 0])  # (K,H,W)
    s["depth_imgs"] = depth[:, None]
    s["median_depths"] = torch.stack([lower_median(depth[k].reshape(-1)) for k in range(nk)])
    med = s["median_depths"]

    # ---- system layout
    dim = 8 * (nk + nr) + 3 * L
    H = torch.zeros(dim, dim, dtype=F64)
    g = torch.zeros(dim, dtype=F64)
    kf_inds = torch.arange(8 * nk).view(nk, 8)
    rec_inds = 8 * nk + torch.arange(8 * nr).view(nr, 8)
    lm_start = 8 * (nk + nr)
    lm_inds = (3 * lm_ids[..., None] + torch.arange(3)).reshape(nk, 3 * M) + lm_start  # (K,3M)

    def add_block(ri, ci, blk):
        H[ri[:, None], ci[None, :]] += blk

    # ---- photometric factors (a17-a22)
    ref, tgt, ow_kf, ow_id = photometric_pairs(nk, s["kf_timestamps"], s["recent_timestamps"])
    all_ref = ref + ow_kf
    n_kf_pairs = len(ref)
    bs = cfg["photo_construction"]["pairwise_batch_size"]
    photo_err = 0.0
    pair_dbg = []
    for b1 in range(0, len(all_ref), bs):
        pend = []
        for pi in range(b1, min(b1 + bs, len(all_ref))):
            i = all_ref[pi]
            if pi < n_kf_pairs:
                j = tgt[pi]
                Twj, affj, imgj, indj = kf_poses[j], s["kf_aff_params"][j], s["kf_img_and_grads"][j], kf_inds[j]
            else:
                j = ow_id[pi - n_kf_pairs]
                Twj, affj, imgj, indj = (s["recent_poses"][j], s["recent_aff_params"][j],
                                          s["recent_img_and_grads"][j], rec_inds[j])
            affi = s["kf_aff_params"][i]
            Rj = Twj[:3, :3]
            Pcj = (Pw_n[i] - Twj[:3, 3]) @ Rj  # R^T (Pw - t)
            X, Y, Z = Pcj[:, 0], Pcj[:, 1], Pcj[:, 2]
            uu = fx * X / Z + cx
            vv = fy * Y / Z + cy
            Hh, Ww = imgj.shape[-2:]
            valid = (uu >= 1) & (uu < Ww - 1) & (vv >= 1) & (vv < Hh - 1) & (Z > 0)
            smp = bilinear3(imgj, uu, vv)
            It, gx, gy = smp[0], smp[1], smp[2]
            dI = torch.stack((gx * fx / Z, gy * fy / Z, -(gx * fx * X / Z + gy * fy * Y / Z) / Z), -1)  # (N,3)
            vsc = torch.exp(affj[0, 0] - affi[0, 0]) * vals_n[i]
            r = It - vsc + (affj[1, 0] - affi[1, 0])
            pend.append((i, indj, Rj, Pcj, valid, dI, vsc, r))
        allr = torch.cat([p[7][p[4]].abs() for p in pend])
        sigma = 1.4826 * lower_median(allr)
        for (i, indj, Rj, Pcj, valid, dI, vsc, r) in pend:
            if i % world != rank:
                continue
            wr = (r / sigma).abs()
            wgt = torch.where(wr < HUBER_K, torch.ones_like(wr), HUBER_K / wr) * valid.to(F64)
            sc = torch.sqrt(wgt) / sigma
            rs = r * sc
            photo_err += float((rs * rs).sum())
            dIs = dI * sc[:, None]
            daffi = torch.stack((vsc * sc, -sc), -1)  # (N,2)
            dIw = dIs @ Rj.T  # dI/dPw = dI/dPc R_cw,  R_cw = Rj^T  -> row-vector times Rj^T
            alpha = (dIw * RPc[i]).sum(-1)
            # ref pose: dI/dPw [-R Pc^ | R] + alpha q^T
            A6 = torch.cat((-(dIw @ R_wc[i])[:, None, :] @ skew(Pc_n[i]), (dIw @ R_wc[i])[:, None, :]), -1)[:, 0, :]
            Ji = torch.cat((A6 + alpha[:, None] * q_n[i], daffi), -1)  # (N,8)
            Jj = torch.cat(((dIs[:, None, :] @ skew(Pcj))[:, 0, :], -dIs, -daffi), -1)  # (N,8)
            Jz = alpha[:, None] * Kt[i] * u[i][None, :]  # (N,M)
            d3 = dz_dP[i]
            ii, li = kf_inds[i], lm_inds[i]
            g.index_add_(0, ii, -(Ji * rs[:, None]).sum(0))
            g.index_add_(0, indj, -(Jj * rs[:, None]).sum(0))
            gz = -(Jz * rs[:, None]).sum(0)
            g.index_add_(0, li, (gz[:, None] * d3[None, :]).reshape(-1))
            add_block(ii, ii, Ji.T @ Ji)
            add_block(indj, indj, Jj.T @ Jj)
            Hij = Ji.T @ Jj
            add_block(ii, indj, Hij)
            add_block(indj, ii, Hij.T)
            for (pinds, Jp) in ((ii, Ji), (indj, Jj)):
                Hpz = Jp.T @ Jz  # (8,M)
                HpP = (Hpz[:, :, None] * d3[None, None, :]).reshape(8, 3 * M)
                add_block(pinds, li, HpP)
                add_block(li, pinds, HpP.T)
            Hzz = Jz.T @ Jz
            HPP = (d3[None, :, None, None] * Hzz[:, None, :, None] * d3[None, None, None, :]).reshape(3 * M, 3 * M)
            add_block(li, li, HPP)
        pair_dbg.append(float(sigma))
    H_photo, g_photo = H.clone(), g.clone()
    if photo_only:
        return dict(H_photo=H_photo, g_photo=g_photo, photo_err=photo_err, sigmas=pair_dbg)

    # ---- priors (a23)
    logmed = torch.log(med)[:, None, None]
    pose_inds = kf_inds[:, :6]
    eye_m = torch.eye(M, dtype=F64)
    err_gp = err_ld = err_px = 0.0
    obs_ref = s["obs_ref_mask"]
    for k in range(nk):
        li, ti = lm_inds[k], pose_inds[k]
        # GP marginal-likelihood prior, sigma = 1
        Linv = torch.linalg.solve_triangular(s["L_mm"][k], eye_m, upper=False)
        rr = Linv @ (logzm[k] - logmed[k])  # (M,1)
        JP = (Linv[:, :, None] * dlogz_dP[k][None, :, :]).reshape(M, 3 * M)
        JT = Linv @ dlogz_dT[k]  # (M,6)
        g.index_add_(0, li, -(JP * rr).sum(0))
        g.index_add_(0, ti, -(JT * rr).sum(0))
        add_block(li, li, JP.T @ JP)
        add_block(ti, ti, JT.T @ JT)
        add_block(ti, li, JT.T @ JP)
        add_block(li, ti, JP.T @ JT)
        err_gp += float((rr * rr).sum())
        # first-observation log-depth prior (sigma_first = 1) and pixel prior (sigma_first = 1e-2)
        sc = obs_ref[k].to(F64)  # (M,)
        r1 = (logzm[k, :, 0] - logmed[k, 0, 0]) * sc
        JP1, JT1 = dlogz_dP[k], dlogz_dT[k]  # (M,3), (M,6)
        r2 = (pm[k] - s["pm_first_obs"][k]) * sc[:, None]  # (M,2)
        # reference quirk: the per-point scale lives in a float32 scratch tensor (pixel_prior.py:43)
        info2 = float(torch.tensor(1.0 / (1e-2 ** 2), dtype=torch.float32))
        for m in range(M):
            if sc[m] == 0:
                continue
            l3 = li[3 * m: 3 * m + 3]
            for (info, rv, JPm, JTm) in ((1.0, r1[m:m + 1], JP1[m:m + 1], JT1[m:m + 1]),
                                         (info2, r2[m], dp_dP[k, m], dp_dT[k, m])):
                g.index_add_(0, l3, -info * (JPm.T @ rv))
                g.index_add_(0, ti, -info * (JTm.T @ rv))
                add_block(l3, l3, info * JPm.T @ JPm)
                add_block(ti, ti, info * JTm.T @ JTm)
                add_block(ti, l3, info * JTm.T @ JPm)
                add_block(l3, ti, info * JPm.T @ JTm)
        err_ld += float((r1 * r1).sum())
        err_px += float(info2 * (r2 * r2).sum())
    # pose anchor on keyframe 0
    sp = cfg["sigmas"]["pose_prior"]
    xi = -se3_log_quirk(inv_se3(kf_poses[0]) @ s["pose_anchor"][0])
    info = 1.0 / sp * (1.0 / sp)
    # reference quirk: J = info_sqrt * eye(6) is float32, so J^T J adds float32(1e12) = 999999995904
    # to H while the gradient and the error use the double value (pose_prior_factors.py:12-17)
    info_H = float(torch.tensor(1.0 / sp, dtype=torch.float32) * torch.tensor(1.0 / sp, dtype=torch.float32))
    H[pose_inds[0][:, None], pose_inds[0][None, :]] += info_H * torch.eye(6, dtype=F64)
    g[pose_inds[0]] -= info * xi
    err_pose = float(info * (xi * xi).sum())
    # affine anchors on keyframe 0
    ss = cfg["sigmas"]["scale_prior"]
    info = 1.0 / ss * (1.0 / ss)
    err_aff = 0.0
    for c in range(2):
        idx = kf_inds[0, 6 + c]
        rv = s["kf_aff_params"][0, c, 0] - s["aff_anchor"][0, c, 0]
        g[idx] += -info * rv
        H[idx, idx] += info
        err_aff += float(info * rv * rv)
    err_scale = err_fixed = 0.0
    if s["window_full"]:
        fix = corr[0]
        rv = (s["P_m"][fix] - s["P_m_anchors"]).reshape(-1)
        idx = (lm_start + 3 * torch.nonzero(fix)[:, 0][:, None] + torch.arange(3)[None, :]).reshape(-1)
        g[idx] += -info * rv
        H[idx, idx] += info
        err_fixed = float(info * (rv * rv).sum())
    else:
        sm = cfg["sigmas"]["mean_depth_prior"]
        info_m = 1.0 / (sm ** 2)
        Kfull = s["Knm_Kmminv"][0].reshape(-1, M)
        nfull = Kfull.shape[0]
        rv = (Kfull @ logzm[0]).mean() - s["init_scale_anchor"].reshape(())
        dr = Kfull.sum(0) / nfull  # (M,)
        JP = (dr[:, None] * dlogz_dP[0]).reshape(1, 3 * M)
        JT = (dr[None, :] @ dlogz_dT[0])  # (1,6)
        li, ti = lm_inds[0], pose_inds[0]
        g.index_add_(0, li, -info_m * JP[0] * rv)
        g.index_add_(0, ti, -info_m * JT[0] * rv)
        add_block(li, li, info_m * JP.T @ JP)
        add_block(ti, ti, info_m * JT.T @ JT)
        add_block(ti, li, info_m * JT.T @ JP)
        add_block(li, ti, info_m * JP.T @ JT)
        err_scale = float(info_m * rv * rv)
    total_err = photo_err + err_gp + err_ld + err_px + err_pose + err_aff + err_scale + err_fixed

    # ---- solve + update (a24)
    Lc, _ = torch.linalg.cholesky_ex(H, upper=False, check_errors=False)
    delta = torch.cholesky_solve(g[:, None], Lc, upper=False)[:, 0]
    dk = delta[kf_inds]
    s["kf_poses"] = kf_poses @ se3_exp_batch(dk[:, :6])
    s["kf_aff_params"] = s["kf_aff_params"] + dk[:, 6:, None]
    if nr > 0:
        dr_ = delta[rec_inds]
        s["recent_poses"] = s["recent_poses"] @ se3_exp_batch(dr_[:, :6])
        s["recent_aff_params"] = s["recent_aff_params"] + dr_[:, 6:, None]
    s["P_m"] = s["P_m"] + delta[lm_start:].view(-1, 3)
    return dict(H=H, g=g, delta=delta, H_photo=H_photo, g_photo=g_photo, photo_err=photo_err, total_err=total_err,
                pairs=(ref, tgt, ow_kf, ow_id), coords_n=coords_n, Pwn=Pw_n, vals_n=vals_n, median_depths_n=med_n,
                sigmas=pair_dbg, errs=dict(gp=err_gp, ld=err_ld, px=err_px, pose=err_pose, aff=err_aff,
                                          scale=err_scale, fixed=err_fixed))


def state_from_golden(g, prefix="in_"):
    """Builds the state dict from a tests/golden/ba_*.npz file."""
    s = {}
    for k in g.files:
        if k.startswith(prefix):
            v = g[k]
            name = k[len(prefix):]
            if name in ("kf_timestamps", "recent_timestamps"):
                s[name] = [float(x) for x in v]
            elif name == "window_full":
                s[name] = bool(v)
            else:
                s[name] = torch.from_numpy(v)
    return s


def img_and_grads(rgb):
    """Mapping.get_img_and_grads, color "gray" (como/odom/Mapping.py:368-376): torchvision rgb_to_grayscale
    (0.2989 R + 0.587 G + 0.114 B) and the Scharr/32 gradients with reflect padding
    (como/utils/image_processing.py:8-44).  rgb (1,3,H,W) float64 -> (1,3,H,W) [I, gx, gy]."""
    img = (0.2989 * rgb[:, 0:1] + 0.587 * rgb[:, 1:2] + 0.114 * rgb[:, 2:3]).to(rgb.dtype)
    kx = torch.tensor([[-3.0, 0.0, 3.0], [-10.0, 0.0, 10.0], [-3.0, 0.0, 3.0]], dtype=rgb.dtype) * (1.0 / 32.0)
    ky = torch.tensor([[-3.0, -10.0, -3.0], [0.0, 0.0, 0.0], [3.0, 10.0, 3.0]], dtype=rgb.dtype) * (1.0 / 32.0)
    xp = torch.nn.functional.pad(img, (1, 1, 1, 1), mode="reflect")
    gx = torch.nn.functional.conv2d(xp, kx.view(1, 1, 3, 3))
    gy = torch.nn.functional.conv2d(xp, ky.view(1, 1, 3, 3))
    return torch.cat((img, gx, gy), dim=1)


def state_from_compact_golden(g, img_fn=None, slab_fn=None):
    """State dict from a compact golden (gen_golden.gen_ba_compact): the image stacks are regenerated from the stored
    texture + frame offsets (`img_fn(rgb) -> (1,3,H,W)`, default: the restatement above) and the dense predictor slab
    from the stored covariance images, anchors and the reference's own K_mm^-1 (`slab_fn(cov (K,4,H,W) f64,
    coords_m (K,M,2), Kmm_inv (K,M,M), scale) -> (K,H,W,M)`, default: oracle/depthcov_oracle.py), after which the rows
    at the sampled pixels are replaced by the reference's exact rows.  Returns (state, report) -- report holds the
    agreement of the regenerated parts with the stored checksums."""
    from oracle import depthcov_oracle as DO

    s = state_from_golden(g)
    W = int(g["W"])
    tex = torch.from_numpy(g["tex"])
    img_fn = img_fn or img_and_grads
    frames = lambda xs: torch.cat([img_fn(tex[..., int(x0):int(x0) + W].clone()).to("cpu") for x0 in xs], 0)
    s["kf_img_and_grads"] = frames(g["kf_x0"])
    s["recent_img_and_grads"] = frames(g["recent_x0"])
    cov = torch.from_numpy(g["in_cov_params_img_f32"]).double()
    s["cov_params_img"] = cov
    s.pop("cov_params_img_f32", None)
    if slab_fn is None:
        def slab_fn(cov, cm, Kinv, scale):
            K, _, H, Wc = cov.shape
            cmn = DO._normalize(cm.double(), (H, Wc))
            E_m = DO._interp_cov(cov, cmn)
            rr, cc = torch.meshgrid(torch.arange(H), torch.arange(Wc), indexing="ij")
            cn = DO._normalize(torch.stack((rr.reshape(-1), cc.reshape(-1)), 1)[None].double(), (H, Wc))
            out = []
            for k in range(K):
                E_n = DO._interp_cov(cov[k:k + 1], cn)
                out.append((DO.cov_python(cn, E_n, cmn[k:k + 1], E_m[k:k + 1], scale) @ Kinv[k:k + 1]).reshape(1, H, Wc, -1))
            return torch.cat(out, 0)
    # the predictor was built at the first-observation pixels; state stores them as (x, y), prep_predictor takes (row, col)
    coords_m = s["pm_first_obs"].flip(-1).contiguous()
    slab = slab_fn(cov, coords_m, s["Kmm_inv"], float(g["gp_scale"])).to("cpu").clone()
    rows = torch.from_numpy(g["in_Knm_rows"])
    cn = torch.from_numpy(g["coords_n"])
    kidx = torch.arange(slab.shape[0])[:, None]
    mine = slab[kidx, cn[..., 0], cn[..., 1]]
    rep = dict(
        img0=float((s["kf_img_and_grads"][0] - torch.from_numpy(g["chk_kf_img_and_grads_0"])).abs().max()),
        img_sum=float((s["kf_img_and_grads"].sum((2, 3)) - torch.from_numpy(g["chk_kf_img_and_grads_sum"])).abs().max()),
        rec_sum=float((s["recent_img_and_grads"].sum((2, 3)) - torch.from_numpy(g["chk_recent_img_and_grads_sum"])).abs().max()),
        rows=float((mine - rows).abs().max() / rows.abs().max()),
        colsum=float((slab.sum((1, 2)) - torch.from_numpy(g["chk_Knm_colsum"])).abs().max()
                     / torch.from_numpy(g["chk_Knm_colsum"]).abs().max()))
    slab[kidx, cn[..., 0], cn[..., 1]] = rows
    s["Knm_Kmminv"] = slab
    s.pop("Knm_rows", None)
    s["depth_imgs"] = torch.zeros(slab.shape[0], 1, slab.shape[1], slab.shape[2], dtype=torch.float64)
    return s, rep


def cfg_from_golden(g):
    return dict(
        photo_construction=dict(nonmax_suppression_window=4, pairwise_batch_size=int(g["photo_cfg_batch"])),
        sigmas=dict(mean_depth_prior=float(g["sigma_mean_depth_prior"]), scale_prior=float(g["sigma_scale_prior"]),
                    pose_prior=float(g["sigma_pose_prior"])),
    )
