"""Seeded synthetic scenes shared by tests, bench.py and the golden generator.

Pure torch, CPU or CUDA; no dataset, no reference import.
"""
import math

import torch


def make_rgb(H, W, seed=0, cell=16, noise=0.05, dtype=torch.float32, device="cpu", extra_w=0):
    """Band-limited random texture: uniform noise on a coarse grid, bicubic-upsampled, + fine noise."""
    g = torch.Generator().manual_seed(seed)
    Wt = W + extra_w
    low = torch.rand(1, 3, max(H // cell, 2), max(Wt // cell, 2), generator=g)
    img = torch.nn.functional.interpolate(low, size=(H, Wt), mode="bicubic", align_corners=False)
    img = img + noise * torch.rand(1, 3, H, Wt, generator=g)
    return img.clamp(0.0, 1.0).to(dtype=dtype, device=device)


def make_depth(H, W, dtype=torch.float32, device="cpu"):
    y = torch.linspace(0, 1, H).view(H, 1)
    x = torch.linspace(0, 1, W).view(1, W)
    d = 2.0 + 0.5 * torch.sin(3 * x) * torch.cos(2 * y)
    return d.view(1, 1, H, W).to(dtype=dtype, device=device)


def make_intrinsics(H, W, dtype=torch.float32, device="cpu"):
    f = 525.0 * W / 640.0
    return torch.tensor([[f, 0, W / 2.0], [0, f, H / 2.0], [0, 0, 1.0]], dtype=dtype, device=device)


def se3_exp_wv(xi):
    """Host helper: SE(3) exponential with COMO ordering xi = [omega(3), v(3)] -> (4,4) float64 tensor."""
    w = [float(v) for v in xi[:3]]
    v = [float(t) for t in xi[3:]]
    th2 = w[0] ** 2 + w[1] ** 2 + w[2] ** 2
    th = math.sqrt(th2)
    if th2 < 1e-12:
        A, B, C = 1 - th2 / 6, 0.5 - th2 / 24, 1.0 / 6 - th2 / 120
    else:
        A, B, C = math.sin(th) / th, (1 - math.cos(th)) / th2, (th - math.sin(th)) / (th2 * th)
    Wm = torch.tensor([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]], dtype=torch.float64)
    WW = Wm @ Wm
    I = torch.eye(3, dtype=torch.float64)
    T = torch.eye(4, dtype=torch.float64)
    T[:3, :3] = I + A * Wm + B * WW
    T[:3, 3] = (I + B * Wm + C * WW) @ torch.tensor(v, dtype=torch.float64)
    return T


def make_cov_image(H, W, seed=0, dtype=torch.float32, device="cpu"):
    """Smooth SPD 2x2 covariance-parameter image (1,4,H,W) [E00,E01,E10,E11], like the DepthCov UNet head's output
    after gaussian_kernel.kernel_params_to_covariance (x, z > 0, |rho| < 0.99)."""
    g = torch.Generator().manual_seed(seed)
    low = torch.rand(1, 3, max(H // 12, 2), max(W // 12, 2), generator=g)
    f = torch.nn.functional.interpolate(low, size=(H, W), mode="bicubic", align_corners=False).clamp(0.02, 0.98)
    x = 2e-3 * torch.exp(3.0 * f[:, 0])
    z = 2e-3 * torch.exp(3.0 * f[:, 1])
    rho = 0.9 * (2 * f[:, 2] - 1)
    off = torch.sqrt(x * z - 1e-8) * rho
    return torch.stack((x, off, off, z), dim=1).to(dtype=dtype, device=device)


def make_cov_image_wide(H, W, seed=0, dtype=torch.float64, device="cpu"):
    """Variant with long, slowly varying length scales (std 0.45-0.65 in normalised coordinates): the predictor
    rows sum to ~1 (0.9..1.06), i.e. the GP interpolates between anchors instead of reverting to its zero mean --
    what the trained DepthCov network produces on smooth surfaces."""
    g = torch.Generator().manual_seed(seed)
    low = torch.rand(1, 3, max(H // 24, 2), max(W // 24, 2), generator=g)
    f = torch.nn.functional.interpolate(low, size=(H, W), mode="bicubic", align_corners=False).clamp(0.02, 0.98)
    x = 2e-1 * torch.exp(0.7 * f[:, 0])
    z = 2e-1 * torch.exp(0.7 * f[:, 1])
    rho = 0.5 * (2 * f[:, 2] - 1)
    off = torch.sqrt(x * z - 1e-8) * rho
    return torch.stack((x, off, off, z), dim=1).to(dtype=dtype, device=device)


TRACK_PERTURB = (0.01, -0.008, 0.005, 0.02, -0.01, 0.015)  # [omega, v], SURVEY 8d


# ---------------------------------------------------------------------------------------------
# Synthetic tracking case in the reference's operand layout (what Tracking.update_kf_reference
# hands to photo_tracking_pyr), built with plain torch ops.  Input generation only.
# ---------------------------------------------------------------------------------------------
def _gray(rgb):
    return (0.2989 * rgb[:, 0:1] + 0.587 * rgb[:, 1:2] + 0.114 * rgb[:, 2:3]).to(rgb.dtype)


def _blur_down(x):
    k = torch.tensor([[1.0, 2.0, 1.0], [2.0, 4.0, 2.0], [1.0, 2.0, 1.0]], dtype=x.dtype, device=x.device) / 16.0
    xp = torch.nn.functional.pad(x, (1, 1, 1, 1), mode="reflect")
    return torch.nn.functional.conv2d(xp, k.view(1, 1, 3, 3))[:, :, 0::2, 0::2]


def _scharr(x):
    kx = torch.tensor([[-3.0, 0.0, 3.0], [-10.0, 0.0, 10.0], [-3.0, 0.0, 3.0]], dtype=x.dtype, device=x.device) / 32.0
    xp = torch.nn.functional.pad(x, (1, 1, 1, 1), mode="reflect")
    gx = torch.nn.functional.conv2d(xp, kx.view(1, 1, 3, 3))
    gy = torch.nn.functional.conv2d(xp, kx.t().contiguous().view(1, 1, 3, 3))
    return gx, gy


def image_pyramid(gray, num_levels):
    """Coarsest first (the reference's convention)."""
    pyr = [gray]
    for _ in range(num_levels - 1):
        pyr.insert(0, _blur_down(pyr[0]))
    return pyr


def intrinsics_pyramid(K, num_levels):
    """The reference's resize_intrinsics quirk: the scale is ADDED to the principal point."""
    out = []
    for i in range(num_levels):
        s = 2.0 ** (-i)
        Tm = torch.tensor([[s, 0, s], [0, s, s], [0, 0, 1.0]], dtype=K.dtype, device=K.device)
        out.insert(0, Tm @ K)
    return out


def make_tracking_case(H, W, num_levels, seed=0, cell=16, device="cpu", noise2=0.03):
    """Returns dict with per-level lists vals (1,N,1), P (1,N,3), dI_dT (1,N,1,8), mask (1,N) bool,
    K (3,3), img2 (1,1,h,w) plus T_init (1,4,4), aff_init (1,2,1)."""
    rgb = make_rgb(H, W, seed=seed, cell=cell)
    depth = make_depth(H, W)
    K = make_intrinsics(H, W)
    g = torch.Generator().manual_seed(1000 + seed)
    rgb2 = (rgb * 1.05 + 0.02 + noise2 * (torch.rand(rgb.shape, generator=g) - 0.5)).clamp(0, 1)
    rgb, depth, K, rgb2 = rgb.to(device), depth.to(device), K.to(device), rgb2.to(device)
    pyr1 = image_pyramid(_gray(rgb), num_levels)
    pyr2 = image_pyramid(_gray(rgb2), num_levels)
    Kp = intrinsics_pyramid(K, num_levels)
    dpyr = [depth]
    for _ in range(num_levels - 1):
        dpyr.insert(0, dpyr[0][:, :, 0::2, 0::2])
    out = dict(vals=[], P=[], dI_dT=[], mask=[], K=Kp, img=pyr2, rgb=rgb, rgb2=rgb2, depth=depth, K0=K)
    for l in range(num_levels):
        img = pyr1[l]
        h, w = img.shape[-2:]
        gx, gy = _scharr(img)
        ys, xs = torch.meshgrid(torch.arange(h, device=device), torch.arange(w, device=device), indexing="ij")
        xs = xs.reshape(-1).float()
        ys = ys.reshape(-1).float()
        z = dpyr[l].reshape(-1)
        Kl = Kp[l]
        X = (xs - Kl[0, 2]) / Kl[0, 0] * z
        Y = (ys - Kl[1, 2]) / Kl[1, 1] * z
        P = torch.stack((X, Y, z), dim=1)
        fx, fy = Kl[0, 0], Kl[1, 1]
        gxv, gyv = gx.reshape(-1), gy.reshape(-1)
        a = gxv * fx / z
        b = gyv * fy / z
        c = -(gxv * fx * X / z + gyv * fy * Y / z) / z
        J = torch.stack((b * (-z) + c * Y, a * z - c * X, -a * Y + b * X, a, b, c, img.reshape(-1),
                         torch.ones_like(z)), dim=1)
        px = fx * X / z + Kl[0, 2]
        py = fy * Y / z + Kl[1, 2]
        mask = (px >= -50) & (px <= w - 1 + 50) & (py >= -50) & (py <= h - 1 + 50) & (z > 1e-4)
        out["vals"].append(img.reshape(1, -1, 1).contiguous())
        out["P"].append(P.reshape(1, -1, 3).contiguous())
        out["dI_dT"].append(J.reshape(1, -1, 1, 8).contiguous())
        out["mask"].append(mask.reshape(1, -1))
    out["T_init"] = se3_exp_wv(TRACK_PERTURB).float()[None].to(device)
    out["aff_init"] = torch.zeros(1, 2, 1, device=device)
    return out


# ---------------------------------------------------------------------------------------------
# Synthetic BA window: fronto-parallel textured plane at Z = 2, camera translating +x by
# `step` px per keyframe; anchors from the device sampler, ~`ndrop` anchors replaced per keyframe so the
# landmark count grows like the real system (L ~ M + ndrop (K-1)); predictors from the fused K-matrix
# kernel.  Needs CUDA.  Returns a mapping_core.WindowState with the reference Mapping attribute names.
# ---------------------------------------------------------------------------------------------
def make_ba_window(K, R, H, W, M=64, device="cuda", seed=0, step=6.0, ndrop=20, pose_noise=2e-3, window_full=False,
                   sampler=None, predictor=None, img_noise=0.04):
    """`sampler(cov, n, curr_coords_or_None) -> (1,k,2) coords` and `predictor(cov, coords) -> (Kinv, L, slab)` default
    to the CUDA product implementations; tests inject CPU oracles to build the same scene without a GPU."""
    from como_b200.state import WindowState

    if sampler is None:
        from como_b200.depth_cov.core.samplers import sample_sparse_coords as _ssc

        def sampler(cov, n, curr):
            c, _ = _ssc(cov, n, "greedy_conditional_entropy", max_stdev_thresh=1e-2, border=3, dist_thresh=0.1,
                        signal_var=torch.tensor(1.0), fixed_var=0.0, curr_coords=curr)
            return c
    if predictor is None:
        from como_b200.depth_cov.core.predictor import prep_predictor as _pp

        def predictor(cov, coords):
            return _pp(cov, coords, 1.0)

    dev = torch.device(device)
    f64 = torch.float64
    g = torch.Generator().manual_seed(seed)
    Z = 2.0
    fx = 525.0 * W / 640.0
    Kmat = torch.tensor([[fx, 0, W / 2.0], [0, fx, H / 2.0], [0, 0, 1.0]], dtype=f64)
    tex = make_rgb(H, W, seed=seed, cell=16, extra_w=int(step * (K + 2)) + 8, dtype=f64, device=dev)

    def frame(k):
        x0 = int(k * step)
        rgb = tex[..., x0:x0 + W]
        # independent sensor noise per frame: keeps the robust scale sigma = 1.4826 med|r| at a realistic level
        # (noise-free copies of one texture let sigma -> 0 and the photometric weights explode)
        rgb = (rgb + img_noise * (torch.rand(rgb.shape, generator=g, dtype=f64) - 0.5).to(dev)).clamp(0, 1)
        gray = _gray(rgb)
        gx, gy = _scharr(gray)
        T = torch.eye(4, dtype=f64)
        T[0, 3] = k * step * Z / fx
        return torch.cat((gray, gx, gy), dim=1), T

    def noisy(T):
        T = T.clone()
        T[:3, 3] += pose_noise * (torch.rand(3, generator=g, dtype=f64) - 0.5)
        return T

    imgs, poses, covs, coords_all, Ls, Kinvs, slabs = [], [], [], [], [], [], []
    lm_of_slot = []  # per keyframe: landmark id of each slot
    first_dim = []
    nlm = 0
    P_list = []
    border = 3
    for k in range(K):
        img, T = frame(k)
        cov = make_cov_image_wide(H, W, seed=1000 + seed * 100 + k, dtype=f64, device=dev)
        if k == 0:
            c = sampler(cov, M, None)
            coords = c[0].to(f64)
            lms = list(range(M))
            nlm = M
            new_dim = M
        else:
            prev = coords_all[-1].clone()
            prev[:, 1] -= step  # points move -x when the camera moves +x (plane at constant depth)
            inside = (prev[:, 1] >= border + 1) & (prev[:, 1] <= W - border - 2)
            order = torch.randperm(M, generator=g)[:ndrop].tolist()
            keep = [m for m in range(M) if bool(inside[m]) and m not in order]
            tracked = prev[keep]
            c = sampler(cov, M, tracked[None].float())
            new = c[0].to(f64)
            coords = torch.cat((tracked, new), 0)[:M]
            new_dim = coords.shape[0] - len(keep)
            lms = [lm_of_slot[-1][m] for m in keep] + list(range(nlm, nlm + new_dim))
            nlm += new_dim
        assert coords.shape[0] == M, "sampler returned too few anchors for the synthetic window"
        Kinv, Lm, slab = predictor(cov, coords[None])
        zs = Z * (1.0 + 0.01 * (torch.rand(M, generator=g, dtype=f64) - 0.5)).to(dev)
        # world points of the NEW landmarks of this keyframe (GT pose, perturbed depth)
        pix = torch.stack((coords[:, 1], coords[:, 0]), -1)  # (x, y)
        Pc = torch.stack(((pix[:, 0] - Kmat[0, 2]) / fx * zs, (pix[:, 1] - Kmat[1, 2]) / fx * zs, zs), -1)
        Pw = Pc + T[:3, 3].to(dev)
        P_list.append(Pw[M - new_dim:])
        imgs.append(img)
        poses.append(noisy(T) if k > 0 else T)
        covs.append(cov)
        coords_all.append(coords)
        lm_of_slot.append(lms)
        first_dim.append(new_dim)
        Ls.append(Lm)
        Kinvs.append(Kinv)
        slabs.append(slab)
    L = nlm
    corr = torch.zeros(K, L, dtype=torch.bool)
    for k in range(K):
        corr[k, lm_of_slot[k]] = True
    obs_ref = torch.zeros(K, M, dtype=torch.bool)
    for k in range(K):
        obs_ref[k, M - first_dim[k]:] = True
    rec_imgs, rec_poses, rec_ts = [], [], []
    for j in range(R):
        kk = (j % (K - 1)) + 0.5
        img, T = frame(kk)
        rec_imgs.append(img)
        rec_poses.append(noisy(T))
        rec_ts.append(1.0 + kk + 0.001 * j)
    order = sorted(range(R), key=lambda i: rec_ts[i])
    pm = torch.stack([torch.stack((c[:, 1], c[:, 0]), -1) for c in coords_all]).to(dev)
    s = WindowState(
        intrinsics=Kmat[None].to(dev),
        kf_timestamps=[1.0 + k for k in range(K)],
        recent_timestamps=[rec_ts[i] for i in order],
        kf_poses=torch.stack(poses).to(dev).contiguous(),
        kf_aff_params=(0.01 * (torch.rand(K, 2, 1, generator=g, dtype=f64) - 0.5)).to(dev).contiguous(),
        recent_poses=(torch.stack([rec_poses[i] for i in order]).to(dev).contiguous() if R else torch.empty(0, dtype=f64, device=dev)),
        recent_aff_params=torch.zeros(R, 2, 1, dtype=f64, device=dev),
        kf_img_and_grads=torch.cat(imgs, 0).contiguous(),
        recent_img_and_grads=(torch.cat([rec_imgs[i] for i in order], 0).contiguous() if R else torch.empty(0, dtype=f64, device=dev)),
        cov_params_img=torch.cat(covs, 0),
        pm_first_obs=pm.contiguous(), pm=pm.clone(),
        logzm=torch.full((K, M, 1), math.log(Z), dtype=f64, device=dev),
        L_mm=torch.cat(Ls, 0).contiguous(), Kmm_inv=torch.cat(Kinvs, 0).contiguous(),
        Knm_Kmminv=torch.cat(slabs, 0).contiguous(),
        correspondence_mask=corr.to(dev), P_m=torch.cat(P_list, 0).contiguous(), obs_ref_mask=obs_ref.to(dev),
        pose_anchor=torch.stack(poses[:1]).to(dev), aff_anchor=torch.zeros(1, 2, 1, dtype=f64, device=dev),
        median_depths=torch.full((K,), Z, dtype=f64, device=dev), depth_imgs=None,
        window_full=bool(window_full), init_scale_anchor=torch.tensor(math.log(Z), dtype=f64, device=dev).view(1, 1, 1),
        iter=0, converged=False,
    )
    s.kf_aff_params[0] = 0.0
    if window_full:
        s.P_m_anchors = s.P_m[s.correspondence_mask[0]].clone()
    return s


def ba_cfg(batch=128):
    """The reference's mapping config values used by iterate (config/como.yml:26-51)."""
    return dict(photo_construction=dict(nonmax_suppression_window=4, pairwise_batch_size=batch, radius_thresh=0.0,
                                        degrees_thresh=0.0),
                sigmas=dict(photo=1e-1, mean_depth_prior=1e-2, scale_prior=1e-4, pose_prior=1e-6))
