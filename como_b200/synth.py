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
