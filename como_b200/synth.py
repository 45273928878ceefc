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


TRACK_PERTURB = (0.01, -0.008, 0.005, 0.02, -0.01, 0.015)  # [omega, v], SURVEY 8d
