"""ORACLE (test infrastructure, never on the product path): DepthCov GP pieces on the CPU.

* `ref_backends()` loads oracle/_ref/como_backends.so -- the reference's OWN C++ sources
  (como/backend/src/{cov.cpp,cov_cpu.cpp,depth_cov_backends.cpp}) compiled where they lie by
  oracle/ref_harness.build_ref_backends(); it travels to the GPU box as a built file.
* `cross_covariance_np` restates cov_cpu.cpp:17-64 in numpy (float32 expression types kept).
* `sample_sparse_coords` restates the greedy conditional-entropy sampler
  (como/depth_cov/core/samplers.py:36-326) in torch CPU on top of either backend.
* `prep_predictor` restates Mapping.prep_predictor (como/odom/Mapping.py:430-468) with the Python
  covariance formula (como/depth_cov/core/kernels.py:22-89, covariance.py:10-39).
Pinned by tests/golden/depthcov.npz and ba_*.npz (tests/test_oracle_depthcov.py).
"""
import importlib.util
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
_ref = None


def ref_backends():
    global _ref
    if _ref is None:
        path = os.path.join(HERE, "_ref", "como_backends.so")
        if not os.path.exists(path):
            return None
        spec = importlib.util.spec_from_file_location("como_backends", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _ref = mod
    return _ref


def cross_covariance_np(x1, E1, x2, E2, scale):
    f = np.float32
    x1, E1, x2, E2 = (np.asarray(a, dtype=f) for a in (x1, E1, x2, E2))
    dx = x1[:, :, None, 0] - x2[:, None, :, 0]
    dy = x1[:, :, None, 1] - x2[:, None, :, 1]
    E00 = E1[:, :, None, 0, 0] + E2[:, None, :, 0, 0]
    E01 = E1[:, :, None, 0, 1] + E2[:, None, :, 0, 1]
    E11 = E1[:, :, None, 1, 1] + E2[:, None, :, 1, 1]
    det = (E00 * E11 - E01 * E01).astype(f)
    det_inv = (1.0 / det.astype(np.float64)).astype(f)
    Q = ((E11 * dx * dx) - f(2) * (E01 * dx * dy) + (E00 * dy * dy)).astype(f)
    Q = (Q.astype(np.float64) * (0.5 * det_inv.astype(np.float64))).astype(f)
    d1 = (E1[..., 0, 0] * E1[..., 1, 1] - E1[..., 0, 1] * E1[..., 1, 0]).astype(f)[:, :, None]
    d2 = (E2[..., 0, 0] * E2[..., 1, 1] - E2[..., 0, 1] * E2[..., 1, 0]).astype(f)[:, None, :]
    pw = np.sqrt(np.sqrt((d1 * d2).astype(f).astype(np.float64))).astype(f)
    ssq = np.sqrt(det_inv.astype(np.float64) + 1e-8).astype(f)
    C = (2.0 * pw.astype(np.float64) * ssq.astype(np.float64)).astype(f)
    sq = np.sqrt(Q.astype(np.float64) + 1e-8).astype(f)
    tmp = (1.73205080757 * sq.astype(np.float64)).astype(f)
    mat = ((f(1) + tmp) * np.exp(-tmp.astype(np.float64)).astype(f)).astype(f)
    return (f(scale) * C * mat).astype(f)


def _cc(x1, E1, x2, E2, scale):
    rb = ref_backends()
    if rb is not None:
        return rb.cross_covariance(x1.contiguous(), E1.contiguous(), x2.contiguous(), E2.contiguous(), float(scale))
    return torch.from_numpy(cross_covariance_np(x1.numpy(), E1.numpy(), x2.numpy(), E2.numpy(), scale))


def _normalize(x_pixel, dims):
    A = 1.0 / torch.as_tensor(dims, dtype=x_pixel.dtype)
    return 2 * A * x_pixel + A - 1


def _interp_cov(cov, x_norm):
    grid = torch.stack((x_norm[..., 1], x_norm[..., 0]), -1).unsqueeze(1)
    s = torch.nn.functional.grid_sample(cov, grid, mode="bilinear", padding_mode="border", align_corners=False)
    return torch.permute(s.squeeze(2), (0, 2, 1)).reshape(cov.shape[0], -1, 2, 2)


def sample_sparse_coords(cov, num_samples, max_stdev_thresh=-1e8, border=0, terminate_early=False, dist_thresh=0.0,
                         signal_var=1.0, fixed_var=None, curr_coords=None, coords_domain=None):
    """Greedy conditional-entropy anchor selection; returns (coords (1,k,2) long/float, domain_inds (1,k)).
    coords_domain (1,d,2): explicit (fractional) candidate set instead of the bordered pixel grid
    (samplers.py:68-74)."""
    cov = cov.float()
    b, _, h, w = cov.shape
    assert b == 1
    if coords_domain is None:
        rr, cc = torch.meshgrid(torch.arange(border, h - border), torch.arange(border, w - border), indexing="ij")
        dom = torch.stack((rr.reshape(-1), cc.reshape(-1)), 1)[None]
        dn = _normalize(dom, (h, w)).float()
        Ed = torch.permute(cov[:, :, dom[0, :, 0], dom[0, :, 1]], (0, 2, 1)).reshape(1, -1, 2, 2).contiguous()
    else:
        dom = coords_domain
        dn = _normalize(dom, (h, w)).float()
        Ed = _interp_cov(cov, dn).contiguous()
    d = dn.shape[1]
    n = min(num_samples, d)
    sv = float(signal_var)
    idx = torch.full((1, n), -1, dtype=torch.long)
    xy = torch.zeros(1, n, 2)
    En = torch.zeros(1, n, 2, 2)
    L = torch.eye(n)[None].clone()
    obs = torch.zeros(1, n, d)
    if curr_coords is not None and curr_coords.shape[1] > 0:
        cn = _normalize(curr_coords.float(), (h, w)).float()
        m = cn.shape[1]
        xy[:, :m] = cn
        En[:, :m] = _interp_cov(cov, cn)
    else:
        areas = Ed[..., 0, 0] * Ed[..., 1, 1] - Ed[..., 0, 1] * Ed[..., 1, 0]
        best = int(torch.argmax(areas.view(-1)))
        idx[0, 0] = best
        xy[0, 0] = dn[0, best]
        En[0, 0] = Ed[0, best]
        m = 1
    Knn = _cc(xy[:, :m], En[:, :m], xy[:, :m].clone(), En[:, :m].clone(), sv)
    if fixed_var is not None:
        Knn = Knn + torch.diag_embed(float(fixed_var) * torch.ones(1, m))
    L[:, :m, :m] = torch.linalg.cholesky(Knn)
    Kmd = _cc(xy[:, :m], En[:, :m], dn, Ed, sv)
    obs[:, :m] = torch.linalg.solve_triangular(L[:, :m, :m], Kmd, upper=False)
    var = sv - torch.sum(obs[:, :m] * obs[:, :m], dim=1)
    th2 = dist_thresh * dist_thresh

    def pick(i):
        sd = torch.sqrt(var)
        sd[sd.isnan()] = 0.0
        sd = sd + 1e-10
        d2 = torch.sum(torch.square(xy[:, :i, None, :] - dn[:, None, :, :]), dim=-1)
        ok = (d2 > th2).all(dim=1)
        j = int(torch.argmax(sd * ok, dim=1))
        return j, float(sd[0, j])

    j, smax = pick(m)
    cnt = n
    rb = ref_backends()
    for i in range(m, n):
        if terminate_early and smax < max_stdev_thresh:
            cnt = i
            break
        idx[0, i] = j
        xy[0, i] = dn[0, j]
        En[0, i] = Ed[0, j]
        k_ni = _cc(xy[:, :i], En[:, :i], xy[:, i:i + 1], En[:, i:i + 1], sv)
        k_id = _cc(xy[:, i:i + 1], En[:, i:i + 1], dn, Ed, sv)
        k_ii = sv + (float(fixed_var) if fixed_var is not None else 0.0)
        if rb is not None:
            rb.get_new_chol_obs_info(L, obs, var, k_ni.contiguous(), k_id.contiguous(), k_ii, i)
        else:
            l_ni = torch.linalg.solve_triangular(L[:, :i, :i], k_ni, upper=False)
            l_ii = torch.sqrt(k_ii - torch.sum(torch.square(l_ni), dim=1, keepdim=True))
            new = (k_id - torch.sum(l_ni * obs[:, :i], dim=1, keepdim=True)) / l_ii
            L[:, i:i + 1, :i] = l_ni.transpose(1, 2)
            L[:, i, i] = l_ii[:, 0, 0]
            obs[:, i:i + 1] = new
            var -= (new * new).squeeze(1)
        j, smax = pick(i + 1)
    ids = idx[:, :cnt]
    ids = ids[:, ids[0] >= 0]
    return dom[0][ids[0]][None], ids


def cov_python(x1, E1, x2, E2, scale):
    """covariance.py:22-39 / kernels.py:22-89 in float64 with the float32-rounded coordinate difference."""
    dfl = (x1[:, :, None, :] - x2[:, None, :, :]).float()
    diff = dfl.double()
    sq = torch.square(dfl).double()  # torch.square of the float32 difference (kernels.py:27-29,37-39)
    s00 = E1[:, :, None, 0, 0] + E2[:, None, :, 0, 0]
    s01 = E1[:, :, None, 0, 1] + E2[:, None, :, 0, 1]
    s11 = E1[:, :, None, 1, 1] + E2[:, None, :, 1, 1]
    Q = s11 * sq[..., 0]
    Q = Q + (-2 * s01 * diff[..., 0] * diff[..., 1])
    Q = Q + s00 * sq[..., 1]
    det = s00 * s11 - s01 ** 2
    Q = Q / det * 0.5
    r1 = (E1[..., 0, 0] * E1[..., 1, 1] - E1[..., 0, 1] * E1[..., 1, 0]) ** 0.25
    r2 = (E2[..., 0, 0] * E2[..., 1, 1] - E2[..., 0, 1] * E2[..., 1, 0]) ** 0.25
    C = 2.0 * r1[:, :, None] * r2[:, None, :] / torch.sqrt(det + 1e-8)
    t = np.sqrt(3) * torch.sqrt(Q + 1e-8)
    return (1 + t) * torch.exp(-t) * C * scale


def prep_predictor(cov_img, coords_m, scale, jitter=1e-6):
    """cov_img (B,4,H,W) float64, coords_m (B,M,2) [row,col] -> (Kmm_inv, L_mm, Knm_Kmminv (B,H,W,M))."""
    B, _, H, W = cov_img.shape
    M = coords_m.shape[1]
    cm = _normalize(coords_m.double(), (H, W))
    E_m = _interp_cov(cov_img, cm)
    rr, cc = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    cn = _normalize(torch.stack((rr.reshape(-1), cc.reshape(-1)), 1)[None].repeat(B, 1, 1).double(), (H, W))
    E_n = _interp_cov(cov_img, cn)
    K_mm = cov_python(cm, E_m, cm, E_m, scale) + jitter * torch.eye(M, dtype=torch.float64)
    L_mm, _ = torch.linalg.cholesky_ex(K_mm)
    Kinv = torch.cholesky_solve(torch.eye(M, dtype=torch.float64).expand(B, M, M), L_mm)
    K_nm = cov_python(cn, E_n, cm, E_m, scale)
    return Kinv, L_mm, (K_nm @ Kinv).reshape(B, H, W, M)
