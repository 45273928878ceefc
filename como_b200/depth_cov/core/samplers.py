"""Anchor (depth inducing point) selection: drop-in for `sample_sparse_coords`
(como/depth_cov/core/samplers.py:36-109) in mode "greedy_conditional_entropy".  The set-up mirrors
precalc_entropy_vars (samplers.py:115-208); the 63-step greedy loop (samplers.py:211-302) runs entirely on
the device (csrc/depthcov.cu) without host round trips.  Index selection is the bit-exact contract.
"""
import torch

from como_b200 import _lib
from como_b200 import como_backends


def get_coords_domain(cov_params_img, border=0):
    b, c, h, w = cov_params_img.shape
    dev = cov_params_img.device
    rr, cc = torch.meshgrid(torch.arange(border, h - border, device=dev), torch.arange(border, w - border, device=dev),
                            indexing="ij")
    vec = torch.stack((rr.reshape(-1), cc.reshape(-1)), dim=1)
    return vec.unsqueeze(0).repeat(b, 1, 1)


def _normalize(x_pixel, dims):
    # como/utils/coords.py:12-15: A = 1/dims in the coordinate dtype (integer coords give a float32 A)
    A = 1.0 / torch.as_tensor(dims, device=x_pixel.device, dtype=x_pixel.dtype)
    return 2 * A * x_pixel + A - 1


def _interp_cov(cov_img, x_norm):
    # gaussian_kernel.py:52-79 (bilinear, border padding); x_norm (B,N,2) in (row, col) order
    grid = torch.stack((x_norm[..., 1], x_norm[..., 0]), dim=-1).unsqueeze(1)
    s = torch.nn.functional.grid_sample(cov_img, grid, mode="bilinear", padding_mode="border", align_corners=False)
    return torch.permute(s.squeeze(2), (0, 2, 1)).reshape(cov_img.shape[0], -1, 2, 2)


def sample_sparse_coords(cov_params_img, num_samples, mode, max_stdev_thresh=-1e8, border=0, terminate_early=False,
                         dist_thresh=0.0, signal_var=None, fixed_var=None, curr_coords=None, curr_var=None,
                         coords_domain=None, dtype=torch.float):
    if mode != "greedy_conditional_entropy":
        raise ValueError("sample_sparse_coords mode: " + mode + " is not implemented.")
    if dtype != torch.float:
        raise NotImplementedError("como_b200 sampler runs in float32 like the reference kernels")
    dev = _lib.require_cuda(cov_params_img)
    b = cov_params_img.shape[0]
    img_size = cov_params_img.shape[-2:]
    cov = cov_params_img.to(dtype=dtype)
    if curr_coords is None:
        curr_coords = torch.empty((b, 0, 2), device=dev, dtype=dtype)
    if curr_var is None:
        curr_var = torch.zeros((b, 0), device=dev, dtype=dtype)
    if coords_domain is None:
        coords_domain = get_coords_domain(cov, border=border)
        dom_norm = _normalize(coords_domain, img_size).to(dtype)
        E_dom = torch.permute(cov[:, :, coords_domain[0, :, 0], coords_domain[0, :, 1]], (0, 2, 1)).reshape(b, -1, 2, 2)
        E_dom = E_dom.contiguous()
    else:
        dom_norm = _normalize(coords_domain, img_size).to(dtype)
        E_dom = _interp_cov(cov, dom_norm).contiguous()
    d = dom_norm.shape[1]
    n = min(int(num_samples), d)
    scale = float(signal_var)
    curr_norm = _normalize(curr_coords, img_size).to(dtype)
    m = curr_norm.shape[1]

    sel_idx = torch.full((b, n), -1, dtype=torch.int64, device=dev)
    sel_xy = torch.zeros((b, n, 2), dtype=dtype, device=dev)
    sel_E = torch.zeros((b, n, 2, 2), dtype=dtype, device=dev)
    L = torch.eye(n, dtype=dtype, device=dev).unsqueeze(0).repeat(b, 1, 1).contiguous()
    obs_info = torch.zeros((b, n, d), dtype=dtype, device=dev)
    bi = torch.arange(b, device=dev)
    if m > 0:
        m = min(m, n)
        sel_xy[:, :m] = curr_norm[:, :m]
        sel_E[:, :m] = _interp_cov(cov, curr_norm[:, :m])
    else:
        areas = E_dom[..., 0, 0] * E_dom[..., 1, 1] - E_dom[..., 0, 1] * E_dom[..., 1, 0]
        best = torch.argmax(areas.view(b, -1), dim=1)
        sel_idx[:, 0] = best
        sel_xy[:, 0] = dom_norm[bi, best]
        sel_E[:, 0] = E_dom[bi, best]
        m = 1
    K_nn = como_backends.cross_covariance(sel_xy[:, :m], sel_E[:, :m], sel_xy[:, :m].clone(), sel_E[:, :m].clone(), scale)
    if curr_var.shape[1] > 0:
        K_nn = K_nn + torch.diag_embed(curr_var[:, :m])
    has_fixed = fixed_var is not None
    if has_fixed:
        K_nn = K_nn + torch.diag_embed(float(fixed_var) * torch.ones(b, m, device=dev, dtype=dtype))
    L[:, :m, :m] = torch.linalg.cholesky(K_nn, upper=False)
    K_md = como_backends.cross_covariance(sel_xy[:, :m], sel_E[:, :m], dom_norm.reshape(b, -1, 2), E_dom, scale)
    if m == 1:
        obs_info[:, :1] = K_md / L[:, :1, :1]
    else:
        obs_info[:, :m] = torch.linalg.solve_triangular(L[:, :m, :m], K_md, upper=False)
    var = (scale - torch.sum(obs_info[:, :m] * obs_info[:, :m], dim=1)).contiguous()
    d2 = torch.sum(torch.square(sel_xy[:, :m, None, :] - dom_norm[:, None, :, :]), dim=-1)
    dist_ok = (d2 > float(dist_thresh) * float(dist_thresh)).all(dim=1).to(torch.uint8).contiguous()

    count = torch.zeros(b, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        ws = torch.empty(int(_lib.sampler_workspace_bytes(b, d)), dtype=torch.uint8, device=dev)
        st = _lib.sampler_greedy(_lib.ptr(dom_norm.contiguous()), _lib.ptr(E_dom), b, d, n, m, _lib.ptr(sel_xy),
                                 _lib.ptr(sel_E), _lib.ptr(sel_idx), _lib.ptr(L), _lib.ptr(obs_info), _lib.ptr(var),
                                 _lib.ptr(dist_ok), scale, float(fixed_var) if has_fixed else 0.0, 1 if has_fixed else 0,
                                 float(dist_thresh), float(max_stdev_thresh), 1 if terminate_early else 0,
                                 _lib.ptr(count), _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev))
    _lib.check(st, "como_b200_sampler_greedy")
    cnt = int(count[0].item())
    inds = sel_idx[:, :cnt]
    domain_inds = inds[:, inds[0, :] >= 0]
    batch_inds = torch.arange(b, device=dev).unsqueeze(1).repeat(1, domain_inds.shape[1])
    return coords_domain[batch_inds, domain_inds, :], domain_inds
