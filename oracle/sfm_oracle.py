"""ORACLE (test infrastructure, never on the product path): two-frame SfM bootstrap on the CPU.

Restates two_frame_sfm_pyr / two_frame_sfm / construct_photo_system / linearize_photo / robustify_photo / the
two depth priors / update_vars (como/odom/frontend/two_frame_sfm.py:15-392) in plain torch-CPU float64.  The
reference materialises dI/dd as an (N, M) matrix; here it is the rank-one form  dI/dd_n = beta_n k_n  with
beta_n = dI/dP_i . P_i  (because dP_i/dlogz = P_i) and k_n the predictor row, the same algebra the CUDA path uses.
Pinned by tests/golden/sfm_64x48.npz, generated from the unmodified reference (oracle/gen_golden.py sfm).
"""
import torch

from oracle import ba_oracle as BO

F64 = torch.float64
HUBER_K = 1.345


def linearize(T, d, coords_rc, vals, Knm, img_j, Kmat):
    """One linearisation: returns r (N), valid (N) bool, J_T (N,6), beta (N), logz (N), Pj (N,3), pj (N,2) [x,y]."""
    logz = (Knm @ d).reshape(-1)                                  # geometry/depth.py:21-24
    z = torch.exp(logz)
    x, y = coords_rc[:, 1].to(F64), coords_rc[:, 0].to(F64)
    Pi = torch.stack(((x - Kmat[0, 2]) / Kmat[0, 0], (y - Kmat[1, 2]) / Kmat[1, 1], torch.ones_like(x)), -1) * z[:, None]
    R, t = T[:3, :3], T[:3, 3]
    Pj = Pi @ R.T + t
    u = Kmat[0, 0] * Pj[:, 0] / Pj[:, 2] + Kmat[0, 2]
    v = Kmat[1, 1] * Pj[:, 1] / Pj[:, 2] + Kmat[1, 2]
    Himg, Wimg = img_j.shape[-2:]
    valid = (u >= 1) & (u < Wimg - 1) & (v >= 1) & (v < Himg - 1) & (Pj[:, 2] > 0)     # photo_utils.py:12-19, :194
    # grid_sample(zeros padding) at the normalised/unnormalised coordinate.  Reference quirk (two_frame_sfm.py:186-189):
    # A_norm = 1.0 / torch.as_tensor((W, H)) is a FLOAT32 tensor (integer input), so the normalisation uses
    # float32(1/W), float32(1/H) inside float64 arithmetic: a 3e-8 relative shift of the sampling position
    Ax = float(torch.tensor(1.0, dtype=torch.float32) / torch.tensor(float(Wimg), dtype=torch.float32))
    Ay = float(torch.tensor(1.0, dtype=torch.float32) / torch.tensor(float(Himg), dtype=torch.float32))
    un = ((2 * Ax * u + Ax - 1) + 1) * Wimg / 2 - 0.5
    vn = ((2 * Ay * v + Ay - 1) + 1) * Himg / 2 - 0.5
    s = BO.bilinear3(img_j[0], un, vn)                            # (3, N): I, gx, gy
    r = s[0] - vals
    dIdw = torch.stack((s[1], s[2]), -1)                          # (N, 2)
    iz = 1.0 / Pj[:, 2]
    # dI/dPj = dI/dw dpj/dPj  (camera.py:27-36)
    dIdP = torch.stack((dIdw[:, 0] * Kmat[0, 0] * iz, dIdw[:, 1] * Kmat[1, 1] * iz,
                        -(dIdw[:, 0] * Kmat[0, 0] * Pj[:, 0] + dIdw[:, 1] * Kmat[1, 1] * Pj[:, 1]) * iz * iz), -1)
    dIdPi = dIdP @ R                                              # row vector times R (transforms.py:30-32)
    # dPj/dT = [-R Pi^ | R] (transforms.py:24-28): dI/dT = [-(dI/dPi) Pi^ | dI/dPi], and  a^T (-[p]x) = (p x a)^T
    JT = torch.cat((torch.linalg.cross(Pi, dIdPi), dIdPi), -1)
    beta = (dIdPi * Pi).sum(-1)
    return r, valid, JT, beta, logz, Pj, torch.stack((u, v), -1)


def level(T, d, coords_rc, vals, Knm, img_j, Kmat, dr_prior, H_prior, init_cfg, trace=None):
    """two_frame_sfm (two_frame_sfm.py:306-392) for one pyramid level.  T (4,4), d (M,1)."""
    M = d.shape[0]
    N = Knm.shape[0]
    dr_mean = Knm.sum(0, keepdim=True) / N                         # :127-133
    it, prev = 0, float("inf")
    while True:
        r, valid, JT, beta, logz, Pj, pj = linearize(T, d, coords_rc, vals, Knm, img_j, Kmat)
        sigma = 1.4826 * torch.median(torch.abs(r[valid]))         # :259-262 (lower median)
        wr = r / sigma
        w = torch.where(wr.abs() < HUBER_K, torch.ones_like(wr), HUBER_K / wr.abs())
        w = torch.where(valid, w, torch.zeros_like(w))
        sc = torch.sqrt(w) / sigma
        photo_err = torch.sum((torch.sqrt(w) * wr) ** 2)
        rs, Js, bs = r * sc, JT * sc[:, None], beta * sc
        H = torch.zeros(6 + M, 6 + M, dtype=F64)
        g = torch.zeros(6 + M, dtype=F64)
        g[:6] = -(Js * rs[:, None]).sum(0)
        g[6:] = -((bs * rs)[:, None] * Knm).sum(0)
        H[:6, :6] = Js.T @ Js
        H[6:, 6:] = (Knm * (bs * bs)[:, None]).T @ Knm
        HTd = (Js * bs[:, None]).T @ Knm
        H[:6, 6:] = HTd
        H[6:, :6] = HTd.T
        rp = dr_prior @ d                                           # :136-146
        prior_err = torch.sum(rp ** 2)
        g[6:] -= (dr_prior * rp).sum(0)
        H[6:, 6:] += H_prior
        rm = logz.mean()                                            # :149-163, sigma = 1
        g[6:] -= dr_mean[0] * rm
        H[6:, 6:] += dr_mean.T @ dr_mean
        total = photo_err + prior_err + rm * rm
        L, _ = torch.linalg.cholesky_ex(H)
        delta = torch.cholesky_solve(g[:, None], L)
        if trace is not None and it == 0:
            trace.update(H0=H.clone(), g0=g.clone(), delta0=delta.clone(), num_valid=int(valid.sum()), sigma0=float(sigma))
        T = T @ BO.se3_exp_batch(delta[:6, 0][None])[0]             # update_vars / batch_se3 (lie_algebra.py:52-56)
        d = d + delta[6:]
        it += 1
        dn = torch.norm(delta[:6])
        dec = prev - total
        rel = torch.abs(torch.as_tensor(dec)) / prev
        done = it >= init_cfg["max_iter"] or dn < init_cfg["delta_norm"] or (rel < init_cfg["rel_tol"] and dec > 0)
        prev = total
        if done:
            break
    return T, d, logz.mean(), pj[valid], Pj[valid, 2], it


def two_frame_sfm_pyr(T0, d0, coords_pyr, vals_pyr, Knm_pyr, img_pyr, K_pyr, dr_prior, H_prior, init_cfg, traces=None):
    """two_frame_sfm_pyr (:15-52): coarse to fine.  Returns (T, d, mean_log_depth, pj_valid [x,y], depths_valid, iters)."""
    T, d = T0.clone(), d0.clone()
    iters = []
    out = None
    for l in range(len(vals_pyr)):
        tr = {} if traces is not None else None
        T, d, mld, pjv, zv, it = level(T, d, coords_pyr[l], vals_pyr[l], Knm_pyr[l], img_pyr[l], K_pyr[l], dr_prior, H_prior,
                                       init_cfg, tr)
        if traces is not None:
            tr.update(T=T.clone(), d=d.clone(), mean_log_depth=float(mld), iters=it)
            traces.append(tr)
        iters.append(it)
        out = (mld, pjv, zv)
    return T, d, out[0], out[1], out[2], iters
