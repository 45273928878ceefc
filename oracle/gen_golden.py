"""TEST INFRASTRUCTURE ONLY: generates tests/golden/*.npz by running the UNMODIFIED reference on CPU.

Run in the authoring container only:  python oracle/gen_golden.py [track|kfref|cov|ba|all]
The committed fixtures are what travels to the GPU box (there is no /root/reference there).
"""
import copy
import os
import sys

import numpy as np
import torch
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_harness  # noqa: E402
from como_b200 import synth  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def _np(t):
    if isinstance(t, torch.Tensor):
        return t.detach().cpu().numpy().copy()  # copy: the reference mutates several tensors in place later
    return np.array(t)


def ref_cfg():
    return yaml.safe_load(open(os.path.join(ref_harness.REF, "config", "como.yml")))


def make_tracker(H, W, end_level, max_iter=50, color="gray"):
    ref_harness.load_reference()
    import como.odom.Tracking as TR

    TR.init_gpu = lambda d: None
    cfg = copy.deepcopy(ref_cfg()["tracking"])
    cfg["device"] = "cpu"
    cfg["color"] = color
    cfg["pyr"]["end_level"] = end_level
    cfg["term_criteria"]["max_iter"] = max_iter
    K = synth.make_intrinsics(H, W)
    tr = TR.Tracking(cfg, K, (H, W))
    tr.setup()
    return tr, cfg


def gen_track(name, H, W, end_level, max_iter, cell, color="gray"):
    """photo_tracking_pyr inputs/outputs + per-iteration trace + handle_frame decisions."""
    ref_harness.load_reference()
    import como.odom.frontend.photo_tracking as PT

    torch.manual_seed(0)
    tr, cfg = make_tracker(H, W, end_level, max_iter, color)
    rgb = synth.make_rgb(H, W, seed=0, cell=cell)
    depth = synth.make_depth(H, W)
    pose = torch.eye(4)[None]
    aff = torch.zeros(1, 2, 1)
    tr.update_kf_reference(([1.0], rgb, pose, aff, depth))
    # second call exercises the "same timestamp" branch (geometry only); must not change anything
    out = {"H": H, "W": W, "end_level": end_level, "max_iter": max_iter}
    out["rgb"] = _np(rgb)
    out["depth"] = _np(depth)
    out["K"] = _np(tr.intrinsics)
    nl = len(tr.vals_pyr)
    out["num_levels"] = nl
    for l in range(nl):
        out[f"vals_{l}"] = _np(tr.vals_pyr[l])
        out[f"P_{l}"] = _np(tr.P_pyr[l])
        out[f"dI_dT_{l}"] = _np(tr.dI_dT_pyr[l])
        out[f"mask_{l}"] = _np(tr.mask_pyr[l])
        out[f"K_{l}"] = _np(tr.intrinsics_pyr[l])
        out[f"grads_{l}"] = _np(tr.img_grads_pyr[l])
    T0 = synth.se3_exp_wv(synth.TRACK_PERTURB).float()[None]
    tr.T_curr_kf = T0.clone()
    out["T_init"] = _np(T0)
    out["aff_init"] = _np(tr.aff_curr_kf)
    # second frame: brightness-changed copy of the KF image so the affine terms are exercised
    # plus independent sensor noise so residuals at convergence stay O(1e-2) (a noise-free copy makes the
    # converged residual pure fp32 cancellation error, which no two implementations agree on)
    gen = torch.Generator().manual_seed(123)
    rgb2 = (rgb * 1.05 + 0.02 + 0.03 * (torch.rand(rgb.shape, generator=gen) - 0.5)).clamp(0, 1)
    out["rgb2"] = _np(rgb2)
    img_pyr = tr.prep_tracking_img(rgb2)
    for l in range(nl):
        out[f"img_{l}"] = _np(img_pyr[l])

    trace = []
    orig_iter = PT.tracking_iter

    def wrapped(Tji, Pi, intr, img_j, aff_, vals_i, dI_dT, photo_sigma, A_norm):
        res = orig_iter(Tji, Pi, intr, img_j, aff_, vals_i, dI_dT, photo_sigma, A_norm)
        Tn, an, delta, mse, gn, pj, vm, dj = res
        trace.append(
            dict(n=Pi.shape[1], T_in=_np(Tji), aff_in=_np(aff_), T_out=_np(Tn), aff_out=_np(an),
                 delta=_np(delta), mse=float(mse), gnorm=float(gn), nvalid=int(vm.sum()))
        )
        return res

    PT.tracking_iter = wrapped
    try:
        viz, map_data = tr.handle_frame((2.0, rgb2))
    finally:
        PT.tracking_iter = orig_iter
    out["T_final"] = _np(tr.T_curr_kf)
    out["aff_final"] = _np(tr.aff_curr_kf)
    out["trace_n"] = np.array([t["n"] for t in trace])
    out["trace_T_in"] = np.stack([t["T_in"] for t in trace])
    out["trace_aff_in"] = np.stack([t["aff_in"] for t in trace])
    out["trace_T_out"] = np.stack([t["T_out"] for t in trace])
    out["trace_aff_out"] = np.stack([t["aff_out"] for t in trace])
    out["trace_delta"] = np.stack([t["delta"] for t in trace])
    out["trace_mse"] = np.array([t["mse"] for t in trace])
    out["trace_gnorm"] = np.array([t["gnorm"] for t in trace])
    out["trace_nvalid"] = np.array([t["nvalid"] for t in trace])
    # keyframe decision inputs (a10)
    reproj = tr.get_reproj_last_kf(tr.T_curr_kf)
    vmask = ~torch.isnan(reproj)
    out["reproj_count"] = int(torch.count_nonzero(vmask))
    out["reproj_median"] = float(torch.median(reproj[vmask]))
    out["decision"] = "none" if map_data is None else map_data[0]
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    print(name, "iters", len(trace), "levels", nl, "decision", out["decision"],
          "final t", out["T_final"][0, :3, 3], "aff", out["aff_final"].ravel())


MAP_STATE_TENSORS = [
    "kf_poses", "kf_aff_params", "recent_poses", "recent_aff_params", "kf_img_and_grads", "recent_img_and_grads",
    "cov_params_img", "pm_first_obs", "pm", "logzm", "L_mm", "Kmm_inv", "Knm_Kmminv", "correspondence_mask", "P_m",
    "obs_ref_mask", "pose_anchor", "aff_anchor", "median_depths", "depth_imgs",
]


def snapshot_mapping(m, prefix, out):
    for k in MAP_STATE_TENSORS:
        out[prefix + k] = _np(getattr(m, k))
    out[prefix + "kf_timestamps"] = np.array(m.kf_timestamps, dtype=np.float64)
    out[prefix + "recent_timestamps"] = np.array(m.recent_timestamps, dtype=np.float64)
    out[prefix + "window_full"] = bool(m.window_full)
    out[prefix + "init_scale_anchor"] = _np(m.init_scale_anchor)
    out[prefix + "intrinsics"] = _np(m.intrinsics)
    if hasattr(m, "P_m_anchors"):
        out[prefix + "P_m_anchors"] = _np(m.P_m_anchors)


def build_reference_window(H, W, NKF, NOW, M, cfg_num_kf, step=3.0):
    """Drives the reference's own init_keyframe/add_keyframe/add_one_way_frame (real DepthCov UNet,
    sampler and correspondence code) on the synthetic translating-plane scene of SURVEY 8d."""
    ref_harness.load_reference()
    import como.odom.Mapping as MP
    from como.depth_cov.core.samplers import sample_sparse_coords

    MP.init_gpu = lambda d: None
    cfg = copy.deepcopy(ref_cfg()["mapping"])
    cfg["device"] = "cpu"
    cfg["model_path"] = os.path.join(ref_harness.REF, "models", "scannet.ckpt")
    cfg["graph"]["num_keyframes"] = cfg_num_kf
    cfg["sampling"]["max_num_coords"] = M
    f = 525.0 * W / 640
    Z = 2.0
    K = torch.tensor([[f, 0, W / 2], [0, f, H / 2], [0, 0, 1]])
    tex = synth.make_rgb(H, W, seed=0, cell=8, extra_w=int(step * (NKF + 2)) + 8).double()

    x0_log = []

    def frame(k):
        dx = k * step
        x0 = int(dx)
        x0_log.append(x0)
        rgb = tex[..., x0:x0 + W].clone()
        T = torch.eye(4, dtype=torch.double)[None]
        T[0, 0, 3] = dx * Z / f
        return rgb, T

    torch.manual_seed(0)
    m = MP.Mapping(cfg, K)
    m.setup()
    rgb0, T0 = frame(0)
    cov0 = m.run_model(rgb0)
    with torch.no_grad():
        coords_m, _ = sample_sparse_coords(
            cov0, M, mode="greedy_conditional_entropy", max_stdev_thresh=1e-2, border=3, dist_thresh=0.1,
            signal_var=m.model.get_scale(-1), fixed_var=0.0)
    coords_m = coords_m.double()
    assert coords_m.shape[1] == M, coords_m.shape
    logz = torch.log(torch.full((1, M, 1), Z, dtype=torch.double))
    m.init_keyframe(rgb0, cov0, coords_m, T0, logz, torch.zeros(1, 2, 1, dtype=torch.double), 1.0)
    m.init_scale_anchor = torch.log(torch.tensor(Z, dtype=torch.double)).view(1, 1, 1)
    m.is_init = True
    for k in range(1, NKF):
        rgb, T = frame(k)
        # perturbed initial pose/affine so the optimiser has something to do
        T = T.clone()
        T[0, 0, 3] += 0.002 * ((-1) ** k)
        T[0, 1, 3] += 0.0007 * k + 0.0003  # distinct per keyframe: equal offsets put rows exactly on the v = 1 border
        m.add_keyframe(rgb, T, torch.tensor([[[0.01 * k], [-0.005 * k]]], dtype=torch.double), 1.0 + k)
    nwin = m.kf_poses.shape[0]
    t0 = m.kf_timestamps[0]
    for j in range(NOW):
        kk = (j % (nwin - 1)) + 0.5 + (t0 - 1.0)
        rgb, T = frame(kk)
        T = T.clone()
        T[0, 1, 3] += 0.00053 + 0.00021 * j   # keep projections off the exact v = 1 / v = H-2 rows
        T[0, 0, 3] += 0.0011 * ((-1) ** j)
        m.add_one_way_frame(rgb, T, torch.zeros(1, 2, 1, dtype=torch.double), 1.0 + kk + 0.001 * j)
    # the reference keeps recent frames time ordered; emulate arrival order by sorting all recent state
    order = sorted(range(len(m.recent_timestamps)), key=lambda i: m.recent_timestamps[i])
    m.recent_timestamps = [m.recent_timestamps[i] for i in order]
    m.recent_img_and_grads = m.recent_img_and_grads[order]
    m.recent_poses = m.recent_poses[order]
    m.recent_aff_params = m.recent_aff_params[order]
    # bookkeeping for the compact goldens (frames are slices of one texture)
    m._golden_tex = tex
    m._golden_kf_x0 = x0_log[:NKF][-nwin:]
    m._golden_recent_x0 = [x0_log[NKF + i] for i in order]
    return m, cfg


def gen_ba(name, H, W, NKF, NOW, M, cfg_num_kf, iters=3):
    """Full Mapping.iterate golden: state in, (H, g, delta, pairs, coords_n, errors) of the first
    iteration, state after every iteration."""
    ref_harness.load_reference()
    import como.odom.backend.linear_system as LS
    import como.odom.backend.photo as PH
    import como.odom.backend.sparse_map as SM
    import como.odom.Mapping as MP

    m, cfg = build_reference_window(H, W, NKF, NOW, M, cfg_num_kf)
    out = {"H": H, "W": W, "M": M, "cfg_num_kf": cfg_num_kf, "iters": iters}
    out["photo_cfg_batch"] = cfg["photo_construction"]["pairwise_batch_size"]
    out["sigma_mean_depth_prior"] = cfg["sigmas"]["mean_depth_prior"]
    out["sigma_scale_prior"] = cfg["sigmas"]["scale_prior"]
    out["sigma_pose_prior"] = cfg["sigmas"]["pose_prior"]
    snapshot_mapping(m, "in_", out)
    out["gp_scale"] = float(m.model.get_scale(-1))
    cap = {}
    orig_solve = LS.solve_system

    def solve_hook(Hm, g):
        d = orig_solve(Hm, g)
        cap.setdefault("H", []).append(_np(Hm).copy())
        cap.setdefault("g", []).append(_np(g).copy())
        cap.setdefault("delta", []).append(_np(d).copy())
        return d

    orig_cps = MP.create_photo_system

    def cps_hook(*a, **k):
        r = orig_cps(*a, **k)
        Hm, g = a[15], a[16]
        cap.setdefault("photo_err", []).append(float(r[0]))
        cap.setdefault("H_photo", []).append(_np(Hm).copy())
        cap.setdefault("g_photo", []).append(_np(g).copy())
        cap.setdefault("pairs", []).append((list(r[1][0]), list(r[1][1]), list(r[2][0]), list(r[2][1])))
        cap.setdefault("Pwn", []).append(_np(a[4]).copy())
        cap.setdefault("vals_n", []).append(_np(a[9]).copy())
        cap.setdefault("median_depths_n", []).append(_np(a[8]).copy())
        return r

    orig_sub = MP.subselect_pixels

    def sub_hook(*a, **k):
        r = orig_sub(*a, **k)
        cap.setdefault("coords_n", []).append(_np(r[0]).copy())
        return r

    LS.solve_system = solve_hook
    MP.lin_sys.solve_system = solve_hook
    MP.create_photo_system = cps_hook
    MP.subselect_pixels = sub_hook
    try:
        for it in range(iters):
            m.iterate()
            out[f"it{it}_total_err"] = float(m.total_err_prev)
            out[f"it{it}_depth_imgs"] = _np(m.depth_imgs)
            out[f"it{it}_kf_poses"] = _np(m.kf_poses)
            out[f"it{it}_kf_aff_params"] = _np(m.kf_aff_params)
            out[f"it{it}_recent_poses"] = _np(m.recent_poses)
            out[f"it{it}_recent_aff_params"] = _np(m.recent_aff_params)
            out[f"it{it}_P_m"] = _np(m.P_m)
            out[f"it{it}_median_depths"] = _np(m.median_depths)
    finally:
        LS.solve_system = orig_solve
        MP.lin_sys.solve_system = orig_solve
        MP.create_photo_system = orig_cps
        MP.subselect_pixels = orig_sub
    out["H0"] = cap["H"][0]
    out["g0"] = cap["g"][0]
    out["delta0"] = cap["delta"][0]
    out["H0_photo"] = cap["H_photo"][0]
    out["g0_photo"] = cap["g_photo"][0]
    for it in range(iters):
        out[f"it{it}_photo_err"] = cap["photo_err"][it]
    pr = cap["pairs"][0]
    out["kf_ref_ids"], out["kf_target_ids"] = np.array(pr[0]), np.array(pr[1])
    out["one_way_kf_ids"], out["one_way_target_ids"] = np.array(pr[2]), np.array(pr[3])
    out["coords_n"] = cap["coords_n"][0]
    out["Pwn0"] = cap["Pwn"][0]
    out["vals_n0"] = cap["vals_n"][0]
    out["median_depths_n0"] = cap["median_depths_n"][0]
    # drop bulky fields the tests do not need
    for k in list(out.keys()):
        if k.endswith("_rgb") or k.endswith("depth_imgs") and not k.startswith("it"):
            pass
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    sz = os.path.getsize(os.path.join(GOLD, name + ".npz"))
    print(name, "K", m.kf_poses.shape[0], "R", m.recent_poses.shape[0], "L", m.P_m.shape[0], "dim", out["H0"].shape,
          "pairs", len(pr[0]), len(pr[2]), "full", m.window_full, "errs", [out[f"it{i}_total_err"] for i in range(iters)],
          "size %.1f MB" % (sz / 1e6))


def gen_ba_compact(name, H, W, NKF, NOW, M, cfg_num_kf):
    """One Mapping.iterate of the reference on a window at the network resolution with M = 64 anchors (BASELINE
    configs 3/4 shape, shrunk in pixels only).  Compact: instead of the (K,H,W,M) predictor slab, the kf/recent
    image stacks and the dense depth images, it stores what they are made from (covariance images + anchors,
    the texture + frame offsets) plus the slab rows at the sampled pixels, and checksums/subsamples of the rest."""
    ref_harness.load_reference()
    import como.odom.backend.linear_system as LS
    import como.odom.Mapping as MP

    m, cfg = build_reference_window(H, W, NKF, NOW, M, cfg_num_kf)
    out = {"H": H, "W": W, "M": M, "cfg_num_kf": cfg_num_kf, "iters": 1, "compact": True}
    out["photo_cfg_batch"] = cfg["photo_construction"]["pairwise_batch_size"]
    out["sigma_mean_depth_prior"] = cfg["sigmas"]["mean_depth_prior"]
    out["sigma_scale_prior"] = cfg["sigmas"]["scale_prior"]
    out["sigma_pose_prior"] = cfg["sigmas"]["pose_prior"]
    tmp = {}
    snapshot_mapping(m, "in_", tmp)
    heavy = ("in_Knm_Kmminv", "in_kf_img_and_grads", "in_recent_img_and_grads", "in_depth_imgs", "in_cov_params_img")
    for k, v in tmp.items():
        if k not in heavy:
            out[k] = v
    cov = tmp["in_cov_params_img"]
    assert np.array_equal(cov.astype(np.float32).astype(np.float64), cov), "covariance images are not float32 values"
    out["in_cov_params_img_f32"] = cov.astype(np.float32)
    out["tex"] = _np(m._golden_tex)
    out["kf_x0"] = np.array(m._golden_kf_x0)
    out["recent_x0"] = np.array(m._golden_recent_x0)
    kfi, rci = tmp["in_kf_img_and_grads"], tmp["in_recent_img_and_grads"]
    out["chk_kf_img_and_grads_sum"] = kfi.sum(axis=(2, 3))
    out["chk_recent_img_and_grads_sum"] = rci.sum(axis=(2, 3))
    out["chk_kf_img_and_grads_0"] = kfi[0]                  # one full keyframe: pins the regenerated stack
    out["gp_scale"] = float(m.model.get_scale(-1))
    slab = tmp["in_Knm_Kmminv"]
    cap = {}
    orig_solve, orig_cps, orig_sub = LS.solve_system, MP.create_photo_system, MP.subselect_pixels

    def solve_hook(Hm, g):
        d = orig_solve(Hm, g)
        cap["H"], cap["g"], cap["delta"] = _np(Hm), _np(g), _np(d)
        return d

    def cps_hook(*a, **k):
        r = orig_cps(*a, **k)
        cap["photo_err"] = float(r[0])
        cap["H_photo"], cap["g_photo"] = _np(a[15]), _np(a[16])
        cap["pairs"] = (list(r[1][0]), list(r[1][1]), list(r[2][0]), list(r[2][1]))
        return r

    def sub_hook(*a, **k):
        r = orig_sub(*a, **k)
        cap["coords_n"] = _np(r[0])
        return r

    LS.solve_system = solve_hook
    MP.lin_sys.solve_system = solve_hook
    MP.create_photo_system = cps_hook
    MP.subselect_pixels = sub_hook
    try:
        m.iterate()
    finally:
        LS.solve_system = orig_solve
        MP.lin_sys.solve_system = orig_solve
        MP.create_photo_system = orig_cps
        MP.subselect_pixels = orig_sub
    cn = cap["coords_n"]                                    # (K,N,2) [row, col]
    Kw = slab.shape[0]
    out["in_Knm_rows"] = np.stack([slab[k, cn[k, :, 0], cn[k, :, 1]] for k in range(Kw)])   # (K,N,M)
    out["chk_Knm_colsum"] = slab.sum(axis=(1, 2))          # (K,M): pins the regenerated dense slab
    out["coords_n"] = cn
    for k in ("H", "g", "delta", "H_photo", "g_photo"):
        out[k + "0" if not k.endswith("photo") else k.replace("_photo", "0_photo")] = cap[k]
    out["it0_photo_err"] = cap["photo_err"]
    out["it0_total_err"] = float(m.total_err_prev)
    for k in ("kf_poses", "kf_aff_params", "recent_poses", "recent_aff_params", "P_m", "median_depths"):
        out["it0_" + k] = _np(getattr(m, k))
    out["it0_depth_imgs_sub8"] = _np(m.depth_imgs)[:, :, ::8, ::8]
    pr = cap["pairs"]
    out["kf_ref_ids"], out["kf_target_ids"] = np.array(pr[0]), np.array(pr[1])
    out["one_way_kf_ids"], out["one_way_target_ids"] = np.array(pr[2]), np.array(pr[3])
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    sz = os.path.getsize(os.path.join(GOLD, name + ".npz"))
    print(name, "K", m.kf_poses.shape[0], "R", m.recent_poses.shape[0], "L", m.P_m.shape[0], "dim", out["H0"].shape,
          "pairs", len(pr[0]), len(pr[2]), "full", m.window_full, "err", out["it0_total_err"], "size %.1f MB" % (sz / 1e6))


def synth_cov_image(H, W, seed):
    """Smooth SPD 2x2 covariance-parameter image (1,4,H,W) float32, the shape gaussian_kernel.py produces."""
    g = torch.Generator().manual_seed(seed)
    low = torch.rand(1, 3, max(H // 12, 2), max(W // 12, 2), generator=g)
    f = torch.nn.functional.interpolate(low, size=(H, W), mode="bicubic", align_corners=False).clamp(0.02, 0.98)
    x = 2e-3 * torch.exp(3.0 * f[:, 0])
    z = 2e-3 * torch.exp(3.0 * f[:, 1])
    rho = 0.9 * (2 * f[:, 2] - 1)
    off = torch.sqrt(x * z - 1e-8) * rho
    return torch.stack((x, off, off, z), dim=1).float()


def gen_cov(name):
    """como_backends.cross_covariance / get_new_chol_obs_info (reference CPU build) and the reference
    sampler's anchor indices on synthetic covariance images."""
    ref_harness.load_reference()
    import como_backends
    from como.depth_cov.core.samplers import sample_sparse_coords

    out = {}
    g = torch.Generator().manual_seed(7)
    B, n1, n2 = 2, 37, 53
    x1 = torch.rand(B, n1, 2, generator=g) * 2 - 1
    x2 = torch.rand(B, n2, 2, generator=g) * 2 - 1

    def randE(n):
        a = 1e-3 + 5e-2 * torch.rand(B, n, generator=g)
        c = 1e-3 + 5e-2 * torch.rand(B, n, generator=g)
        r = 0.9 * (2 * torch.rand(B, n, generator=g) - 1)
        o = torch.sqrt(a * c) * r
        return torch.stack((a, o, o, c), -1).reshape(B, n, 2, 2)

    E1, E2 = randE(n1), randE(n2)
    out.update(cc_x1=_np(x1), cc_E1=_np(E1), cc_x2=_np(x2), cc_E2=_np(E2), cc_scale=0.83)
    out["cc_K"] = _np(como_backends.cross_covariance(x1, E1, x2, E2, 0.83))
    # chol append sequence
    n, d = 6, 200
    xd = torch.rand(1, d, 2, generator=g) * 2 - 1
    a = 1e-2 + 3e-2 * torch.rand(1, d, generator=g)
    Ed = torch.stack((a, 0 * a, 0 * a, a), -1).reshape(1, d, 2, 2)
    L = torch.eye(n).unsqueeze(0).clone()
    obs = torch.zeros(1, n, d)
    sel = [3, 50, 120, 7, 160, 90]
    sv = 1.0
    K00 = como_backends.cross_covariance(xd[:, sel[:1]], Ed[:, sel[:1]], xd[:, sel[:1]].clone(), Ed[:, sel[:1]].clone(), sv)
    L[:, :1, :1] = torch.linalg.cholesky(K00)
    Kmd = como_backends.cross_covariance(xd[:, sel[:1]], Ed[:, sel[:1]], xd, Ed, sv)
    obs[:, :1] = Kmd / L[:, :1, :1]
    var = sv - torch.sum(obs[:, :1] * obs[:, :1], dim=1)
    out.update(ca_xd=_np(xd), ca_Ed=_np(Ed), ca_sel=np.array(sel), ca_L0=_np(L), ca_obs0=_np(obs), ca_var0=_np(var))
    for i in range(1, n):
        k_ni = como_backends.cross_covariance(xd[:, sel[:i]], Ed[:, sel[:i]], xd[:, sel[i:i + 1]], Ed[:, sel[i:i + 1]], sv)
        k_id = como_backends.cross_covariance(xd[:, sel[i:i + 1]], Ed[:, sel[i:i + 1]], xd, Ed, sv)
        como_backends.get_new_chol_obs_info(L, obs, var, k_ni, k_id, torch.tensor(sv).clone(), i)
    out.update(ca_L=_np(L), ca_obs=_np(obs), ca_var=_np(var))
    # sampler indices
    for tag, (H, W, seed, ns) in {"a": (48, 64, 1, 64), "b": (96, 128, 2, 64), "c": (60, 80, 3, 24)}.items():
        cov = synth_cov_image(H, W, seed)
        with torch.no_grad():
            coords, inds = sample_sparse_coords(cov, ns, mode="greedy_conditional_entropy", max_stdev_thresh=1e-2,
                                                border=3, dist_thresh=0.1, signal_var=torch.tensor(1.0), fixed_var=0.0)
        out[f"s{tag}_cov"] = _np(cov)
        out[f"s{tag}_coords"] = _np(coords)
        out[f"s{tag}_inds"] = _np(inds)
        out[f"s{tag}_n"] = ns
        # continuation from existing coordinates (corr.py:202-213 call shape)
        keep = coords[:, : ns // 2].float()
        with torch.no_grad():
            c2, i2 = sample_sparse_coords(cov, ns, mode="greedy_conditional_entropy", max_stdev_thresh=1e-2, border=3,
                                          dist_thresh=0.1, signal_var=torch.tensor(1.0), fixed_var=0.0, curr_coords=keep)
        out[f"s{tag}_coords2"] = _np(c2)
        out[f"s{tag}_inds2"] = _np(i2)
        print("sampler", tag, tuple(coords.shape), tuple(c2.shape))
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    print(name, "size %.2f MB" % (os.path.getsize(os.path.join(GOLD, name + ".npz")) / 1e6))


def gen_kfinit(name, H, W, NKF, M, iters=2):
    """Keyframe-creation golden: inputs and outputs of the LAST track_and_init call of a window
    built by the reference (corr.py:60-242), with the intermediate results of distill_depth_from_scratch /
    distill_conditional_depth_from_scratch and of the two sampler calls.  A couple of BA iterations run between
    keyframes so that depths and poses are not the trivial initial ones."""
    ref_harness.load_reference()
    import como.odom.Mapping as MP
    import como.odom.frontend.corr as CR

    cap = []
    orig_tai = MP.track_and_init
    orig_dd, orig_dc, orig_ss = CR.distill_depth_from_scratch, CR.distill_conditional_depth_from_scratch, CR.sample_sparse_coords

    def tai_hook(*a, **k):
        rec = {"in": [(_np(x).copy() if torch.is_tensor(x) else x) for x in a[:7]], "corr": dict(a[8]), "samp": dict(a[9]),
               "rgb_img_size": tuple(a[10]), "sub": {}}
        cap.append(rec)
        r = orig_tai(*a, **k)
        rec["out"] = [_np(x).copy() for x in r]
        return r

    def dd_hook(*a, **k):
        r = orig_dd(*a, **k)
        cap[-1]["sub"]["dd_coords_m"] = _np(a[0]).copy()
        cap[-1]["sub"]["dd_logz_m"] = _np(r[0]).copy()
        cap[-1]["sub"]["dd_res_std"] = float(torch.std(r[1]))
        cap[-1]["sub"]["dd_n"] = int(r[1].shape[1])
        return r

    def dc_hook(*a, **k):
        r = orig_dc(*a, **k)
        cap[-1]["sub"]["dc_coords_m"] = _np(a[0]).copy()
        cap[-1]["sub"]["dc_logz_2"] = _np(r).copy()
        return r

    def ss_hook(*a, **k):
        r = orig_ss(*a, **k)
        i = sum(1 for q in cap[-1]["sub"] if q.startswith("ss") and q.endswith("_coords"))
        cap[-1]["sub"][f"ss{i}_coords"] = _np(r[0]).copy()
        cap[-1]["sub"][f"ss{i}_inds"] = _np(r[1]).copy()
        return r

    MP.track_and_init = tai_hook
    CR.distill_depth_from_scratch, CR.distill_conditional_depth_from_scratch, CR.sample_sparse_coords = dd_hook, dc_hook, ss_hook
    orig_add = MP.Mapping.add_keyframe

    def add_hook(self, *a, **k):
        r = orig_add(self, *a, **k)
        for _ in range(iters):
            self.iterate()
        return r

    MP.Mapping.add_keyframe = add_hook
    try:
        m, cfg = build_reference_window(H, W, NKF, 0, M, NKF + 1)
    finally:
        MP.track_and_init = orig_tai
        CR.distill_depth_from_scratch, CR.distill_conditional_depth_from_scratch, CR.sample_sparse_coords = orig_dd, orig_dc, orig_ss
        MP.Mapping.add_keyframe = orig_add
    out = {"H": H, "W": W, "M": M, "ncalls": len(cap), "gp_scale": float(m.model.get_scale(-1))}
    names = ["pose1", "pose2", "coords_m1", "z_m1", "z_img1", "cov_params_img2", "K"]
    for ci, rec in enumerate(cap):
        pre = f"c{ci}_"
        for nm, v in zip(names, rec["in"]):
            out[pre + nm] = v
        for nm, v in zip(["coords_2", "z2", "corr_mask", "coords_all", "z_all"], rec["out"]):
            out[pre + nm] = v
        for k, v in rec["sub"].items():
            out[pre + k] = v
        out[pre + "rgb_img_size"] = np.array(rec["rgb_img_size"])
    for k, v in cap[-1]["corr"].items():
        out["corr_" + k] = v
    for k, v in cap[-1]["samp"].items():
        out["samp_" + k] = v
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    print(name, "calls", len(cap), [(int(r["out"][0].shape[1]), int(r["out"][3].shape[1]), int(r["out"][2].sum())) for r in cap],
          "size %.2f MB" % (os.path.getsize(os.path.join(GOLD, name + ".npz")) / 1e6))


def gen_sfm(name, H, W, M, shift_px=2.5):
    """Two-frame SfM bootstrap golden: the reference's TwoFrameSfm (como/odom/frontend/TwoFrameSfm.py,
    two_frame_sfm.py) initialised on frame 0 and aligned against a translated frame 1.  Captured: everything
    setup_reference produced (test coordinates in the reference's random pixel order, values, predictor pyramids,
    intrinsics pyramid, prior linearisation), the target pyramid, the first normal equations of every level and the
    per-level results."""
    ref_harness.load_reference()
    import como.odom.Mapping as MP
    import como.odom.frontend.two_frame_sfm as TF2

    MP.init_gpu = lambda d: None
    cfg = copy.deepcopy(ref_cfg()["mapping"])
    cfg["device"] = "cpu"
    cfg["model_path"] = os.path.join(ref_harness.REF, "models", "scannet.ckpt")
    cfg["sampling"]["max_num_coords"] = M
    f = 525.0 * W / 640
    K = torch.tensor([[f, 0, W / 2], [0, f, H / 2], [0, 0, 1]])
    tex = synth.make_rgb(H, W, seed=1, cell=8, extra_w=16).double()
    torch.manual_seed(0)
    m = MP.Mapping(cfg, K)
    m.setup()
    sfm = m.two_frame_sfm
    rgb0 = tex[..., 0:W].clone()
    x1 = int(shift_px)
    a = shift_px - x1
    rgb1 = ((1 - a) * tex[..., x1:x1 + W] + a * tex[..., x1 + 1:x1 + 1 + W]).clone()
    sfm.handle_frame(rgb0, 1.0)
    cap = {"levels": [], "solve": []}
    orig_level, orig_solve = TF2.two_frame_sfm, TF2.solve_delta

    def level_hook(*a, **k):
        cap["solve"].append([])
        r = orig_level(*a, **k)
        cap["levels"].append([_np(x).copy() for x in r])
        return r

    def solve_hook(Hm, g):
        d = orig_solve(Hm, g)
        cap["solve"][-1].append((_np(Hm).copy(), _np(g).copy(), _np(d).copy()))
        return d

    TF2.two_frame_sfm, TF2.solve_delta = level_hook, solve_hook
    try:
        img_and_grads1 = sfm.get_img_gradient_pyr(rgb1)
        res = sfm.align_frame([x.clone() for x in img_and_grads1])
    finally:
        TF2.two_frame_sfm, TF2.solve_delta = orig_level, orig_solve
    L = len(sfm.vals_pyr)
    out = {"H": H, "W": W, "M": M, "levels": L, "gp_scale": float(m.model.get_scale(-1)),
           "cov_params_img": _np(sfm.cov_params_img), "coords_m": _np(sfm.coords_m),
           "dr_prior_dd": _np(sfm.dr_prior_dd), "H_prior_d_d": _np(sfm.H_prior_d_d),
           "T_init": _np(sfm.T_curr_kf), "sparse_log_depth_init": _np(sfm.sparse_log_depth),
           "sigma_photo": float(cfg["sigmas"]["photo"])}
    for k, v in cfg["init"].items():
        out["init_" + k] = v
    for l in range(L):
        out[f"l{l}_test_coords"] = _np(sfm.test_coords_pyr[l])
        out[f"l{l}_vals"] = _np(sfm.vals_pyr[l])
        out[f"l{l}_Knm_Kmminv"] = _np(sfm.Knm_Kmminv_pyr[l])
        out[f"l{l}_intrinsics"] = _np(sfm.intrinsics_pyr[l])
        out[f"l{l}_img_and_grads_ref"] = _np(sfm.img_and_grads[l])
        out[f"l{l}_img_and_grads_j"] = _np(img_and_grads1[l])
        Tl, dl, affl, cj, dj, mld = cap["levels"][l]
        out[f"l{l}_T"], out[f"l{l}_sparse_log_depth"], out[f"l{l}_mean_log_depth"] = Tl, dl, mld
        out[f"l{l}_num_valid"] = cj.shape[1]
        out[f"l{l}_iters"] = len(cap["solve"][l])
        out[f"l{l}_H0"], out[f"l{l}_g0"], out[f"l{l}_delta0"] = cap["solve"][l][0]
    out["T_final"], out["sparse_log_depth_final"] = _np(res[0]), _np(res[1])
    out["depth_final_sorted"] = np.sort(_np(res[4]).reshape(-1))
    out["mean_log_depth_final"] = _np(res[5])
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    print(name, "levels", L, "iters", [out[f"l{l}_iters"] for l in range(L)], "valid", [out[f"l{l}_num_valid"] for l in range(L)],
          "t", res[0][0, :3, 3].tolist(), "size %.2f MB" % (os.path.getsize(os.path.join(GOLD, name + ".npz")) / 1e6))


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    os.makedirs(GOLD, exist_ok=True)
    if what in ("track", "all"):
        gen_track("track_80x60_l3", 60, 80, 3, 50, 4)
        gen_track("track_80x60_l3_it1", 60, 80, 3, 1, 4)  # BASELINE config 1 shape (max_iter 1), shrunk
        gen_track("track_80x60_l3_rgb", 60, 80, 3, 50, 4, color="rgb")  # tracking.color: rgb (C = 3)
        gen_track("track_160x120_l4", 120, 160, 4, 50, 8)
    if what in ("cov", "all"):
        gen_cov("depthcov")
    if what in ("ba", "all"):
        gen_ba("ba_k4_notfull", 48, 64, 4, 3, 16, 5)
        gen_ba("ba_k4_full", 48, 64, 5, 3, 16, 4)
    if what in ("ba_compact", "all"):
        gen_ba_compact("ba_k8_m64_256x192", 192, 256, 8, 6, 64, 9)
    if what in ("kfinit", "all"):
        gen_kfinit("kfinit_64x48", 48, 64, 4, 16)
    if what in ("kfinit2", "all"):
        gen_kfinit("kfinit_96x72", 72, 96, 3, 24)     # second shape: pins the oracle only (tests/test_oracle_kfinit.py)
    if what in ("sfm", "all"):
        gen_sfm("sfm_64x48", 48, 64, 16)
    if what in ("sfm2", "all"):
        gen_sfm("sfm_96x72", 72, 96, 24, shift_px=3.25)   # second shape: pins the oracle only (tests/test_oracle_sfm.py)


if __name__ == "__main__":
    main()
