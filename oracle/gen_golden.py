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
        return t.detach().cpu().numpy()
    return np.asarray(t)


def ref_cfg():
    return yaml.safe_load(open(os.path.join(ref_harness.REF, "config", "como.yml")))


def make_tracker(H, W, end_level, max_iter=50):
    ref_harness.load_reference()
    import como.odom.Tracking as TR

    TR.init_gpu = lambda d: None
    cfg = copy.deepcopy(ref_cfg()["tracking"])
    cfg["device"] = "cpu"
    cfg["pyr"]["end_level"] = end_level
    cfg["term_criteria"]["max_iter"] = max_iter
    K = synth.make_intrinsics(H, W)
    tr = TR.Tracking(cfg, K, (H, W))
    tr.setup()
    return tr, cfg


def gen_track(name, H, W, end_level, max_iter, cell):
    """photo_tracking_pyr inputs/outputs + per-iteration trace + handle_frame decisions."""
    ref_harness.load_reference()
    import como.odom.frontend.photo_tracking as PT

    torch.manual_seed(0)
    tr, cfg = make_tracker(H, W, end_level, max_iter)
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


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    os.makedirs(GOLD, exist_ok=True)
    if what in ("track", "all"):
        gen_track("track_80x60_l3", 60, 80, 3, 50, 4)
        gen_track("track_80x60_l3_it1", 60, 80, 3, 1, 4)  # BASELINE config 1 shape (max_iter 1), shrunk
        gen_track("track_160x120_l4", 120, 160, 4, 50, 8)


if __name__ == "__main__":
    main()
