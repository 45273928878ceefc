"""Pins oracle/ba_oracle.py (one Mapping.iterate) against golden vectors recorded from the unmodified
reference (real DepthCov UNet / sampler / correspondence code, CPU fp64)."""
import os

import numpy as np
import pytest
import torch

from oracle import ba_oracle as BO


def rel(a, b):
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    return float((a - b).abs().max() / b.abs().max())


@pytest.mark.parametrize("name", ["ba_k4_notfull", "ba_k4_full"])
def test_iterate_matches_reference(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    s = BO.state_from_golden(g)
    cfg = BO.cfg_from_golden(g)
    for it in range(int(g["iters"])):
        out = BO.iterate(s, cfg)
        if it == 0:
            # integer selections: bit exact
            np.testing.assert_array_equal(out["coords_n"].numpy(), g["coords_n"])
            assert out["pairs"][0] == list(g["kf_ref_ids"]) and out["pairs"][1] == list(g["kf_target_ids"])
            assert out["pairs"][2] == list(g["one_way_kf_ids"]) and out["pairs"][3] == list(g["one_way_target_ids"])
            assert rel(out["Pwn"], g["Pwn0"]) < 1e-13
            assert rel(out["H_photo"], g["H0_photo"]) < 1e-12
            assert rel(out["g_photo"], g["g0_photo"]) < 1e-12
            assert rel(out["H"], g["H0"]) < 1e-12
            assert rel(out["g"], g["g0"]) < 1e-12
            assert rel(out["delta"], g["delta0"][:, 0]) < 1e-8
        assert abs(out["photo_err"] - float(g[f"it{it}_photo_err"])) <= 1e-9 * float(g[f"it{it}_photo_err"])
        assert abs(out["total_err"] - float(g[f"it{it}_total_err"])) <= 1e-9 * float(g[f"it{it}_total_err"])
        assert rel(s["kf_poses"], g[f"it{it}_kf_poses"]) < 1e-8
        assert rel(s["kf_aff_params"], g[f"it{it}_kf_aff_params"]) < 1e-7
        assert rel(s["recent_poses"], g[f"it{it}_recent_poses"]) < 1e-8
        assert rel(s["P_m"], g[f"it{it}_P_m"]) < 1e-8
        assert rel(s["median_depths"], g[f"it{it}_median_depths"]) < 1e-10
        assert rel(s["depth_imgs"], g[f"it{it}_depth_imgs"]) < 1e-10


def test_iterate_matches_reference_k8_m64_256x192(golden_dir):
    """BASELINE anchor count and window size class: K = 8 keyframes, R = 6 one-way frames, M = 64 anchors at the
    network resolution 256x192 (dim 511), built by the reference's own keyframe pipeline.  The fixture is compact
    (texture + offsets, covariance images + anchors, predictor rows at the sampled pixels); the regenerated image
    stack and predictor slab must agree with the reference's checksums before the iteration is compared."""
    g = np.load(os.path.join(golden_dir, "ba_k8_m64_256x192.npz"))
    s, rep = BO.state_from_compact_golden(g)
    assert rep["img0"] == 0.0 and rep["img_sum"] < 1e-10 and rep["rec_sum"] < 1e-10, rep
    assert rep["rows"] < 1e-12 and rep["colsum"] < 1e-12, rep
    cfg = BO.cfg_from_golden(g)
    out = BO.iterate(s, cfg)
    np.testing.assert_array_equal(out["coords_n"].numpy(), g["coords_n"])
    assert out["pairs"][0] == list(g["kf_ref_ids"]) and out["pairs"][1] == list(g["kf_target_ids"])
    assert out["pairs"][2] == list(g["one_way_kf_ids"]) and out["pairs"][3] == list(g["one_way_target_ids"])
    assert rel(out["H_photo"], g["H0_photo"]) < 1e-12
    assert rel(out["g_photo"], g["g0_photo"]) < 1e-12
    assert rel(out["H"], g["H0"]) < 1e-12
    assert rel(out["g"], g["g0"]) < 1e-11
    assert rel(out["delta"], g["delta0"][:, 0]) < 1e-8
    assert abs(out["photo_err"] - float(g["it0_photo_err"])) <= 1e-9 * float(g["it0_photo_err"])
    assert abs(out["total_err"] - float(g["it0_total_err"])) <= 1e-9 * float(g["it0_total_err"])
    assert rel(s["kf_poses"], g["it0_kf_poses"]) < 1e-8
    assert rel(s["recent_poses"], g["it0_recent_poses"]) < 1e-8
    assert rel(s["P_m"], g["it0_P_m"]) < 1e-8
    assert rel(s["median_depths"], g["it0_median_depths"]) < 1e-10
    assert rel(s["depth_imgs"][:, :, ::8, ::8], g["it0_depth_imgs_sub8"]) < 1e-10
