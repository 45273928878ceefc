"""Pins oracle/depthcov_oracle.py against reference-generated goldens (CPU)."""
import os

import numpy as np
import pytest
import torch

from oracle import depthcov_oracle as DO


def G(golden_dir):
    return np.load(os.path.join(golden_dir, "depthcov.npz"))


def test_cross_covariance_numpy_restatement(golden_dir):
    g = G(golden_dir)
    K = DO.cross_covariance_np(g["cc_x1"], g["cc_E1"], g["cc_x2"], g["cc_E2"], float(g["cc_scale"]))
    np.testing.assert_allclose(K, g["cc_K"], rtol=3e-7, atol=0)
    assert (K == g["cc_K"]).mean() > 0.9  # almost everywhere bit identical


def test_ref_backend_if_built_matches_golden(golden_dir):
    rb = DO.ref_backends()
    if rb is None:
        pytest.skip("oracle/_ref not built")
    g = G(golden_dir)
    t = lambda k: torch.from_numpy(g[k])
    K = rb.cross_covariance(t("cc_x1"), t("cc_E1"), t("cc_x2"), t("cc_E2"), float(g["cc_scale"]))
    np.testing.assert_array_equal(K.numpy(), g["cc_K"])


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_sampler_indices_bit_exact(golden_dir, tag):
    g = G(golden_dir)
    cov = torch.from_numpy(g[f"s{tag}_cov"])
    n = int(g[f"s{tag}_n"])
    coords, inds = DO.sample_sparse_coords(cov, n, max_stdev_thresh=1e-2, border=3, dist_thresh=0.1, signal_var=1.0,
                                           fixed_var=0.0)
    np.testing.assert_array_equal(inds.numpy(), g[f"s{tag}_inds"])
    np.testing.assert_array_equal(coords.numpy(), g[f"s{tag}_coords"])
    keep = torch.from_numpy(g[f"s{tag}_coords"])[:, : n // 2].float()
    c2, i2 = DO.sample_sparse_coords(cov, n, max_stdev_thresh=1e-2, border=3, dist_thresh=0.1, signal_var=1.0,
                                     fixed_var=0.0, curr_coords=keep)
    np.testing.assert_array_equal(i2.numpy(), g[f"s{tag}_inds2"])


def test_prep_predictor_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "ba_k4_notfull.npz"))
    cov = torch.from_numpy(g["in_cov_params_img"])
    pm = torch.from_numpy(g["in_pm_first_obs"])
    coords_m = torch.stack((pm[..., 1], pm[..., 0]), -1)
    Kinv, L, KK = DO.prep_predictor(cov, coords_m, float(g["gp_scale"]))
    rel = lambda a, b: float((a - torch.from_numpy(b)).abs().max() / np.abs(b).max())
    # K_mm is ill conditioned (jitter 1e-6; tracked anchors can sit close together): last-bit differences
    # in K_mm are amplified by cond(K_mm) ~ 1e7 in the factor and the predictor
    assert rel(L, g["in_L_mm"]) < 1e-7
    assert rel(Kinv, g["in_Kmm_inv"]) < 1e-5
    assert rel(KK, g["in_Knm_Kmminv"]) < 1e-6
    # keyframe 0 has integer anchor coordinates and a well separated anchor set: much tighter
    assert float((L[0] - torch.from_numpy(g["in_L_mm"][0])).abs().max()) < 1e-9
