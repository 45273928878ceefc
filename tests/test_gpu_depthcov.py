"""GPU parity for the DepthCov kernels: `como_backends` operators, the device-resident greedy sampler
(bit-exact anchor indices) and the fused K-matrix / predictor."""
import os

import numpy as np
import pytest
import torch

from oracle import depthcov_oracle as DO

pytestmark = pytest.mark.gpu


def G(golden_dir):
    return np.load(os.path.join(golden_dir, "depthcov.npz"))


def test_cross_covariance_vs_reference(golden_dir):
    from como_b200 import como_backends as CB

    g = G(golden_dir)
    c = lambda k: torch.from_numpy(g[k]).cuda()
    K = CB.cross_covariance(c("cc_x1"), c("cc_E1"), c("cc_x2"), c("cc_E2"), float(g["cc_scale"])).cpu().numpy()
    np.testing.assert_allclose(K, g["cc_K"], rtol=3e-7, atol=0)
    assert (K == g["cc_K"]).mean() > 0.9
    # non-contiguous views are accepted like the reference; float64 path runs
    x1 = c("cc_x1")
    K2 = CB.cross_covariance(x1[:, ::2], c("cc_E1")[:, ::2], c("cc_x2"), c("cc_E2"), 0.83).cpu().numpy()
    np.testing.assert_allclose(K2, g["cc_K"][:, ::2], rtol=3e-7)
    K3 = CB.cross_covariance(x1.double(), c("cc_E1").double(), c("cc_x2").double(), c("cc_E2").double(), 0.83)
    np.testing.assert_allclose(K3.cpu().numpy(), g["cc_K"], rtol=2e-5, atol=1e-9)  # fp64 result vs the fp32 golden
    with pytest.raises(RuntimeError, match="same device"):
        CB.cross_covariance(x1.cpu(), c("cc_E1"), c("cc_x2"), c("cc_E2"), 0.83)


def test_cross_covariance_large_random_vs_ref_backend():
    rb = DO.ref_backends()
    if rb is None:
        pytest.skip("oracle/_ref not built")
    from como_b200 import como_backends as CB

    g = torch.Generator().manual_seed(3)
    n1, n2 = 64, 20000
    x1, x2 = torch.rand(1, n1, 2, generator=g) * 2 - 1, torch.rand(1, n2, 2, generator=g) * 2 - 1

    def E(n):
        a = 1e-3 + 5e-2 * torch.rand(1, n, generator=g)
        c = 1e-3 + 5e-2 * torch.rand(1, n, generator=g)
        o = torch.sqrt(a * c) * 0.9 * (2 * torch.rand(1, n, generator=g) - 1)
        return torch.stack((a, o, o, c), -1).reshape(1, n, 2, 2)

    E1, E2 = E(n1), E(n2)
    ref = rb.cross_covariance(x1, E1, x2, E2, 1.0).numpy()
    out = CB.cross_covariance(x1.cuda(), E1.cuda(), x2.cuda(), E2.cuda(), 1.0).cpu().numpy()
    np.testing.assert_allclose(out, ref, rtol=4e-7, atol=1e-30)
    assert (out == ref).mean() > 0.9


def test_get_new_chol_obs_info_sequence(golden_dir):
    from como_b200 import como_backends as CB

    g = G(golden_dir)
    c = lambda k: torch.from_numpy(g[k]).cuda()
    xd, Ed, sel = c("ca_xd"), c("ca_Ed"), list(g["ca_sel"])
    L, obs, var = c("ca_L0").clone(), c("ca_obs0").clone(), c("ca_var0").clone()
    for i in range(1, len(sel)):
        k_ni = CB.cross_covariance(xd[:, sel[:i]], Ed[:, sel[:i]], xd[:, sel[i:i + 1]], Ed[:, sel[i:i + 1]], 1.0)
        k_id = CB.cross_covariance(xd[:, sel[i:i + 1]], Ed[:, sel[i:i + 1]], xd, Ed, 1.0)
        CB.get_new_chol_obs_info(L, obs, var, k_ni, k_id, 1.0, i)
    np.testing.assert_allclose(L.cpu().numpy(), g["ca_L"], rtol=2e-5, atol=1e-6)
    np.testing.assert_allclose(obs.cpu().numpy(), g["ca_obs"], rtol=1e-4, atol=2e-6)
    np.testing.assert_allclose(var.cpu().numpy(), g["ca_var"], rtol=1e-4, atol=2e-6)
    with pytest.raises(RuntimeError, match="contiguous"):
        CB.get_new_chol_obs_info(L, obs.transpose(1, 2), var, k_ni, k_id, 1.0, 1)


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_sampler_indices_bit_exact_vs_reference(golden_dir, tag):
    from como_b200.depth_cov.core.samplers import sample_sparse_coords

    g = G(golden_dir)
    cov = torch.from_numpy(g[f"s{tag}_cov"]).cuda()
    n = int(g[f"s{tag}_n"])
    coords, inds = sample_sparse_coords(cov, n, "greedy_conditional_entropy", max_stdev_thresh=1e-2, border=3,
                                        dist_thresh=0.1, signal_var=torch.tensor(1.0), fixed_var=0.0)
    np.testing.assert_array_equal(inds.cpu().numpy(), g[f"s{tag}_inds"])
    np.testing.assert_array_equal(coords.cpu().numpy(), g[f"s{tag}_coords"])
    keep = torch.from_numpy(g[f"s{tag}_coords"]).cuda()[:, : n // 2].float()
    c2, i2 = sample_sparse_coords(cov, n, "greedy_conditional_entropy", max_stdev_thresh=1e-2, border=3, dist_thresh=0.1,
                                  signal_var=torch.tensor(1.0), fixed_var=0.0, curr_coords=keep)
    np.testing.assert_array_equal(i2.cpu().numpy(), g[f"s{tag}_inds2"])


def test_sampler_full_resolution_vs_oracle():
    """640x480 domain (d = 300k): anchor indices identical to the oracle (reference backend on the CPU)."""
    from como_b200.depth_cov.core.samplers import sample_sparse_coords
    from como_b200 import synth

    cov = synth.make_cov_image(480, 640, seed=5)
    co, io = DO.sample_sparse_coords(cov, 64, max_stdev_thresh=1e-2, border=3, dist_thresh=0.1, signal_var=1.0, fixed_var=0.0)
    cg, ig = sample_sparse_coords(cov.cuda(), 64, "greedy_conditional_entropy", max_stdev_thresh=1e-2, border=3,
                                  dist_thresh=0.1, signal_var=torch.tensor(1.0), fixed_var=0.0)
    np.testing.assert_array_equal(ig.cpu().numpy(), io.numpy())
    # property: all anchors further apart than dist_thresh in normalised coordinates
    xy = (2 * cg[0].cpu().double() + 1) / torch.tensor([480.0, 640.0]) - 1
    dmin = torch.cdist(xy, xy) + 10 * torch.eye(xy.shape[0])
    assert float(dmin.min()) > 0.1


@pytest.mark.parametrize("name", ["ba_k4_notfull"])
def test_prep_predictor_vs_reference(golden_dir, name):
    from como_b200.depth_cov.core.predictor import prep_predictor

    g = np.load(os.path.join(golden_dir, name + ".npz"))
    cov = torch.from_numpy(g["in_cov_params_img"]).cuda()
    pm = torch.from_numpy(g["in_pm_first_obs"]).cuda()
    coords_m = torch.stack((pm[..., 1], pm[..., 0]), -1)
    Kinv, L, KK = prep_predictor(cov, coords_m, float(g["gp_scale"]))
    rel = lambda a, b: float((a.cpu() - torch.from_numpy(b)).abs().max() / np.abs(b).max())
    assert rel(L, g["in_L_mm"]) < 1e-7
    assert rel(Kinv, g["in_Kmm_inv"]) < 1e-5
    assert rel(KK, g["in_Knm_Kmminv"]) < 1e-6
    assert float((L[0].cpu() - torch.from_numpy(g["in_L_mm"][0])).abs().max()) < 1e-9
