"""Oracle (CPU restatement of the two-frame SfM bootstrap) vs goldens from the unmodified reference."""
import os

import numpy as np
import pytest
import torch

from oracle import sfm_oracle as SO


def load(golden_dir, name="sfm_64x48"):
    g = np.load(os.path.join(golden_dir, name + ".npz"), allow_pickle=True)
    L = int(g["levels"])
    t = lambda k: torch.from_numpy(g[k])
    coords = [t(f"l{l}_test_coords")[0] for l in range(L)]
    vals = [t(f"l{l}_vals")[0, 0] for l in range(L)]
    Knm = [t(f"l{l}_Knm_Kmminv")[0] for l in range(L)]
    imgs = [t(f"l{l}_img_and_grads_j") for l in range(L)]
    Ks = [t(f"l{l}_intrinsics") for l in range(L)]
    init = {k[5:]: g[k].item() for k in g.files if k.startswith("init_")}
    return g, L, coords, vals, Knm, imgs, Ks, init


@pytest.mark.parametrize("name", ["sfm_64x48", "sfm_96x72"])
def test_two_frame_sfm_oracle_vs_reference(golden_dir, name):
    g, L, coords, vals, Knm, imgs, Ks, init = load(golden_dir, name)
    traces = []
    T, d, mld, pjv, zv, iters = SO.two_frame_sfm_pyr(
        torch.from_numpy(g["T_init"])[0], torch.from_numpy(g["sparse_log_depth_init"])[0], coords, vals, Knm, imgs, Ks,
        torch.from_numpy(g["dr_prior_dd"])[0], torch.from_numpy(g["H_prior_d_d"])[0], init, traces)
    for l in range(L):
        tr = traces[l]
        H0, g0 = g[f"l{l}_H0"], g[f"l{l}_g0"]
        np.testing.assert_allclose(tr["H0"].numpy(), H0, rtol=0, atol=1e-9 * np.abs(H0).max())
        np.testing.assert_allclose(tr["g0"].numpy(), g0, rtol=0, atol=1e-9 * np.abs(g0).max())
        assert tr["iters"] == int(g[f"l{l}_iters"])
        np.testing.assert_allclose(tr["T"].numpy(), g[f"l{l}_T"][0], rtol=0, atol=1e-7)
        np.testing.assert_allclose(tr["d"].numpy(), g[f"l{l}_sparse_log_depth"][0], rtol=0, atol=1e-6)
    np.testing.assert_allclose(T.numpy(), g["T_final"][0], rtol=0, atol=1e-7)
    np.testing.assert_allclose(d.numpy(), g["sparse_log_depth_final"][0], rtol=0, atol=1e-6)
    assert zv.numel() == int(g[f"l{L - 1}_num_valid"])
    np.testing.assert_allclose(np.sort(zv.numpy()), g["depth_final_sorted"], rtol=1e-6)
    np.testing.assert_allclose(float(mld), float(np.asarray(g["mean_log_depth_final"]).reshape(-1)[0]), atol=1e-7)
