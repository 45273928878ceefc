"""Keyframe-creation path on the GPU, through the C ABI, vs the reference goldens and the oracle."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _anchors(H, W, M, seed):
    """M well separated anchors (jittered grid): random anchors can nearly coincide, and K_mm^-1 then amplifies
    the CPU/GPU factorisation differences far beyond the kernels' own rounding."""
    g = torch.Generator().manual_seed(seed)
    cols = int(np.ceil(np.sqrt(M * W / H)))
    rows = int(np.ceil(M / cols))
    rr, cc = torch.meshgrid(torch.arange(rows), torch.arange(cols), indexing="ij")
    base = torch.stack(((rr.reshape(-1) + 0.5) * H / rows, (cc.reshape(-1) + 0.5) * W / cols), -1)[:M]
    return (base + (torch.rand(M, 2, generator=g) - 0.5) * 0.3 * min(H / rows, W / cols))[None].double()


def _case(g, ci, dev):
    pre = f"c{ci}_"
    ins = [torch.from_numpy(g[pre + k]).to(dev) for k in ("pose1", "pose2", "coords_m1", "z_m1", "z_img1", "cov_params_img2", "K")]
    corr = {k[5:]: g[k].item() for k in g.files if k.startswith("corr_")}
    samp = {k[5:]: g[k].item() for k in g.files if k.startswith("samp_")}
    return pre, ins, corr, samp


@pytest.mark.parametrize("ci", [0, 1, 2])
def test_track_and_init_vs_reference(golden_dir, ci):
    from como_b200.odom.frontend.corr import track_and_init
    g = np.load(os.path.join(golden_dir, "kfinit_64x48.npz"), allow_pickle=True)
    pre, ins, corr, samp = _case(g, ci, "cuda:0")
    dbg = {}
    c2, z2, mask, call, zall = track_and_init(*ins, float(g["gp_scale"]), corr, samp, tuple(g[pre + "rgb_img_size"]), debug=dbg)
    np.testing.assert_allclose(dbg["dd_coords_m"].cpu().numpy(), g[pre + "dd_coords_m"], rtol=1e-12)
    assert dbg["dd_n"] == int(g[pre + "dd_n"])
    np.testing.assert_allclose(dbg["dd_logz_m"].cpu().numpy(), g[pre + "dd_logz_m"], rtol=0, atol=1e-7)
    np.testing.assert_allclose(dbg["dd_res_std"], float(g[pre + "dd_res_std"]), rtol=1e-6)
    np.testing.assert_array_equal(dbg["ss0_inds"].cpu().numpy(), g[pre + "ss0_inds"])
    np.testing.assert_array_equal(dbg["ss1_inds"].cpu().numpy(), g[pre + "ss1_inds"])
    np.testing.assert_array_equal(mask.cpu().numpy(), g[pre + "corr_mask"])
    np.testing.assert_array_equal(c2.cpu().numpy(), g[pre + "coords_2"])
    np.testing.assert_allclose(call.cpu().numpy(), g[pre + "coords_all"], rtol=1e-12)
    np.testing.assert_allclose(dbg["dc_logz_2"].cpu().numpy(), g[pre + "dc_logz_2"], rtol=0, atol=1e-7)
    np.testing.assert_allclose(zall.cpu().numpy(), g[pre + "z_all"], rtol=1e-7)


def test_distill_dropins_vs_oracle():
    """The reference-shaped entry points (compacted inputs) against the oracle on a synthetic 160x120 case with
    64 anchors, invalid observations and out-of-image points."""
    from como_b200 import synth
    from como_b200.depth_cov.core import distill_depth as DD
    from oracle import kfinit_oracle as KO
    torch.manual_seed(3)
    H, W, M = 120, 160, 64
    cov = synth.make_cov_image_wide(H, W, seed=5).double()
    n = 7000
    coords_n = torch.stack((torch.rand(n) * (H + 6) - 3, torch.rand(n) * (W + 6) - 3), -1)[None].double()
    coords_m = _anchors(H, W, M, 1)
    z = (1.5 + torch.rand(1, n, 1)).double()
    z[0, ::17, 0] = 0.0
    scale = 0.09
    lo, ro = KO.distill_depth_from_scratch(coords_m, coords_n, z, cov, scale, True, 0.0)
    lg, rg = DD.distill_depth_from_scratch(coords_m.cuda(), coords_n.cuda(), z.cuda(), cov.cuda(), scale, True, 0.0)
    np.testing.assert_allclose(lg.cpu().numpy(), lo.numpy(), rtol=0, atol=2e-7)
    np.testing.assert_allclose(rg.cpu().numpy(), ro.numpy(), rtol=0, atol=2e-7)
    lo2, _ = KO.distill_depth_from_scratch(coords_m, coords_n, z, cov, scale, False, 0.0)
    lg2, _ = DD.distill_depth_from_scratch(coords_m.cuda(), coords_n.cuda(), z.cuda(), cov.cuda(), scale, False, 0.0)
    np.testing.assert_allclose(lg2.cpu().numpy(), lo2.numpy(), rtol=0, atol=2e-5)
    z1 = torch.exp(lo[:, :40])
    co = KO.distill_conditional_from_scratch(coords_m, z1, coords_n, cov, z, scale, 0.0, 0.03)
    cg = DD.distill_conditional_depth_from_scratch(coords_m.cuda(), z1.cuda(), coords_n.cuda(), cov.cuda(), z.cuda(), scale, 0.0, 0.03)
    np.testing.assert_allclose(cg.cpu().numpy(), co.numpy(), rtol=0, atol=2e-7)


def test_kmat_rows_variance_and_gram_vs_torch():
    """kmat_rows (fractional coords, mask, variance, min) and weighted_gram against plain torch fp64 of the oracle
    K-matrices; M = 52 exercises the padded DMMA tiles."""
    from como_b200 import synth
    from como_b200.depth_cov.core import distill_depth as DD
    from oracle import kfinit_oracle as KO
    torch.manual_seed(4)
    H, W, M, n = 96, 128, 52, 3001
    cov = synth.make_cov_image_wide(H, W, seed=2).double()
    coords_n = torch.stack((torch.rand(n) * (H - 1), torch.rand(n) * (W - 1)), -1)[None].double()
    coords_m = _anchors(H, W, M, 2)
    mask = (torch.rand(n) > 0.2).to(torch.uint8)
    K_mm, K_nm, K_nn = KO.kernel_matrices(coords_m, coords_n, cov, 0.11)
    L, _ = torch.linalg.cholesky_ex(K_mm)
    KK = K_nm @ torch.cholesky_solve(torch.eye(M, dtype=torch.float64)[None], L)
    var = (K_nn - torch.sum(K_nm * KK, dim=2))[0]
    rows, Lg, var_g, vmin = DD.predictor_rows(coords_m.cuda(), coords_n.cuda(), mask.cuda(), cov.cuda(), 0.11, True)
    mb = mask.bool()
    scl = float(KK.abs().max())
    np.testing.assert_allclose(rows.cpu().numpy()[mb.numpy()], KK[0].numpy()[mb.numpy()], rtol=0, atol=2e-6 * scl)
    assert float(rows.cpu()[~mb].abs().max()) == 0.0
    np.testing.assert_allclose(var_g.cpu().numpy()[mb.numpy()], var.numpy()[mb.numpy()], rtol=0, atol=1e-7)
    np.testing.assert_allclose(float(vmin), float(var[mb].min()), rtol=0, atol=1e-7)
    y = torch.randn(n).double()
    w = 1.0 / (var_g.cpu() + 0.37)
    w[~mb] = 0.0
    R = rows.cpu()
    G_ref = (R * w[:, None]).T @ R
    h_ref = (R * (w * y)[:, None]).sum(0)
    G, h = DD._gram(rows, y.cuda(), var_g, mask.cuda(), 0.37, 1.0)
    np.testing.assert_allclose(G.cpu().numpy(), G_ref.numpy(), rtol=1e-11, atol=1e-11 * float(G_ref.abs().max()))
    np.testing.assert_allclose(h.cpu().numpy(), h_ref.numpy(), rtol=1e-11, atol=1e-11 * float(h_ref.abs().max()))
