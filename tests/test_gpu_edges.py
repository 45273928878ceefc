"""Edge cases of the C ABI on the GPU: empty / fully masked / ragged inputs, argument errors (no CPU fallback)."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
F64 = torch.float64


def _anchors(H, W, M):
    cols = int(np.ceil(np.sqrt(M * W / H)))
    rows = int(np.ceil(M / cols))
    rr, cc = torch.meshgrid(torch.arange(rows), torch.arange(cols), indexing="ij")
    return torch.stack(((rr.reshape(-1) + 0.5) * H / rows, (cc.reshape(-1) + 0.5) * W / cols), -1)[:M][None].double()


def test_weighted_gram_and_residual_with_everything_masked():
    from como_b200 import _lib
    from como_b200.depth_cov.core import distill_depth as DD
    n, m = 1000, 24
    rows = torch.randn(n, m, dtype=F64, device="cuda")
    y = torch.randn(n, dtype=F64, device="cuda")
    mask = torch.zeros(n, dtype=torch.uint8, device="cuda")
    G, h = DD._gram(rows, y, None, mask, 0.0, 1.0)
    assert float(G.abs().max()) == 0.0 and float(h.abs().max()) == 0.0
    res = torch.empty(n, dtype=F64, device="cuda")
    st3 = torch.empty(3, dtype=F64, device="cuda")
    x = torch.randn(m, dtype=F64, device="cuda")
    st = _lib.rows_residual(_lib.ptr(rows), _lib.ptr(x), _lib.ptr(y), _lib.ptr(mask), n, m, _lib.ptr(res), _lib.ptr(st3),
                            _lib.stream_ptr(rows.device))
    assert st == 0
    assert st3.cpu().tolist() == [0.0, 0.0, 0.0]


def test_weighted_gram_ragged_sizes_match_torch():
    """n not a multiple of the 32-row warp step, m not a multiple of 8: padded tiles must stay zero."""
    from como_b200.depth_cov.core import distill_depth as DD
    for n, m in ((1, 2), (33, 7), (257, 63), (4099, 50)):
        g = torch.Generator().manual_seed(n)
        rows = torch.randn(n, m, dtype=F64, generator=g)
        y = torch.randn(n, dtype=F64, generator=g)
        G, h = DD._gram(rows.cuda(), y.cuda(), None, None, 0.0, 0.25)
        np.testing.assert_allclose(G.cpu().numpy(), (0.25 * rows.T @ rows).numpy(), rtol=1e-11, atol=1e-11)
        np.testing.assert_allclose(h.cpu().numpy(), (0.25 * rows.T @ y).numpy(), rtol=1e-11, atol=1e-11)


def test_kmat_rows_zero_points_and_all_points_masked():
    from como_b200 import synth
    from como_b200.depth_cov.core import distill_depth as DD
    H, W, M = 48, 64, 16
    cov = synth.make_cov_image_wide(H, W, seed=1).cuda()
    cm = _anchors(H, W, M).cuda()
    rows, L, var, vmin = DD.predictor_rows(cm, torch.empty(1, 0, 2, dtype=F64, device="cuda"), None, cov, 0.1, True)
    assert rows.shape == (0, M) and var.shape == (0,)
    cn = torch.rand(1, 100, 2, dtype=F64, device="cuda") * 40
    mask = torch.zeros(100, dtype=torch.uint8, device="cuda")
    rows, L, var, vmin = DD.predictor_rows(cm, cn, mask, cov, 0.1, True)
    assert float(rows.abs().max()) == 0.0
    assert float(vmin) == 1e300      # "no valid point" sentinel: nothing took part in the minimum


def test_kmat_predictor_small_anchor_counts_vs_oracle():
    """M = 4 and M = 12 (padded DMMA tiles, fewer than 32 anchors per lane group)."""
    from como_b200 import synth
    from como_b200.depth_cov.core.predictor import prep_predictor
    from oracle import depthcov_oracle as DO
    H, W = 24, 40
    cov = synth.make_cov_image_wide(H, W, seed=3)
    for M in (4, 12):
        cm = _anchors(H, W, M)
        Kinv_o, L_o, KK_o = DO.prep_predictor(cov, cm, 0.09)
        Kinv, L, KK = prep_predictor(cov.cuda(), cm.cuda(), 0.09)
        np.testing.assert_allclose(L.cpu().numpy(), L_o.numpy(), rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(KK.cpu().numpy(), KK_o.numpy(), rtol=0, atol=1e-8 * float(KK_o.abs().max()))


def test_reproject_dense_mask_matches_reference_filter():
    from como_b200 import _lib
    H, W = 30, 40
    z = (1.0 + torch.rand(H, W, dtype=F64)).cuda()
    z[3, 4] = -1.0                                   # behind the camera after the (identity) motion
    T12 = (C.c_double * 12)(1, 0, 0, 0.05, 0, 1, 0, -0.02, 0, 0, 1, 0.0)
    intr = (C.c_double * 4)(30.0, 30.0, W / 2, H / 2)
    cj = torch.empty(H * W, 2, dtype=F64, device="cuda")
    lz = torch.empty(H * W, dtype=F64, device="cuda")
    zj = torch.empty(H * W, dtype=F64, device="cuda")
    mk = torch.empty(H * W, dtype=torch.uint8, device="cuda")
    st = _lib.reproject_dense(_lib.ptr(z), H, W, T12, intr, 0.0, _lib.ptr(cj), _lib.ptr(lz), _lib.ptr(zj), _lib.ptr(mk),
                              _lib.stream_ptr(z.device))
    assert st == 0
    rr, cc = torch.meshgrid(torch.arange(H, dtype=F64), torch.arange(W, dtype=F64), indexing="ij")
    zc = z.cpu()
    X = (cc - W / 2) / 30.0 * zc + 0.05
    Y = (rr - H / 2) / 30.0 * zc - 0.02
    u, v = 30.0 * X / zc + W / 2, 30.0 * Y / zc + H / 2
    ok = (u >= 1) & (u < W - 1) & (v >= 1) & (v < H - 1) & (zc > 0.0)
    np.testing.assert_array_equal(mk.cpu().numpy().reshape(H, W).astype(bool), ok.numpy())
    np.testing.assert_allclose(cj.cpu().numpy().reshape(H, W, 2)[..., 1][ok.numpy()], u.numpy()[ok.numpy()], rtol=1e-13)
    assert not bool(mk.cpu().reshape(H, W)[3, 4])


def test_abi_rejects_bad_arguments_with_message():
    from como_b200 import _lib
    st = _lib.chol_solve(None, None, 4, None, None, 0, None)
    assert st == -1 and b"null pointer" in _lib.last_error()
    x = torch.zeros(4, dtype=F64, device="cuda")
    st = _lib.chol_solve(_lib.ptr(x), _lib.ptr(x), 4, _lib.ptr(x), _lib.ptr(x), 8, None)
    assert st == -3 and b"workspace" in _lib.last_error()
    with pytest.raises(RuntimeError, match="same device"):
        from como_b200.odom.mapping_core import solve_system
        solve_system(torch.eye(3, dtype=F64), torch.ones(3, dtype=F64))


def test_ba_iterate_without_one_way_frames_vs_oracle():
    """R = 0: keyframe pairs only (the reference's state right after initialisation)."""
    from como_b200 import synth
    from como_b200.odom import mapping_core as MC
    from oracle import ba_oracle as BO
    s = synth.make_ba_window(4, 0, 48, 64, M=16, device="cuda", seed=4, ndrop=4)
    cfg = synth.ba_cfg()
    sc = {}
    for k, v in s.__dict__.items():
        if not k.startswith("_"):
            sc[k] = v.detach().cpu().clone() if isinstance(v, torch.Tensor) else (list(v) if isinstance(v, list) else v)
    o = BO.iterate(sc, cfg)
    dbg = MC.iterate(s, cfg, return_debug=True)

    def rel(a, b):
        a, b = torch.as_tensor(a).detach().cpu().double(), torch.as_tensor(b).detach().cpu().double()
        return float((a - b).abs().max() / b.abs().max().clamp_min(1e-300))

    assert rel(dbg["H"], o["H"]) < 1e-6
    assert rel(dbg["g"], o["g"]) < 1e-6
    assert rel(s.kf_poses, sc["kf_poses"]) < 1e-5
    assert rel(s.P_m, sc["P_m"]) < 1e-5


def test_img_and_grads_f64_vs_torch():
    """Fused gray + Scharr (Mapping.get_img_and_grads) against the plain torch ops of the reference, odd sizes."""
    from como_b200.odom.mapping_core import get_img_and_grads
    for (H, W) in ((2, 2), (31, 47), (96, 128)):
        g = torch.Generator().manual_seed(H)
        rgb = torch.rand(1, 3, H, W, dtype=F64, generator=g)
        gray = 0.2989 * rgb[:, 0:1] + 0.587 * rgb[:, 1:2] + 0.114 * rgb[:, 2:3]
        kx = torch.tensor([[-3.0, 0, 3], [-10, 0, 10], [-3, 0, 3]], dtype=F64) / 32.0
        ky = kx.t().contiguous()
        xp = torch.nn.functional.pad(gray, (1, 1, 1, 1), mode="reflect")
        ref = torch.cat((gray, torch.nn.functional.conv2d(xp, kx[None, None]), torch.nn.functional.conv2d(xp, ky[None, None])), 1)
        out = get_img_and_grads(rgb.cuda())
        np.testing.assert_allclose(out.cpu().numpy(), ref.numpy(), rtol=0, atol=1e-14)
