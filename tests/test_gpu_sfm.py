"""Two-frame SfM bootstrap on the GPU, through the C ABI, vs the reference golden and the oracle."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _load(golden_dir):
    g = np.load(os.path.join(golden_dir, "sfm_64x48.npz"), allow_pickle=True)
    L = int(g["levels"])
    t = lambda k: torch.from_numpy(g[k]).to(DEV)
    init = {k[5:]: g[k].item() for k in g.files if k.startswith("init_")}
    args = dict(coords=[t(f"l{l}_test_coords") for l in range(L)], vals=[t(f"l{l}_vals") for l in range(L)],
                Knm=[t(f"l{l}_Knm_Kmminv") for l in range(L)], imgs=[t(f"l{l}_img_and_grads_j") for l in range(L)],
                Ks=[t(f"l{l}_intrinsics") for l in range(L)])
    return g, L, init, args


def test_first_normal_equations_vs_reference(golden_dir):
    """Level by level: one iteration from the reference's state at the start of that level reproduces H0, g0."""
    from como_b200.odom.frontend import two_frame_sfm as SF
    g, L, init, a = _load(golden_dir)
    one = dict(init, max_iter=1)
    T = torch.from_numpy(g["T_init"]).to(DEV)
    d = torch.from_numpy(g["sparse_log_depth_init"]).to(DEV)
    aff = torch.zeros(1, 2, 1, dtype=torch.float64, device=DEV)
    for l in range(L):
        out = SF.two_frame_sfm(T, d, aff, a["coords"][l], a["vals"][l], a["Knm"][l], a["imgs"][l],
                               torch.from_numpy(g["dr_prior_dd"]).to(DEV), torch.from_numpy(g["H_prior_d_d"]).to(DEV),
                               a["Ks"][l], {"photo": 0.1}, None, one)
        # after one iteration: T1 = T0 exp(delta0), d1 = d0 + delta0[6:]
        d0 = torch.from_numpy(g[f"l{l}_delta0"]).to(DEV)
        np.testing.assert_allclose((out[1] - d).reshape(-1).cpu().numpy(), d0[6:, 0].cpu().numpy(), rtol=0,
                                   atol=1e-7 * float(d0.abs().max()))
        T, d = torch.from_numpy(g[f"l{l}_T"]).to(DEV), torch.from_numpy(g[f"l{l}_sparse_log_depth"]).to(DEV)


def test_two_frame_sfm_pyr_vs_reference(golden_dir):
    from como_b200.odom.frontend import two_frame_sfm as SF
    g, L, init, a = _load(golden_dir)
    aff = torch.zeros(1, 2, 1, dtype=torch.float64, device=DEV)
    T, d, _, cj, dj, mld = SF.two_frame_sfm_pyr(
        torch.from_numpy(g["T_init"]).to(DEV), torch.from_numpy(g["sparse_log_depth_init"]).to(DEV), aff, a["coords"], a["vals"],
        a["Knm"], a["imgs"], torch.from_numpy(g["dr_prior_dd"]).to(DEV), torch.from_numpy(g["H_prior_d_d"]).to(DEV), a["Ks"],
        {"photo": 0.1}, None, init)
    assert SF.two_frame_sfm_pyr.last_iters == [int(g[f"l{l}_iters"]) for l in range(L)]
    np.testing.assert_allclose(T.cpu().numpy(), g["T_final"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(d.cpu().numpy(), g["sparse_log_depth_final"], rtol=0, atol=1e-5)
    assert dj.shape[1] == int(g[f"l{L - 1}_num_valid"]) and cj.shape[1] == dj.shape[1]
    np.testing.assert_allclose(np.sort(dj.cpu().numpy().reshape(-1)), g["depth_final_sorted"], rtol=1e-5)
    np.testing.assert_allclose(float(mld), float(np.asarray(g["mean_log_depth_final"]).reshape(-1)[0]), atol=1e-6)


def test_single_iteration_blocks_vs_oracle():
    """sfm_linearize + sfm_accumulate at 160x120, M = 40 against the oracle's linearisation (random pose / depths)."""
    import ctypes as C
    from como_b200 import _lib, synth
    from oracle import sfm_oracle as SO
    torch.manual_seed(5)
    H, W, M = 120, 160, 40
    N = H * W
    img = synth.make_rgb(H, W, seed=3, dtype=torch.float64)
    gray = 0.2989 * img[:, 0:1] + 0.587 * img[:, 1:2] + 0.114 * img[:, 2:3]
    gx, gy = synth._scharr(gray)
    iag = torch.cat((gray, gx, gy), 1)
    Knm = torch.rand(N, M, dtype=torch.float64) / M
    d = 0.1 * torch.randn(M, 1, dtype=torch.float64)
    perm = torch.randperm(N)
    rr, cc = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    coords = torch.stack((rr.reshape(-1), cc.reshape(-1)), -1)[perm]
    vals = gray[0, 0, coords[:, 0], coords[:, 1]] + 0.02 * torch.randn(N, dtype=torch.float64)
    Kmat = synth.make_intrinsics(H, W, dtype=torch.float64)
    T = torch.eye(4, dtype=torch.float64)
    T[:3, 3] = torch.tensor([0.02, -0.01, 0.03])
    T[:3, :3] = torch.from_numpy(__import__("scipy.spatial.transform", fromlist=["Rotation"]).Rotation.from_rotvec([0.01, -0.02, 0.015]).as_matrix())
    r, valid, JT, beta, logz, Pj, pj = SO.linearize(T, d, coords, vals, Knm, iag, Kmat)
    rec = torch.empty(N, 8, dtype=torch.float64, device=DEV)
    absr = torch.empty(N, dtype=torch.float64, device=DEV)
    proj = torch.empty(N, 3, dtype=torch.float64, device=DEV)
    stats = torch.empty(2, dtype=torch.float64, device=DEV)
    T12 = (C.c_double * 12)(*T[:3].reshape(-1).tolist())
    intr4 = (C.c_double * 4)(float(Kmat[0, 0]), float(Kmat[1, 1]), float(Kmat[0, 2]), float(Kmat[1, 2]))
    Kd, dd, cd, vd, im = Knm.to(DEV), d.reshape(-1).to(DEV), coords.to(DEV), vals.to(DEV), iag[0].contiguous().to(DEV)
    st = _lib.sfm_linearize(_lib.ptr(Kd), _lib.ptr(dd), _lib.ptr(cd), _lib.ptr(vd), _lib.ptr(im), H, W, N, M, T12, intr4,
                            _lib.ptr(rec), _lib.ptr(absr), _lib.ptr(proj), _lib.ptr(stats), _lib.stream_ptr(torch.device(DEV)))
    assert st == 0
    rc = rec.cpu()
    np.testing.assert_array_equal((~torch.isnan(rc[:, 0])).numpy(), valid.numpy())
    v = valid.numpy()
    np.testing.assert_allclose(rc[:, 0].numpy()[v], r.numpy()[v], rtol=0, atol=1e-12)
    np.testing.assert_allclose(rc[:, 1:7].numpy()[v], JT.numpy()[v], rtol=0, atol=1e-10 * float(JT.abs().max()))
    np.testing.assert_allclose(rc[:, 7].numpy()[v], beta.numpy()[v], rtol=0, atol=1e-10 * float(beta.abs().max()))
    np.testing.assert_allclose(float(stats[0]), float(logz.sum()), rtol=1e-12)
    assert int(stats[1]) == int(valid.sum())
    sigma = 1.4826 * torch.median(r[valid].abs())
    wr = r / sigma
    w = torch.where(wr.abs() < 1.345, torch.ones_like(wr), 1.345 / wr.abs()) * valid
    s = w / sigma ** 2
    G_ref = (Knm * (s * beta * beta)[:, None]).T @ Knm
    St_ref = torch.cat(((JT * (s * beta)[:, None]).T @ Knm, ((s * beta * r)[None]) @ Knm), 0)
    G = torch.empty(M, M, dtype=torch.float64, device=DEV)
    St = torch.empty(7, M, dtype=torch.float64, device=DEV)
    small = torch.empty(28, dtype=torch.float64, device=DEV)
    sg = sigma.reshape(1).to(DEV)
    st = _lib.sfm_accumulate(_lib.ptr(Kd), _lib.ptr(rec), _lib.ptr(sg), N, M, _lib.ptr(G), _lib.ptr(St), _lib.ptr(small),
                             _lib.stream_ptr(torch.device(DEV)))
    assert st == 0
    np.testing.assert_allclose(G.cpu().numpy(), G_ref.numpy(), rtol=0, atol=1e-11 * float(G_ref.abs().max()))
    np.testing.assert_allclose(St.cpu().numpy(), St_ref.numpy(), rtol=0, atol=1e-11 * float(St_ref.abs().max()))
    HTT = (JT * s[:, None]).T @ JT
    iu = np.triu_indices(6)
    np.testing.assert_allclose(small.cpu().numpy()[:21], HTT.numpy()[iu], rtol=0, atol=1e-11 * float(HTT.abs().max()))
    np.testing.assert_allclose(small.cpu().numpy()[21:27], ((JT * (s * r)[:, None]).sum(0)).numpy(), rtol=0,
                               atol=1e-11 * float((JT * (s * r)[:, None]).sum(0).abs().max()))
    np.testing.assert_allclose(float(small[27]), float((w * wr * wr).sum()), rtol=1e-11)


def test_setup_reference_vs_reference(golden_dir):
    """Predictor pyramids, values, intrinsics pyramid and prior linearisation of setup_reference; the reference lists
    the pixels in a random order, so rows are matched through their (row, col)."""
    from como_b200.odom.frontend import two_frame_sfm as SF
    g, L, init, a = _load(golden_dir)
    H, W = int(g["H"]), int(g["W"])
    iag = [torch.from_numpy(g[f"l{l}_img_and_grads_ref"]).to(DEV) for l in range(L)]
    cm = torch.from_numpy(g["coords_m"]).to(DEV)
    dims = torch.tensor([H, W], dtype=torch.float64, device=DEV)
    cm_norm = 2.0 * (1.0 / dims) * cm + (1.0 / dims) - 1.0
    f = 525.0 * W / 640
    K = torch.tensor([[f, 0, W / 2], [0, f, H / 2], [0, 0, 1]], dtype=torch.float64, device=DEV)
    vals, coords, Knm, sizes, Kp, dr, Hp = SF.setup_reference(iag, cm_norm, float(g["gp_scale"]), torch.from_numpy(g["cov_params_img"]).to(DEV), K)
    np.testing.assert_allclose(dr.cpu().numpy(), g["dr_prior_dd"], rtol=0, atol=1e-7 * np.abs(g["dr_prior_dd"]).max())
    np.testing.assert_allclose(Hp.cpu().numpy(), g["H_prior_d_d"], rtol=0, atol=1e-7 * np.abs(g["H_prior_d_d"]).max())
    for l in range(L):
        np.testing.assert_allclose(Kp[l].cpu().numpy(), g[f"l{l}_intrinsics"], rtol=1e-12)
        tc = g[f"l{l}_test_coords"][0]
        w = sizes[l][1]
        idx = tc[:, 0] * w + tc[:, 1]
        np.testing.assert_allclose(vals[l].cpu().numpy()[0, 0][idx], g[f"l{l}_vals"][0, 0], rtol=0, atol=1e-15)
        ref = g[f"l{l}_Knm_Kmminv"][0]
        np.testing.assert_allclose(Knm[l].cpu().numpy()[0][idx], ref, rtol=0, atol=2e-6 * np.abs(ref).max())
