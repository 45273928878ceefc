"""Tiled dataflow Cholesky solve (csrc/chol.cu) through the C ABI vs torch.linalg (fp64)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _spd(n, seed, cond_boost=0.0):
    g = torch.Generator().manual_seed(seed)
    A = torch.randn(n, n + 8, generator=g, dtype=torch.float64)
    H = A @ A.T / n + (1e-3 + cond_boost) * torch.eye(n, dtype=torch.float64)
    b = torch.randn(n, generator=g, dtype=torch.float64)
    return H, b


@pytest.mark.parametrize("n", [1, 7, 64, 65, 129, 200, 449, 1000, 2848])
def test_chol_solve_matches_torch(n):
    from como_b200.odom.mapping_core import solve_system
    H, b = _spd(n, n)
    Hd, bd = H.cuda(), b.cuda()
    Hkeep = Hd.clone()
    x = solve_system(Hd, bd)
    assert x.shape == (n, 1)
    assert torch.equal(Hd, Hkeep)                      # H is not modified
    L = torch.linalg.cholesky(H)
    ref = torch.cholesky_solve(b[:, None], L)
    scale = float(ref.abs().max())
    np.testing.assert_allclose(x.cpu().numpy(), ref.numpy(), rtol=0, atol=1e-9 * scale)
    # residual check in fp64
    r = (H @ x.cpu() - b[:, None]).abs().max() / b.abs().max()
    assert float(r) < 1e-10


def test_chol_solve_upper_triangle_ignored_and_repeatable():
    from como_b200.odom.mapping_core import solve_system
    H, b = _spd(300, 5)
    Hd = H.cuda()
    x0 = solve_system(Hd, b.cuda())
    Hd2 = torch.tril(Hd) + torch.triu(torch.full_like(Hd, 7.0), diagonal=1)   # garbage above the diagonal
    x1 = solve_system(Hd2, b.cuda())
    assert torch.equal(x0, x1)
    x2 = solve_system(Hd, b.cuda())
    assert torch.equal(x0, x2)                          # deterministic (no atomics, fixed tile order)


def test_chol_solve_non_pd_gives_nan_not_error():
    from como_b200.odom.mapping_core import solve_system
    H, b = _spd(130, 9)
    H[70, 70] = -5.0
    x = solve_system(H.cuda(), b.cuda())
    assert not torch.isfinite(x).all()


def test_chol_solve_ba_system(golden_dir):
    """The reference's own normal equations (golden H0, g0 -> delta0)."""
    import os
    from como_b200.odom.mapping_core import solve_system
    g = np.load(os.path.join(golden_dir, "ba_k4_notfull.npz"), allow_pickle=True)
    H, gv, d0 = torch.from_numpy(g["H0"]), torch.from_numpy(g["g0"]), g["delta0"]
    x = solve_system(H.cuda(), gv.reshape(-1).cuda())
    np.testing.assert_allclose(x.cpu().numpy().reshape(-1), d0.reshape(-1), rtol=0, atol=1e-7 * np.abs(d0).max())
