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


@pytest.mark.parametrize("n", [1, 63, 64, 130, 449, 2848])
def test_both_schedules_solve_the_same_system(n):
    """The chain schedule (one CTA walks the critical path, default) and the round-1 dataflow schedule factorise the
    same tiles with the same per-tile arithmetic; only the order of the last subtraction in each diagonal /
    sub-diagonal tile differs.  Both must solve the system to the same accuracy, and each must be repeatable."""
    from como_b200 import _lib
    from como_b200.odom.mapping_core import solve_system

    H, b = _spd(n, 100 + n)
    Hd, bd = H.cuda(), b.cuda()
    xs = {}
    try:
        for mode in (0, 1):
            _lib.chol_schedule(mode)
            xs[mode] = solve_system(Hd, bd).clone()
            assert torch.equal(xs[mode], solve_system(Hd, bd))      # run-to-run bitwise repeatable
    finally:
        _lib.chol_schedule(1)
    scale = float(xs[0].abs().max())
    np.testing.assert_allclose(xs[1].cpu().numpy(), xs[0].cpu().numpy(), rtol=0, atol=1e-9 * scale)
    r = (H @ xs[1].cpu() - b[:, None]).abs().max() / b.abs().max()
    assert float(r) < 1e-10
