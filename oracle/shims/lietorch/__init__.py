"""TEST INFRASTRUCTURE ONLY -- stand-in for the (absent, unpinned) `lietorch` dependency.

The reference only calls ``lietorch.SE3.exp(x).matrix()`` (como/geometry/lie_algebra.py:45-56).
lietorch is a third-party fork (github.com/edexheim/lietorch, unpinned in install.sh:10-13) that
is not vendored under /root/reference, so its published algorithm is restated here:
tangent order [tau(3), phi(3)], R = Exp(phi) (Rodrigues), t = V(phi) tau with V the SO(3) left
Jacobian, Taylor branches for small angles.  Parity for this dependency is therefore "unpinned";
the restatement is cross-checked against scipy in tests/test_oracle_lie.py.
"""
import torch


class _Mat:
    def __init__(self, T):
        self._T = T

    def matrix(self):
        return self._T


class SE3:
    @staticmethod
    def exp(x):
        tau, phi = x[..., :3], x[..., 3:]
        th2 = (phi * phi).sum(-1)
        th = th2.sqrt()
        small = th2 < 1e-12
        ths = torch.where(small, torch.ones_like(th), th)
        A = torch.where(small, 1 - th2 / 6, torch.sin(ths) / ths)
        B = torch.where(small, 0.5 - th2 / 24, (1 - torch.cos(ths)) / (ths * ths))
        C = torch.where(small, 1.0 / 6 - th2 / 120, (ths - torch.sin(ths)) / (ths ** 3))
        n = x.shape[0]
        W = torch.zeros(n, 3, 3, dtype=x.dtype, device=x.device)
        W[:, 0, 1] = -phi[:, 2]
        W[:, 0, 2] = phi[:, 1]
        W[:, 1, 0] = phi[:, 2]
        W[:, 1, 2] = -phi[:, 0]
        W[:, 2, 0] = -phi[:, 1]
        W[:, 2, 1] = phi[:, 0]
        WW = W @ W
        I = torch.eye(3, dtype=x.dtype, device=x.device).expand(n, 3, 3)
        R = I + A[:, None, None] * W + B[:, None, None] * WW
        V = I + B[:, None, None] * W + C[:, None, None] * WW
        T = torch.zeros(n, 4, 4, dtype=x.dtype, device=x.device)
        T[:, :3, :3] = R
        T[:, :3, 3] = (V @ tau[..., None])[..., 0]
        T[:, 3, 3] = 1
        return _Mat(T)
