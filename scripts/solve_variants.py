"""GPU experiment: timing of dense SPD solve variants at the BA system size (n ~ 2848, fp64)."""
import sys
import time

import torch

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2848
torch.manual_seed(0)
A = torch.randn(n, n, dtype=torch.float64, device="cuda")
H = A @ A.T + n * torch.eye(n, dtype=torch.float64, device="cuda")
g = torch.randn(n, dtype=torch.float64, device="cuda")


def timeit(name, fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name:55s} {e0.elapsed_time(e1) / reps * 1e3:9.1f} us")
    return out


def cur():
    L, _ = torch.linalg.cholesky_ex(H, upper=False, check_errors=False)
    return torch.cholesky_solve(g[:, None], L, upper=False)


def upper():
    U, _ = torch.linalg.cholesky_ex(H, upper=True, check_errors=False)
    return torch.cholesky_solve(g[:, None], U, upper=True)


def tri(ncol):
    def f():
        L, _ = torch.linalg.cholesky_ex(H, upper=False, check_errors=False)
        B = g[:, None].expand(n, ncol).contiguous()
        y = torch.linalg.solve_triangular(L, B, upper=False)
        return torch.linalg.solve_triangular(L.mT, y, upper=True)[:, :1]
    return f


x0 = timeit("cholesky_ex(lower) + cholesky_solve", cur)
timeit("cholesky_ex(upper) + cholesky_solve", upper)
timeit("cholesky_ex only", lambda: torch.linalg.cholesky_ex(H, upper=False, check_errors=False))
L, _ = torch.linalg.cholesky_ex(H, upper=False, check_errors=False)
timeit("cholesky_solve only (1 rhs)", lambda: torch.cholesky_solve(g[:, None], L, upper=False))
for nc in (1, 8, 32):
    x = timeit(f"chol + 2x solve_triangular ({nc} cols)", tri(nc))
    print("   max diff vs current", float((x - x0).abs().max()))
B8 = g[:, None].expand(n, 8).contiguous()
timeit("solve_triangular lower only (8 cols)", lambda: torch.linalg.solve_triangular(L, B8, upper=False))
timeit("solve_triangular lower only (1 col)", lambda: torch.linalg.solve_triangular(L, g[:, None], upper=False))
timeit("linalg.solve (LU)", lambda: torch.linalg.solve(H, g))
try:
    torch.backends.cuda.preferred_linalg_library("magma")
    timeit("magma: cholesky_ex + cholesky_solve", cur)
except Exception as e:
    print("magma unavailable", e)
torch.backends.cuda.preferred_linalg_library("cusolver")
Hs = H.clone()
timeit("H.clone()", lambda: H.clone())
timeit("H.zero_()", lambda: Hs.zero_())
