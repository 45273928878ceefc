"""Debug: per-column timeline of the critical CTA of chol_factor2_kernel (globaltimer stamps, ns)."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from como_b200 import _lib
from como_b200.odom.mapping_core import solve_system
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2848
nb = (n + 63) // 64
g = torch.Generator().manual_seed(0)
A = torch.randn(n, n + 8, generator=g, dtype=torch.float64)
H = (A @ A.T / n + 1e-3 * torch.eye(n, dtype=torch.float64)).cuda()
b = torch.randn(n, generator=g, dtype=torch.float64).cuda()
for _ in range(3):
    solve_system(H, b)
torch.cuda.synchronize()
buf = torch.zeros(nb * 8 + 4096, dtype=torch.int64, device="cuda")
f = _lib.lib.como_b200_chol_debug_timeline
f.argtypes = [C.c_void_p]
f(C.c_void_p(buf.data_ptr()))
solve_system(H, b)
torch.cuda.synchronize()
f(C.c_void_p(0))
t = buf.cpu().numpy()[: nb * 8].reshape(nb, 8).astype(np.int64)
t0 = t[0, 0]
print("columns", nb, "span us", (t[-1, 6] - t0) / 1e3)
d = lambda a, b: (t[:, b] - t[:, a]) / 1e3
names = [("form T (own + syrk)", 0, 1), ("potrf64", 1, 2), ("wait P_S", 2, 3), ("stores + fence + flag + fetch", 3, 4), ("trsm64", 4, 5), ("store + fence + flag", 5, 6)]
for nm, a, bb in names:
    x = d(a, bb)
    print(f"{nm:32s} mean {x.mean():6.2f} us  max {x.max():6.2f}  sum {x.sum():7.1f}")
print("per column mean", ((t[:, 6] - t[:, 0]) / 1e3).mean())
