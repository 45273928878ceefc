import ctypes as C, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from como_b200 import _lib
g = torch.Generator().manual_seed(0)
A = torch.randn(64, 80, generator=g, dtype=torch.float64)
H = (A @ A.T / 64 + 1e-2 * torch.eye(64, dtype=torch.float64))
f = _lib.lib.como_b200_chol_debug_probe
f.argtypes = [C.c_void_p, C.c_void_p]
for rep in range(2):
    t = H.clone().cuda(); clk = torch.zeros(40, dtype=torch.int64, device="cuda")
    f(C.c_void_p(t.data_ptr()), C.c_void_p(clk.data_ptr()))
c = clk.cpu().tolist()
Linv = torch.linalg.inv(torch.linalg.cholesky(H))
print("inverse max err", float((t.cpu() - Linv).abs().max()))
for jb in range(8):
    a = c[4*jb+1]-c[4*jb]; b = c[4*jb+2]-c[4*jb+1]; cc = c[4*jb+3]-c[4*jb+2]
    print(f"jb {jb}: (a) {a}  (b) {b}  (c) {cc} clk")
print("loop total", c[32]-c[0], " inv8", c[33]-c[32], " levels", c[34]-c[33], c[35]-c[34], c[36]-c[35], " all", c[36]-c[0])
