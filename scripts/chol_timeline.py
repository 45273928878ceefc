"""Debug: per-tile timeline of the tiled Cholesky (globaltimer stamps) -> where does the critical path go?"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from como_b200 import _lib
from como_b200.odom.mapping_core import solve_system

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2848
nb = (n + 63) // 64
ntiles = nb * (nb + 1) // 2 + nb
g = torch.Generator().manual_seed(0)
A = torch.randn(n, n + 8, generator=g, dtype=torch.float64)
H = (A @ A.T / n + 1e-3 * torch.eye(n, dtype=torch.float64)).cuda()
b = torch.randn(n, generator=g, dtype=torch.float64).cuda()
for _ in range(3):
    solve_system(H, b)
torch.cuda.synchronize()
buf = torch.zeros(ntiles * 4, dtype=torch.int64, device="cuda")
f = _lib.lib.como_b200_chol_debug_timeline
f.argtypes = [C.c_void_p]
f(C.c_void_p(buf.data_ptr()))
solve_system(H, b)
torch.cuda.synchronize()
f(C.c_void_p(0))
t = buf.cpu().numpy().reshape(ntiles, 4).astype(np.int64)
tiles = [(i, k) for k in range(nb) for i in range(k, nb + 1)]
t0 = t[:, 0].min()
t = np.where(t > 0, t - t0, 0)
print("total span us", (t[:, 3].max()) / 1e3)
diag = {k: idx for idx, (i, k) in enumerate(tiles) if i == k}
prev_end = 0
print(" k  start   upd_done  end   (potrf us)  gap_from_prev_diag_end")
for k in range(nb):
    s, u, _, e = t[diag[k]]
    if k % 4 == 0 or k == nb - 1:
        print(f"{k:2d} {s/1e3:8.1f} {u/1e3:8.1f} {e/1e3:8.1f}  potrf {(e-u)/1e3:6.1f}  upd {(u-s)/1e3:6.1f}  since prev diag {(e-prev_end)/1e3:6.1f}")
    prev_end = e
# subdiagonal tile (k+1,k)
print("subdiag tiles: start, upd_done, flag_kk_seen, end")
for k in range(0, nb - 1, 6):
    idx = tiles.index((k + 1, k))
    s, u, w, e = t[idx]
    print(f"({k+1},{k}) {s/1e3:8.1f} {u/1e3:8.1f} {w/1e3:8.1f} {e/1e3:8.1f}   diag end {t[diag[k]][3]/1e3:8.1f}")
# per-step cost in throughput regime: tiles with many updates
dur = (t[:, 1] - t[:, 0]) / 1e3
ks = np.array([k for (i, k) in tiles])
for kk in (10, 20, 30, 40):
    sel = (ks == kk)
    print(f"column {kk}: update phase mean {dur[sel].mean():.1f} us = {dur[sel].mean()/kk:.2f} us/step; trsm phase {((t[sel,3]-t[sel,2])/1e3).mean():.1f} us")
