"""Times the K-matrix/predictor kernels at 640x480, M=64 (one keyframe) with CUDA events."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from como_b200 import _lib, synth
from como_b200.depth_cov.core.predictor import prep_predictor
from como_b200.depth_cov.core import distill_depth as DD

dev = "cuda:0"
H, W, M = 480, 640, 64
cov = synth.make_cov_image_wide(H, W, seed=0).double().to(dev)
torch.manual_seed(0)
rr, cc = torch.meshgrid(torch.arange(8), torch.arange(8), indexing="ij")
cm = torch.stack(((rr.reshape(-1) + 0.5) * H / 8, (cc.reshape(-1) + 0.5) * W / 8), -1)[None].double().to(dev)


def timeit(f, n=10):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


Kinv, L, KK = prep_predictor(cov, cm, 0.09)
E_m = torch.empty(1, M, 4, dtype=torch.float64, device=dev)
K_mm = torch.empty(1, M, M, dtype=torch.float64, device=dev)
out = torch.empty(1, H, W, M, dtype=torch.float64, device=dev)
s = _lib.stream_ptr(dev)
t = timeit(lambda: _lib.kmat_predictor(_lib.ptr(cov), 1, H, W, _lib.ptr(cm), _lib.ptr(E_m), _lib.ptr(Kinv), M, 0.09, _lib.ptr(out), s))
flops = 2.0 * H * W * M * M
print(f"kmat_predictor (grid) {t*1e3:.1f} us  GEMM part {flops/t*1e-9:.2f} TFLOP/s  write {H*W*M*8/t*1e-6:.0f} GB/s")
print("whole prep_predictor %.1f us" % (timeit(lambda: prep_predictor(cov, cm, 0.09)) * 1e3))
n = H * W
coords_n = torch.stack((torch.rand(n) * (H - 1), torch.rand(n) * (W - 1)), -1)[None].double().to(dev)
mask = torch.ones(n, dtype=torch.uint8, device=dev)
t = timeit(lambda: DD.predictor_rows(cm, coords_n, mask, cov, 0.09, True))
print(f"predictor_rows (fractional, var) {t*1e3:.1f} us")
rows, _, var, vmin = DD.predictor_rows(cm, coords_n, mask, cov, 0.09, True)
y = torch.randn(n, dtype=torch.float64, device=dev)
t = timeit(lambda: DD._gram(rows, y, var, mask, 1e-3, 1.0))
print(f"weighted_gram {t*1e3:.1f} us  ({2.0*n*M*M/t*1e-9:.2f} TFLOP/s full-square equivalent)")
x = torch.randn(1, M, 1, dtype=torch.float64, device=dev)
res = torch.empty(n, dtype=torch.float64, device=dev)
st3 = torch.empty(3, dtype=torch.float64, device=dev)
t = timeit(lambda: _lib.rows_residual(_lib.ptr(rows), _lib.ptr(x), _lib.ptr(y), _lib.ptr(mask), n, M, _lib.ptr(res), _lib.ptr(st3), s))
print(f"rows_residual {t*1e3:.1f} us  {n*M*8/t*1e-6:.0f} GB/s")
