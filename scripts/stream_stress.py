"""Characterise the store_vars corruption seen when the predictor stream co-runs with other kernels."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from como_b200 import _lib
torch.manual_seed(0)
dev = torch.device("cuda", 0)
K, HW, M = 8, 307200, 64
Knm = torch.randn(K, HW, M, dtype=torch.float64, device=dev) * 0.05
scaf = torch.zeros(K, M, 16, dtype=torch.float64, device=dev)
scaf[:, :, 0] = torch.randn(K, M, dtype=torch.float64, device=dev)
ref = torch.empty(K, HW, dtype=torch.float64, device=dev)
_lib.predictor_stream_ctas(0)
_lib.check(_lib.predictor_apply(_lib.ptr(Knm), _lib.ptr(scaf), K, HW, M, _lib.ptr(ref), _lib.stream_ptr()), "pa")
torch.cuda.synchronize()
side = torch.cuda.Stream(dev)
A = torch.randn(4096, 4096, dtype=torch.float64, device=dev)
big = torch.randn(64 * 1024 * 1024, dtype=torch.float32, device=dev)

def load(kind):
    if kind == "matmul":
        return A @ A
    if kind == "memset":
        big.zero_(); big.add_(1.0); return big
    if kind == "median":
        v = torch.rand(4, 300000, dtype=torch.float64, device=dev)
        off = (torch.arange(5, dtype=torch.int64) * 300000).to(dev)
        out = torch.empty(4, dtype=torch.float64, device=dev)
        ws = torch.empty(int(_lib.median_workspace_bytes(4, 8)), dtype=torch.uint8, device=dev)
        for _ in range(3):
            _lib.check(_lib.median_f64(_lib.ptr(v), _lib.ptr(off), 4, 300000, 1.0, _lib.ptr(out), None, _lib.ptr(ws), ws.numel(), _lib.stream_ptr()), "m")
        return out
    return None

for ctas in (0, 148, 96):
    for kind in ("none", "matmul", "memset", "median"):
        bad = 0
        rows = set()
        for rep in range(6):
            out = torch.full((K, HW), float("nan"), dtype=torch.float64, device=dev)
            torch.cuda.synchronize()
            ev = torch.cuda.Event(); ev.record()
            side.wait_event(ev)
            with torch.cuda.stream(side):
                _lib.predictor_stream_ctas(ctas)
                _lib.check(_lib.predictor_apply(_lib.ptr(Knm), _lib.ptr(scaf), K, HW, M, _lib.ptr(out), _lib.stream_ptr(dev)), "pa")
            keep = [load(kind) for _ in range(3)]
            torch.cuda.synchronize()
            d = (out != ref) & ~(torch.isnan(out) & torch.isnan(ref))
            nb = int(d.sum())
            bad += nb
            if nb:
                idx = torch.nonzero(d)[:6].tolist()
                rows.update((i[0], i[1] // 64, i[1] % 64) for i in idx)
        print(f"ctas {ctas:3d} co-running {kind:7s}: differing pixels over 6 reps = {bad}", sorted(rows)[:6])
_lib.predictor_stream_ctas(0)
