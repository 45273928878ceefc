"""predictor_apply must give bit-identical depth images for every grid size (como_b200_predictor_stream_ctas)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from como_b200 import _lib
torch.manual_seed(0)
K, HW, M = 5, 307200, 64
Knm = torch.randn(K, HW, M, dtype=torch.float64, device="cuda") * 0.05
scaf = torch.zeros(K, M, 16, dtype=torch.float64, device="cuda")
scaf[:, :, 0] = torch.randn(K, M, dtype=torch.float64, device="cuda")
ref = torch.exp((Knm @ scaf[:, :, 0:1])[..., 0])
outs = {}
for c in (0, 296, 148, 96, 37, 1):
    _lib.predictor_stream_ctas(c)
    out = torch.full((K, HW), float("nan"), dtype=torch.float64, device="cuda")
    _lib.check(_lib.predictor_apply(_lib.ptr(Knm), _lib.ptr(scaf), K, HW, M, _lib.ptr(out), _lib.stream_ptr()), "pa")
    torch.cuda.synchronize()
    outs[c] = out
    print("ctas", c, "max rel err vs torch", float(((out - ref).abs() / ref).max()), "nan", int(torch.isnan(out).sum()),
          "bit-equal to ctas 0:", bool(torch.equal(out, outs[0])))
_lib.predictor_stream_ctas(0)
