"""Debug helper (GPU box): where does the CUDA BA iteration differ from the oracle on a golden state?"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import ba_oracle as BO  # noqa
from como_b200.odom import mapping_core as MC  # noqa
from test_gpu_ba import cuda_state  # noqa

name = sys.argv[1] if len(sys.argv) > 1 else "ba_k4_full"
g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
cfg = BO.cfg_from_golden(g)
s = cuda_state(BO.state_from_golden(g))
so = BO.state_from_golden(g)
o = BO.iterate(so, cfg)
dbg = MC.iterate(s, cfg, return_debug=True)
print("sigma cuda", dbg["sigma"].cpu().numpy(), "oracle", o["sigmas"])
Hc, Ho = dbg["H_photo"].cpu(), o["H_photo"]
K = s.kf_poses.shape[0]
R = s.recent_poses.shape[0]
nb = K + R
print("block rel diffs (pose-pose 8x8 blocks):")
for a in range(nb):
    row = []
    for b in range(nb):
        A, B = Hc[8 * a:8 * a + 8, 8 * b:8 * b + 8], Ho[8 * a:8 * a + 8, 8 * b:8 * b + 8]
        row.append(float((A - B).abs().max() / (B.abs().max() + 1e-300)) if float(B.abs().max()) > 0 else float(A.abs().max()))
    print(a, " ".join("%.1e" % v for v in row))
lm0 = 8 * nb
print("pose-landmark block diff", float((Hc[:lm0, lm0:] - Ho[:lm0, lm0:]).abs().max() / Ho[:lm0, lm0:].abs().max()))
print("landmark block diff", float((Hc[lm0:, lm0:] - Ho[lm0:, lm0:]).abs().max() / Ho[lm0:, lm0:].abs().max()))
print("g diff", float((dbg["g_photo"].cpu() - o["g_photo"]).abs().max() / o["g_photo"].abs().max()))
print("err", float(dbg["err"][0]), o["photo_err"])
cache = s.__dict__["_b200_cache"]
pp = cache["pair_plan"]
kp = cache["kf_plan"]
rb = pp.rbuf.view(pp.P, kp.N).cpu()
print("pairs", pp.pair_ref.cpu().tolist(), pp.pair_tgt.cpu().tolist())
print("valid per pair (cuda):", [(~torch.isnan(rb[p])).sum().item() for p in range(pp.P)])
scaf = cache["scaf"].cpu()
print("logzm diff", float((scaf[:, :, 0] - so["logzm"][:, :, 0]).abs().max()), "pm diff", float((scaf[:, :, 2:4] - so["pm"]).abs().max()))
print("P_m diff after", float((s.P_m.cpu() - so["P_m"]).abs().max()))
refz = pp.refbuf.view(K, kp.N, 8).cpu()
print("z_n range", float(refz[..., 0].min()), float(refz[..., 0].max()))
print("Pwn diff", "n/a")
