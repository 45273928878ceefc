"""Is the first iteration's store_vars result independent of how the side stream is scheduled?"""
import os, sys, copy
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from como_b200 import synth, _lib
from como_b200.odom import mapping_core as MC

def clone_state(s):
    return MC.WindowState(**{k: (v.clone() if isinstance(v, torch.Tensor) else (list(v) if isinstance(v, list) else v))
                             for k, v in s.__dict__.items() if not k.startswith("_")})

s0 = synth.make_ba_window(32, 24, 480, 640, M=64, seed=0)
cfg = synth.ba_cfg()
res = {}
for name, ctas, overlap in (("two_per_sm", 0, True), ("one_per_sm", -1, True), ("96", 96, True), ("serial", 0, False),
                            ("one_per_sm_again", -1, True)):
    MC._STREAM_CTAS, MC._OVERLAP = ctas, overlap
    s = clone_state(s0)
    s.Knm_Kmminv = s0.Knm_Kmminv   # share the 5 GB slab
    MC.iterate(s, cfg)
    torch.cuda.synchronize()
    res[name] = (s.median_depths.clone(), s.depth_imgs.clone(), s.kf_poses.clone())
    a = res["two_per_sm"]
    print(name, "medians bit-equal:", bool(torch.equal(res[name][0], a[0])), "depth bit-equal:", bool(torch.equal(res[name][1], a[1])),
          "max |d med|", float((res[name][0] - a[0]).abs().max()), "n depth diff", int((res[name][1] != a[1]).sum()))
