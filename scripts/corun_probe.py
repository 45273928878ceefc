"""How do the solve and the store_vars stream behave on part of the chip, alone and side by side?"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from como_b200 import synth, _lib
from como_b200.odom import mapping_core as MC
dev = torch.device("cuda", 0)
s = synth.make_ba_window(32, 24, 480, 640, M=64, device=dev, seed=0)
cfg = synth.ba_cfg()
for _ in range(3):
    MC.iterate(s, cfg)
L = MC.kernel_launchers(s, cfg, dev)
cache = s.__dict__["_b200_cache"]
K, H, W, M = 32, 480, 640, 64
scaf = cache["scaf"]
depth = torch.empty(K, 1, H, W, dtype=torch.float64, device=dev)

def stream(ctas):
    _lib.predictor_stream_ctas(ctas)
    _lib.check(_lib.predictor_apply(_lib.ptr(s.Knm_Kmminv), _lib.ptr(scaf), K, H * W, M, _lib.ptr(depth), _lib.stream_ptr(dev)), "pa")

def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

for c in (0, 128, 112, 96, 80, 64):
    _lib.chol_ctas(c)
    print(f"solve alone, {c or 148} CTAs: {timeit(L['solve']):.3f} ms")
_lib.chol_ctas(0)
for c in (0, 148, 96, 80, 64, 48):
    print(f"stream alone, {c or 296} CTAs: {timeit(lambda: stream(c)):.3f} ms")
side = torch.cuda.Stream(dev)
def both(cc, sc):
    _lib.chol_ctas(cc)
    ev = torch.cuda.Event(); ev.record(); side.wait_event(ev)
    with torch.cuda.stream(side):
        stream(sc)
    L["solve"]()
    ev2 = torch.cuda.Event(); ev2.record(side); torch.cuda.current_stream().wait_event(ev2)
for cc, sc in ((0, 0), (100, 96), (96, 104), (84, 128), (112, 72), (74, 148)):
    print(f"solve {cc or 148} CTAs || stream {sc or 296} CTAs: {timeit(lambda: both(cc, sc)):.3f} ms")
_lib.chol_ctas(0); _lib.predictor_stream_ctas(0)
