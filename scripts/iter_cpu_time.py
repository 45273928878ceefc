"""How much host time does one iterate() take (launch overhead) vs. device time?"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from como_b200 import synth
from como_b200.odom import mapping_core as MC
s = synth.make_ba_window(32, 24, 480, 640, M=64, device="cuda", seed=0)
cfg = synth.ba_cfg()
for _ in range(5):
    MC.iterate(s, cfg)
torch.cuda.synchronize()
t0 = time.time()
for _ in range(20):
    MC.iterate(s, cfg)
t1 = time.time()
torch.cuda.synchronize()
t2 = time.time()
print(f"host time per iterate {1e3*(t1-t0)/20:.2f} ms; wall incl. drain {1e3*(t2-t0)/20:.2f} ms")
