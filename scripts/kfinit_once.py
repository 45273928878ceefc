"""One warm track_and_init call between cudaProfilerStart/Stop (launch list for profiles/)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from como_b200.odom.frontend.corr import track_and_init
c = bench.build_kfinit_case(torch.device("cuda", 0))
def step():
    return track_and_init(c["pose1"], c["pose2"], c["coords_m1"], c["z_m1"], c["z_img1"], c["cov2"], c["K"], c["scale"],
                          bench.KFINIT_CORR, bench.KFINIT_SAMP, (c["H"], c["W"]))
for _ in range(2):
    step()
torch.cuda.synchronize()
print("MARK")
step()
torch.cuda.synchronize()
print("done")
