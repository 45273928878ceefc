"""Debug helper: B independent 640x480 tracking problems through the plan and the stats path, one sync after each."""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from como_b200.odom.frontend.photo_tracking import TrackBatchPlan, photo_tracking_pyr_batch

B = int(sys.argv[1]) if len(sys.argv) > 1 else 230
dev = torch.device("cuda:0")
probs, T0, a0, cases = bench.build_track_problems(B, dev)
plan = TrackBatchPlan(probs, bench.TERM)
torch.cuda.synchronize(); print("plan built", flush=True)
T, aff, nit = plan.run(T0, a0)
torch.cuda.synchronize(); print("plan.run ok", int(nit.sum()), flush=True)
T2, aff2, stats, nit2 = photo_tracking_pyr_batch(T0, a0, probs, bench.TERM, return_stats=True)
torch.cuda.synchronize(); print("batch ok", int(nit2.sum()), bool(torch.equal(T, T2)), flush=True)
