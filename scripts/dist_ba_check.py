"""Multi-GPU check (run under torchrun, one rank per GPU): a BA window whose pair blocks are sharded by reference
keyframe over the ranks (exact global median through all-reduced digit histograms + NCCL all-reduce of H, g, err)
must give the same iteration as the single-GPU path."""
import copy
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from como_b200 import synth  # noqa: E402
from como_b200.odom import mapping_core as MC  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    K, R, H, W, M = 8, 6, 192, 256, 32
    s = synth.make_ba_window(K, R, H, W, M=M, device=dev, seed=11, ndrop=10)
    cfg = synth.ba_cfg()
    s_single = copy.copy(s)
    for k, v in s.__dict__.items():
        if isinstance(v, torch.Tensor):
            setattr(s_single, k, v.clone())
    s_single.__dict__.pop("_b200_cache", None)

    comm = MC.ShardComm(world, rank, dev)
    ok = True
    for it in range(3):
        d1 = MC.iterate(s_single, cfg, return_debug=True)
        d2 = MC.iterate(s, cfg, comm=comm, return_debug=True)
        rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
        res = dict(sigma=rel(d2["sigma"], d1["sigma"]), H=rel(d2["H"], d1["H"]), g=rel(d2["g"], d1["g"]),
                   delta=rel(d2["delta"], d1["delta"]), poses=rel(s.kf_poses, s_single.kf_poses),
                   P_m=rel(s.P_m, s_single.P_m), med=rel(s.median_depths, s_single.median_depths), err=abs(float(d2["err"].sum() - d1["err"].sum())) / float(d1["err"].sum()))
        bad = (res["sigma"] > 1e-9 or res["H"] > 1e-9 or res["g"] > 1e-8 or res["poses"] > 1e-6 or res["P_m"] > 1e-6
               or res["med"] > 1e-9)
        ok = ok and not bad
        if rank == 0:
            print(f"iter {it}: " + " ".join(f"{k}={v:.2e}" for k, v in res.items()), "BAD" if bad else "ok", flush=True)
    t = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("DIST_BA_CHECK", "PASS" if float(t) > 0 else "FAIL", "world", world, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
