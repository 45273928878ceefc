#!/usr/bin/env python
"""bench.py -- measures the COMO photometric Gauss-Newton hot path on B200.

Contract (driver): `python bench.py --gpus N --steps K --warmup W [--impl reference]`, launched under
torchrun for N>1 (one rank per GPU).  Prints ONE JSON line on rank 0.

Workloads (config.workload):
  ba_window (default) one Gauss-Newton iteration of a 32-keyframe 640x480 window (BASELINE.json metric).
  kf_init   one keyframe creation (track_and_init) at 640x480 with 64 anchors.
  track640  B independent 640x480 / 4-level frame-to-keyframe tracking problems per GPU, one cooperative
            launch per step; metric = Gauss-Newton iterations per second (sum over problems).
A "step" is one pass of the hot path over one batch of synthetic input.  The batch (B x 21 MB of
reference operands + B target pyramids) is larger than the 126 MB L2, so no L2 flush is needed.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

TERM = dict(max_iter=50, delta_norm=1e-3, rel_tol=1e-3, grad_norm=1.0)  # config/como.yml:12-17
TRACK_BYTES_PER_PX_ITER = 52  # P 12 + I_ref 4 + J 32 + target 4


# --------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append((time.time(), ln.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ts, ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                clk, mxv = float(f[1]), float(f[2])
            except ValueError:
                continue
            mx = mxv
            if t0 <= ts <= t1 + 0.2:
                sm.append(clk)
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        if not sm:  # region shorter than the sampling period: take everything we saw
            for ts, ln in self.lines:
                f = [x.strip() for x in ln.split(",")]
                try:
                    sm.append(float(f[1]))
                except Exception:
                    pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------------------------- workloads
def build_track_problems(B, device, seed0=0):
    from como_b200 import synth

    probs, cases = [], []
    for b in range(B):
        case = synth.make_tracking_case(480, 640, 4, seed=seed0 + b, device=device)
        probs.append((case["vals"], case["P"], case["dI_dT"], case["mask"], case["K"], case["img"]))
        cases.append(case)
    T0 = torch.cat([c["T_init"] for c in cases], 0)
    a0 = torch.cat([c["aff_init"] for c in cases], 0)
    return probs, T0, a0, cases


def run_track_ours(args, rank, world, device):
    from como_b200 import synth
    from como_b200.odom.frontend.photo_tracking import TrackBatchPlan, photo_tracking_pyr_batch

    B = args.batch
    probs, T0, a0, cases = build_track_problems(B, device, seed0=rank * B)
    npx = [int(m.sum()) for m in cases[0]["mask"]]
    plan = TrackBatchPlan(probs, TERM)   # descriptors built once; a step = copy initial poses + one launch

    def step():
        return plan.run(T0, a0)

    for _ in range(args.warmup):
        T, aff, nit = step()
    torch.cuda.synchronize()
    barrier(world)
    clk = ClockSampler(torch.cuda.current_device() if "CUDA_VISIBLE_DEVICES" not in os.environ else 0)
    if rank == 0:
        clk.start()
        time.sleep(0.3)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters_total = torch.zeros((), dtype=torch.int64, device=device)
    t_wall0 = time.time()
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(args.steps):
        T, aff, nit = step()
        iters_total += nit.sum()
    ev1.record()
    torch.cuda.synchronize()
    t_wall1 = time.time()
    barrier(world)
    ms = ev0.elapsed_time(ev1)
    ms_max = allreduce_max(ms, world, device)
    its = int(allreduce_sum(int(iters_total.item()), world, device))
    clocks = clk.stop(t_wall0, t_wall1) if rank == 0 else None

    # algorithmic bytes of the timed region on this rank: per problem, sum over iterations of n_level*52
    # (stats give the level of each iteration) -- measured once outside the timed loop
    Ts, affs, stats, nit = photo_tracking_pyr_batch(T0, a0, probs, TERM, return_stats=True)
    torch.cuda.synchronize()
    stc = stats.cpu()
    alg_bytes = 0
    for b in range(B):
        nb = int(nit[b])
        lv = stc[b, :nb, 0].long().tolist()
        msk = [int(m.sum()) for m in cases[b]["mask"]]
        alg_bytes += sum(msk[l] * TRACK_BYTES_PER_PX_ITER for l in lv)
    kernel_ms = ms / args.steps  # the step IS one kernel launch (+ a 4 KB descriptor memcpy)
    peak, peak_src = measured_peaks()
    ach = alg_bytes / (kernel_ms * 1e-3) / 1e9

    if args.no_e2e:
        return {"value": its / (ms_max * 1e-3), "ms_per_step": ms_max / args.steps, "n_gpus": world,
                "config": {"workload": "track640", "problems_per_gpu": B, "tuning_only": True},
                "roofline": {"achieved": ach, "peak": peak, "frac": ach / peak, "alg_bytes_per_launch": alg_bytes,
                             "launch_ms": kernel_ms}}, cases
    # e2e through the public API: B `Tracking` objects (one per sequence), each step every tracker handles one new
    # frame: pinned host RGB -> device, gray pyramid, tracking, reprojection statistics, keyframe decision (the
    # reference's handle_frame), pose read back to the host.
    import como_b200.odom.frontend.photo_tracking as PT
    from como_b200.odom.Tracking import Tracking

    tcfg = {"device": str(device), "dtype": "float", "color": "gray",
            "pyr": {"start_level": 0, "end_level": 4, "depth_interp_mode": "nearest_neighbor"},
            "term_criteria": dict(TERM), "sigmas": {"photo": 1.0e-1},
            "keyframing": {"kf_depth_motion_ratio": 0.12, "kf_num_pixels_frac": 0.75, "one_way_freq": 3}}
    trackers = []
    for b in range(B):
        tr = Tracking(tcfg, cases[b]["K0"].cpu(), (480, 640))
        tr.setup()
        tr.update_kf_reference(([1.0], cases[b]["rgb"], torch.eye(4, device=device)[None],
                                torch.zeros(1, 2, 1, device=device), cases[b]["depth"]))
        trackers.append(tr)
    rgb_host = [c["rgb2"].cpu().pin_memory() for c in cases]  # (1,3,480,640) fp32 each
    out_host = torch.empty(B, 16, dtype=torch.float32).pin_memory()

    def e2e_step():
        tot = torch.zeros((), dtype=torch.int64, device=device)
        for b, tr in enumerate(trackers):
            tr.T_curr_kf = cases[b]["T_init"].clone()
            tr.aff_curr_kf = cases[b]["aff_init"].clone()
            rgb = rgb_host[b].to(device, non_blocking=True)
            viz, _ = tr.handle_frame((2.0, rgb))
            tot += PT.last_num_iters[0]
            out_host[b].copy_(viz[1].reshape(16), non_blocking=True)
        return tot

    for _ in range(max(1, args.warmup // 2)):
        e2e_step()
    torch.cuda.synchronize()
    barrier(world)
    e_it = torch.zeros((), dtype=torch.int64, device=device)
    ev0.record()
    for _ in range(args.steps):
        e_it += e2e_step()
    ev1.record()
    torch.cuda.synchronize()
    barrier(world)
    e_ms = allreduce_max(ev0.elapsed_time(ev1), world, device)
    e_its = int(allreduce_sum(int(e_it.item()), world, device))
    h2d_bytes = sum(int(r.numel()) * 4 for r in rgb_host)

    res = {
        "metric": "GN-iterations/sec at 640x480 (tracking, 4-level pyramid)", "value": its / (ms_max * 1e-3),
        "unit": "GN-it/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "track640", "resolution": "640x480", "pyramid_levels": 4, "problems_per_gpu": B,
                   "px_per_level": npx, "l2": "inputs (B x 21 MB operands) exceed the 126 MB L2; no flush",
                   "parallelism": f"replicas x{world} (independent sequences, no collective)"},
        "e2e": {"value": e_its / (e_ms * 1e-3), "unit": "GN-it/s",
                "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": int(out_host.numel() * 4)},
        "gpu_launches": args.steps * 1,
        "roofline": {"kernel": "track_pyr_kernel", "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s",
                     "frac": ach / peak, "traffic": None, "peak_source": peak_src,
                     "alg_bytes_per_launch": alg_bytes, "launch_ms": kernel_ms},
        "clocks": clocks,
    }
    return res, cases


def cpu_track_baseline(cases, budget_s=15.0):
    """Oracle port of the reference tracker on the host cores (torch intra-op threads)."""
    from oracle import track_oracle as TO

    t0 = time.time()
    its = 0
    n = 0
    for c in cases:
        cc = {k: ([v.cpu() for v in val] if isinstance(val, list) else val) for k, val in c.items()}
        _, _, trace = TO.track_pyr(c["T_init"].cpu(), c["aff_init"].cpu(), cc["vals"], cc["P"], cc["dI_dT"],
                                   cc["mask"], cc["K"], cc["img"], TERM)
        its += len(trace)
        n += 1
        if time.time() - t0 > budget_s:
            break
    dt = time.time() - t0
    return {"value": its / dt, "unit": "GN-it/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{n} of the same 640x480 4-level tracking problems ({its} GN iterations, {dt:.1f} s)"}


# --------------------------------------------------------------------------------------------- BA workload
BA_K, BA_R = 32, 24


def ba_alg_bytes(K, N, P, HW, M):
    """Algorithmic bytes of one BA iteration: predictor apply K*HW*M*8, pair kernels
    K*N*M*8 (one predictor row per reference pixel) + P*N*48."""
    return dict(predictor_apply=K * HW * M * 8, photo=K * N * M * 8 + P * N * 48)


def run_ba_ours(args, rank, world, device):
    from como_b200 import synth
    from como_b200.odom import mapping_core as MC

    K, R, H, W, M = args.kf, args.oneway, 480, 640, 64
    shard = world > 1 and args.shard
    # sharded: every rank holds the SAME window (its pair blocks are split); replicas: one window per rank
    s = synth.make_ba_window(K, R, H, W, M=M, device=device, seed=0 if shard else rank)
    cfg = synth.ba_cfg()
    snap = snapshot_small(s) if rank == 0 else None   # CPU baseline runs on the untouched initial state
    allreduce = None
    hist_allreduce = None
    if shard:
        def allreduce(Hm, g, err):
            torch.distributed.all_reduce(Hm)
            torch.distributed.all_reduce(g)
            torch.distributed.all_reduce(err)

        def hist_allreduce(t):
            torch.distributed.all_reduce(t)

    def step():
        MC.iterate(s, cfg, allreduce=allreduce, hist_allreduce=hist_allreduce, rank=rank if shard else 0,
                   world=world if shard else 1)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    barrier(world)
    clk = ClockSampler(0 if "CUDA_VISIBLE_DEVICES" in os.environ else torch.cuda.current_device())
    if rank == 0:
        clk.start()
        time.sleep(0.3)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()   # `ncu --profile-from-start off` then sees exactly the timed steps
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    t1 = time.time()
    barrier(world)
    ms = allreduce_max(ev0.elapsed_time(ev1), world, device)
    clocks = clk.stop(t0, t1) if rank == 0 else None
    kp, pp = MC.get_plans(s, cfg, device, rank if shard else 0, world if shard else 1)

    # per-kernel device time of the dominant kernel (predictor apply), timed alone with CUDA events
    from como_b200 import _lib
    scaf = s.__dict__["_b200_cache"]["scaf"]
    depth = torch.empty(K, H * W, dtype=torch.float64, device=device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(2):
        _lib.predictor_apply(_lib.ptr(s.Knm_Kmminv), _lib.ptr(scaf), K, H * W, M, _lib.ptr(depth), _lib.stream_ptr(device))
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        _lib.predictor_apply(_lib.ptr(s.Knm_Kmminv), _lib.ptr(scaf), K, H * W, M, _lib.ptr(depth), _lib.stream_ptr(device))
    e1.record()
    torch.cuda.synchronize()
    pa_ms = e0.elapsed_time(e1) / 5
    ab = ba_alg_bytes(K, kp.N, pp.P, H * W, M)
    peak, peak_src = measured_peaks()
    ach = ab["predictor_apply"] / (pa_ms * 1e-3) / 1e9

    # e2e: a new one-way frame arrives on the host every step (pinned RGB) -> device -> gray + gradients ->
    # replaces the oldest one-way frame -> iterate -> poses + error back to the host
    rgb_host = synth.make_rgb(H, W, seed=77, dtype=torch.float64).pin_memory()
    res_host = torch.empty((K + R) * 16 + 1, dtype=torch.float64).pin_memory()

    # The upload of frame i+1 runs on a copy stream while iteration i computes (two device buffers, events both ways):
    # every timed step still issues one full H2D copy of its input and the D2H reads of its result.
    copy_stream = torch.cuda.Stream(device)
    rgb_dev = [torch.empty(rgb_host.shape, dtype=rgb_host.dtype, device=device) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    for ev in consumed:
        ev.record()

    def upload(i):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[i % 2])
            rgb_dev[i % 2].copy_(rgb_host, non_blocking=True)
            ready[i % 2].record(copy_stream)

    def e2e_step(i):
        torch.cuda.current_stream().wait_event(ready[i % 2])
        s.recent_img_and_grads[0].copy_(MC.get_img_and_grads(rgb_dev[i % 2])[0])   # fused gray + Scharr kernel
        consumed[i % 2].record()
        upload(i + 1)
        step()
        res_host[: K * 16].copy_(s.kf_poses.reshape(-1), non_blocking=True)
        res_host[K * 16:(K + R) * 16].copy_(s.recent_poses.reshape(-1), non_blocking=True)
        res_host[-1:].copy_(s.total_err_prev.reshape(1), non_blocking=True)

    upload(0)
    for i in range(2):
        e2e_step(i)
    torch.cuda.synchronize()
    barrier(world)
    ev0.record()
    for i in range(2, 2 + args.steps):
        e2e_step(i)
    ev1.record()
    torch.cuda.synchronize()
    barrier(world)
    e_ms = allreduce_max(ev0.elapsed_time(ev1), world, device)
    nwin = 1 if shard else world   # sharded: one window over all GPUs (strong); else one window per GPU (weak)
    res = {
        "metric": "GN-iterations/sec at 640x480, 32-keyframe window", "value": nwin * args.steps / (ms * 1e-3),
        "unit": "GN-it/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "strong" if shard else "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "ba_window", "resolution": "640x480", "keyframes": K, "one_way_frames": R,
                   "anchors_per_kf": M, "landmarks": kp.L, "pairs": pp.P, "pixels_per_kf": kp.N,
                   "system_dim": 8 * (K + R) + 3 * kp.L,
                   "l2": "inputs (5.0 GB predictor slabs) exceed the 126 MB L2; no flush",
                   "parallelism": (f"pair blocks sharded by reference keyframe over {world} GPUs + NCCL allreduce of H,g"
                                   if shard else f"replicas x{world} (one independent window per GPU)")},
        "e2e": {"value": nwin * args.steps / (e_ms * 1e-3), "unit": "GN-it/s",
                "h2d_bytes_per_step": int(rgb_host.numel() * 8), "d2h_bytes_per_step": int(res_host.numel() * 8)},
        "gpu_launches": args.steps * 36,   # own kernels per step, counted in profiles/r01_launches_ba_final.csv
        "roofline": {"kernel": "predictor_stream_kernel", "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s",
                     "frac": ach / peak,
                     # dram__bytes_read + write of one launch at this exact shape (ncu --set full,
                     # profiles/r01_predictor_stream_full.txt): 5.034 GB + 0.081 GB; other shapes: not captured
                     "traffic": 5115372864 if (K, H, W, M) == (32, 480, 640, 64) else None, "peak_source": peak_src,
                     "alg_bytes_per_launch": ab["predictor_apply"], "launch_ms": pa_ms},
        "clocks": clocks,
    }
    res["finite"] = bool(torch.isfinite(s.kf_poses).all() and torch.isfinite(s.P_m).all())
    res["final_total_err"] = float(s.total_err_prev)
    return res, (s, snap), cfg


# --------------------------------------------------------------------------------------------- keyframe-creation workload
FP64_PEAK_TFLOPS = 37.1   # measured on B200 with scripts/micro/dmma_bench.cu (DFMA == DMMA.8x8x4 == 64 FMA/clk/SM)
KFINIT_CORR = dict(corr_mode="logz", corr_thresh=3.0e-2, distill_with_prior=True, min_obs_depth=0.0,
                   logz_grad_mag_thresh=7.0e-2)                                    # config/como.yml:59-64
KFINIT_SAMP = dict(mode="greedy_conditional_entropy", max_num_coords=64, max_stdev_thresh=1.0e-2, border=3,
                   fixed_var=0.0, dist_thresh=1.0e-1)                              # config/como.yml:52-58


def build_kfinit_case(device, seed=0, H=480, W=640, M=64):
    """Synthetic keyframe-creation inputs (SURVEY 8d scene): depth 2 + 0.5 sin cos, camera translating +x by 6 px,
    covariance image with long length scales, anchors of the last keyframe chosen by the sampler itself."""
    from como_b200 import synth
    from como_b200.depth_cov.core.samplers import sample_sparse_coords

    cov1 = synth.make_cov_image_wide(H, W, seed=seed).to(device)
    cov2 = synth.make_cov_image_wide(H, W, seed=seed + 100).to(device)
    z_img = synth.make_depth(H, W, dtype=torch.float64).to(device).reshape(1, 1, H, W)
    Km = synth.make_intrinsics(H, W, dtype=torch.float64).to(device).reshape(1, 3, 3)
    scale = 0.086
    coords_m1, _ = sample_sparse_coords(cov1, M, "greedy_conditional_entropy", 1e-2, border=3, dist_thresh=0.1,
                                        signal_var=scale, fixed_var=0.0)
    z_m1 = z_img[0, 0, coords_m1[0, :, 0], coords_m1[0, :, 1]].reshape(1, -1, 1)
    pose1 = torch.eye(4, dtype=torch.float64, device=device)[None]
    pose2 = pose1.clone()
    pose2[0, 0, 3] = 6.0 * 2.0 / float(Km[0, 0, 0])
    pose2[0, 1, 3] = 0.0013
    return dict(pose1=pose1, pose2=pose2, coords_m1=coords_m1.double(), z_m1=z_m1, z_img1=z_img, cov2=cov2, K=Km,
                scale=scale, H=H, W=W, M=M)


def run_kfinit_ours(args, rank, world, device):
    from como_b200 import _lib
    from como_b200.depth_cov.core import distill_depth as DD
    from como_b200.odom.frontend.corr import track_and_init

    c = build_kfinit_case(device, seed=rank)
    H, W, M = c["H"], c["W"], c["M"]

    def step(cov2=None, z_img=None):
        return track_and_init(c["pose1"], c["pose2"], c["coords_m1"], c["z_m1"], c["z_img1"] if z_img is None else z_img,
                              c["cov2"] if cov2 is None else cov2, c["K"], c["scale"], KFINIT_CORR, KFINIT_SAMP, (H, W))

    for _ in range(args.warmup):
        out = step()
    torch.cuda.synchronize()
    barrier(world)
    sampler = ClockSampler(device.index)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    ev0.record()
    for _ in range(args.steps):
        out = step()
    ev1.record()
    torch.cuda.synchronize()
    t1 = time.time()
    clocks = sampler.stop(t0, t1)
    barrier(world)
    ms_max = allreduce_max(ev0.elapsed_time(ev1), world, device)

    # end to end: covariance + depth images arrive from pinned host memory, the new anchors go back
    cov_host = c["cov2"].cpu().pin_memory()
    z_host = c["z_img1"].cpu().pin_memory()
    out_host = torch.empty(M * 3 + M, dtype=torch.float64).pin_memory()

    def e2e_step():
        cov2 = cov_host.to(device, non_blocking=True)
        zi = z_host.to(device, non_blocking=True)
        c2, z2, mask, call, zall = step(cov2, zi)
        n = call.shape[1]
        out_host[:2 * n].copy_(call.reshape(-1), non_blocking=True)
        out_host[2 * M:2 * M + n].copy_(zall.reshape(-1), non_blocking=True)
        out_host[3 * M:3 * M + mask.numel()].copy_(mask.to(torch.float64), non_blocking=True)

    for _ in range(2):
        e2e_step()
    torch.cuda.synchronize()
    barrier(world)
    ev0.record()
    for _ in range(args.steps):
        e2e_step()
    ev1.record()
    torch.cuda.synchronize()
    barrier(world)
    e_ms = allreduce_max(ev0.elapsed_time(ev1), world, device)

    # roofline of the dominant kernel (K-matrix / predictor rows with variance), timed alone
    n = H * W
    coords_n = torch.stack((torch.rand(n, device=device) * (H - 1), torch.rand(n, device=device) * (W - 1)), -1)[None].double()
    mask = torch.ones(n, dtype=torch.uint8, device=device)
    for _ in range(3):
        DD.predictor_rows(c["coords_m1"], coords_n, mask, c["cov2"], c["scale"], True)
    E_m = torch.empty(1, M, 4, dtype=torch.float64, device=device)
    K_mm = torch.empty(1, M, M, dtype=torch.float64, device=device)
    st = _lib.stream_ptr(device)
    _lib.kmat_kmm(_lib.ptr(c["cov2"]), 1, H, W, _lib.ptr(c["coords_m1"]), M, c["scale"], 0.0, _lib.ptr(E_m), _lib.ptr(K_mm), st)
    Kinv = torch.linalg.inv(K_mm).contiguous()
    rows = torch.empty(n, M, dtype=torch.float64, device=device)
    var = torch.empty(n, dtype=torch.float64, device=device)
    vmin = torch.empty(1, dtype=torch.float64, device=device)
    reps = 10
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(reps):
        _lib.kmat_rows(_lib.ptr(c["cov2"]), 1, H, W, _lib.ptr(c["coords_m1"]), _lib.ptr(E_m), _lib.ptr(Kinv), M, c["scale"],
                       _lib.ptr(coords_n), _lib.ptr(mask), n, _lib.ptr(rows), _lib.ptr(var), _lib.ptr(vmin), st)
    ev1.record()
    torch.cuda.synchronize()
    kernel_ms = ev0.elapsed_time(ev1) / reps
    alg_flop = 2.0 * n * M * M + 30.0 * n * M   # GEMM 2 n m^2 + ~30 flop per kernel evaluation
    ach = alg_flop / (kernel_ms * 1e-3) * 1e-12
    res = {
        "metric": "keyframe creations/sec at 640x480 (track_and_init, 64 anchors)", "value": world * args.steps / (ms_max * 1e-3),
        "unit": "KF/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "kf_init", "resolution": "640x480", "anchors": M, "dense_points": n,
                   "new_anchors": int(out[0].shape[1]), "correspondences": int(out[2].sum()),
                   "l2": "predictor rows (157 MB per pass, written once and read twice) exceed the 126 MB L2; no flush",
                   "parallelism": f"replicas x{world} (independent keyframes, no collective)"},
        "e2e": {"value": world * args.steps / (e_ms * 1e-3), "unit": "KF/s",
                "h2d_bytes_per_step": int(cov_host.numel() + z_host.numel()) * 8, "d2h_bytes_per_step": int(out_host.numel()) * 8},
        "gpu_launches": args.steps * 253,   # own kernels per call, counted in profiles/r01_launches_kfinit_summary.txt
        "roofline": {"kernel": "kmat_rows_kernel", "bound": "tensor", "achieved": ach, "peak": FP64_PEAK_TFLOPS,
                     "unit": "TFLOP/s", "frac": ach / FP64_PEAK_TFLOPS, "traffic": None,
                     "peak_source": "fp64 DFMA/DMMA peak measured with scripts/micro/dmma_bench.cu (MEASURED_PEAKS.json has no fp64 entry)",
                     "alg_flop_per_launch": alg_flop, "launch_ms": kernel_ms},
        "clocks": clocks,
    }
    return res, c


def cpu_kfinit_baseline(c):
    """Oracle port of track_and_init on the host cores, same inputs."""
    from oracle import kfinit_oracle as KO

    cpu = {k: (v.cpu() if isinstance(v, torch.Tensor) else v) for k, v in c.items()}
    t0 = time.time()
    KO.track_and_init(cpu["pose1"], cpu["pose2"], cpu["coords_m1"], cpu["z_m1"], cpu["z_img1"], cpu["cov2"], cpu["K"],
                      cpu["scale"], KFINIT_CORR, KFINIT_SAMP)
    dt = time.time() - t0
    return {"value": 1.0 / dt, "unit": "KF/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"1 keyframe creation on the same inputs ({dt:.1f} s)"}


def snapshot_small(s):
    """Host copy of everything but the big constant tensors (those are copied lazily by the CPU baseline)."""
    sc = {}
    for k, v in s.__dict__.items():
        if k.startswith("_"):
            continue
        if isinstance(v, torch.Tensor) and v.numel() < (1 << 24):
            sc[k] = v.detach().cpu().clone()
        elif not isinstance(v, torch.Tensor):
            sc[k] = list(v) if isinstance(v, list) else v
    return sc


def cpu_ba_baseline(s_and_snap, cfg, max_iters=1):
    """Oracle port of Mapping.iterate on the host cores (torch CPU, fp64), on the same window (initial state)."""
    from oracle import ba_oracle as BO

    s, sc = s_and_snap
    for k, v in s.__dict__.items():
        if k.startswith("_") or k in sc:
            continue
        sc[k] = v.detach().cpu() if isinstance(v, torch.Tensor) else v
    t0 = time.time()
    n = 0
    for _ in range(max_iters):
        BO.iterate(sc, cfg)
        n += 1
    dt = time.time() - t0
    return {"value": n / dt, "unit": "GN-it/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{n} iteration(s) of the same window ({dt:.1f} s)"}


# --------------------------------------------------------------------------------------------- dist helpers
def barrier(world):
    if world > 1:
        torch.distributed.barrier()


def allreduce_max(v, world, device):
    if world == 1:
        return v
    t = torch.tensor([v], dtype=torch.float64, device=device)
    torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    return float(t.item())


def allreduce_sum(v, world, device):
    if world == 1:
        return v
    t = torch.tensor([v], dtype=torch.float64, device=device)
    torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.SUM)
    return float(t.item())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="ba_window", choices=["ba_window", "track640", "kf_init"])
    ap.add_argument("--batch", type=int, default=74,
                    help="track640: independent sequences per launch (74 -> 2 CTAs per sequence on 148 SMs)")
    ap.add_argument("--kf", type=int, default=BA_K)
    ap.add_argument("--oneway", type=int, default=BA_R)
    ap.add_argument("--shard", type=int, default=0, help="1: shard the pair blocks of ONE window over the GPUs")
    ap.add_argument("--no-e2e", type=int, default=0, help="1: kernel-only sweep (tuning; no e2e / cpu_baseline legs)")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        # The reference's own CPU implementation of the path: the Python reference cannot travel to the GPU box,
        # so this arm times the oracle port (oracle/*.py, pinned to the reference by tests/golden) on the host cores.
        if rank != 0:
            return
        from como_b200 import synth

        if args.workload == "track640":
            cases = [synth.make_tracking_case(480, 640, 4, seed=b) for b in range(1)]
            for _ in range(args.warmup):
                cpu_track_baseline(cases[:1], budget_s=0.0)
            t0 = time.time()
            its = 0
            for _ in range(args.steps):
                r = cpu_track_baseline(cases[:1], budget_s=0.0)
                its += int(r["sample"].split("(")[1].split(" ")[0])
            dt = time.time() - t0
            v = its / dt
            metric, cfgd = "GN-iterations/sec at 640x480 (tracking, 4-level pyramid)", {
                "workload": "track640", "resolution": "640x480", "pyramid_levels": 4}
            sample = "one 640x480 4-level tracking problem per step"
            dtype = "f32"
        else:
            # inputs (synthetic window) are generated with the device kernels when a GPU is present -- input
            # generation is outside the timed region; the timed path is the CPU oracle only
            dev = "cuda" if torch.cuda.is_available() else None
            if dev is None:
                print(json.dumps({"impl": "reference", "unavailable": "window generation needs a CUDA device"}))
                return
            s_gpu = synth.make_ba_window(args.kf, args.oneway, 480, 640, M=64, device=dev, seed=0)
            cfg = synth.ba_cfg()
            snap = snapshot_small(s_gpu)
            from oracle import ba_oracle as BO
            for k, val in s_gpu.__dict__.items():
                if not k.startswith("_") and k not in snap:
                    snap[k] = val.detach().cpu() if isinstance(val, torch.Tensor) else val
            del s_gpu
            torch.cuda.empty_cache()
            for _ in range(min(args.warmup, 1)):
                BO.iterate(dict(snap), cfg)
            t0 = time.time()
            n = 0
            for _ in range(args.steps):
                BO.iterate(snap, cfg)
                n += 1
                if time.time() - t0 > 120:   # bounded sample
                    break
            dt = time.time() - t0
            v = n / dt
            metric, cfgd = "GN-iterations/sec at 640x480, 32-keyframe window", {
                "workload": "ba_window", "resolution": "640x480", "keyframes": args.kf, "one_way_frames": args.oneway,
                "anchors_per_kf": 64}
            sample = f"{n} iteration(s) of the same 640x480 K={args.kf} window"
            dtype = "f64"
            args.steps = n
        print(json.dumps({
            "impl": "reference", "metric": metric, "value": v, "unit": "GN-it/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / max(args.steps, 1) * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": dtype, "data": "synthetic", "config": cfgd,
            "cpu_baseline": {"value": v, "unit": "GN-it/s", "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "GN-it/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=device)
    if args.workload == "track640":
        res, cases = run_track_ours(args, rank, world, device)
        if rank == 0:
            res["cpu_baseline"] = cpu_track_baseline(cases) if (world == 1 and not args.no_e2e) else None
            print(json.dumps(res))
    elif args.workload == "kf_init":
        res, case = run_kfinit_ours(args, rank, world, device)
        if rank == 0:
            try:
                res["cpu_baseline"] = cpu_kfinit_baseline(case) if world == 1 else None
            except Exception as ex:
                res["cpu_baseline"] = {"error": repr(ex)[:200]}
            print(json.dumps(res))
    else:
        res, s, cfg = run_ba_ours(args, rank, world, device)
        if rank == 0:
            try:
                res["cpu_baseline"] = cpu_ba_baseline(s, cfg) if world == 1 else None
            except Exception as ex:  # the headline must still be printed
                res["cpu_baseline"] = {"error": repr(ex)[:200]}
            print(json.dumps(res))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
