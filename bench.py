#!/usr/bin/env python
"""bench.py -- measures the COMO photometric Gauss-Newton hot path on B200.

Contract (driver): `python bench.py --gpus N --steps K --warmup W [--impl reference]`, launched under
torchrun for N>1 (one rank per GPU).  Prints ONE JSON line on rank 0.

The headline record is the `ba_window` workload (BASELINE.json's metric: GN-iterations/sec of a 640x480,
32-keyframe window).  With the default `--workload all` the same line also carries `secondary` records so that
every number DESIGN.md quotes is produced by the driver's own run:
  track640        B independent 640x480 / 4-level frame-to-keyframe tracking problems per GPU in ONE launch
                  (the "warp + Jacobian kernel" of the metric; roofline on 52 B per pixel-iteration)
  track640_single the same tracker on ONE live sequence: per-frame latency (kernel, and through Tracking.handle_frame)
  kf_init         one keyframe creation (track_and_init) at 640x480 with 64 anchors
  sfm             one two-frame SfM bootstrap alignment at 640x480, 3 levels, 64 anchors
  ba_shard        (N > 1 only) ONE 32-keyframe window sharded over the N GPUs (strong scaling, NCCL exchange)
A "step" is one pass of the hot path over one batch of synthetic input.  Every workload's inputs exceed the
126 MB L2 (5.0 GB predictor slabs; B x 21 MB tracking operands; 157 MB predictor rows), so no L2 flush is needed.

`--impl reference` times the reference's CPU implementation of the path: the ORACLE PORT (oracle/*.py, pinned to
the unmodified reference by tests/golden -- the Python reference itself cannot travel to the GPU box), on the host
cores, on a window built by CPU generators only (the process never loads libcomo_b200.so and never touches a GPU).
`--impl reference --device cuda` runs the same port with its tensors on the GPU (stock ATen / cuBLAS / cuSOLVER
ops): the stated proxy for the reference's PyTorch-CUDA path (kind "port-cuda").
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
BA_K, BA_R = 32, 24
FP64_PEAK_TFLOPS = 37.1   # measured on B200 with scripts/micro/dmma_bench.cu (DFMA == DMMA.8x8x4 == 64 FMA/clk/SM)


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
                ["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
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
        time.sleep(0.12)
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
            if t0 <= ts <= t1 + 0.1:
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


def gpu_index_for_smi(device):
    # nvidia-smi --id counts physical devices; under CUDA_VISIBLE_DEVICES the visible index is remapped
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[device.index])
        except Exception:
            return 0
    return device.index


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured copy bandwidth (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def time_kernel(fn, min_ms=250.0, min_reps=200, max_reps=20000):
    """Average device time of `fn` (one launch), in ms: CUDA events on the launching stream around >= min_reps
    back-to-back launches lasting >= min_ms in total (a 5-launch loop is not a stable number)."""
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps, total, n = min_reps, 0.0, 0
    while total < min_ms and n < max_reps:
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        total += e0.elapsed_time(e1)
        n += reps
    return total / n, n


def timed_loop(step, steps, world, device, clock=True):
    """The contract's timed region: barrier + synchronize, CUDA events around exactly `steps` steps, max over ranks."""
    torch.cuda.synchronize()
    barrier(world)
    clk = ClockSampler(gpu_index_for_smi(device)) if clock else None
    if clk:
        clk.start()
        time.sleep(0.25)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    torch.cuda.synchronize()
    ev0.record()
    for i in range(steps):
        step(i)
    ev1.record()
    torch.cuda.synchronize()
    t1 = time.time()
    barrier(world)
    ms = allreduce_max(ev0.elapsed_time(ev1), world, device)
    return ms, (clk.stop(t0, t1) if clk else None)


# --------------------------------------------------------------------------------------------- tracking workload
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


def track_config(B, world, npx=(4800, 19200, 76800, 307200)):
    return {"workload": "track640", "resolution": "640x480", "pyramid_levels": 4, "problems_per_gpu": B,
            "px_per_level": list(npx), "l2": "inputs (B x 21 MB operands) exceed the 126 MB L2; no flush",
            "parallelism": f"replicas x{world} (independent sequences, no collective)"}


def make_trackers(cases, device):
    from como_b200.odom.Tracking import Tracking

    tcfg = {"device": str(device), "dtype": "float", "color": "gray",
            "pyr": {"start_level": 0, "end_level": 4, "depth_interp_mode": "nearest_neighbor"},
            "term_criteria": dict(TERM), "sigmas": {"photo": 1.0e-1},
            "keyframing": {"kf_depth_motion_ratio": 0.12, "kf_num_pixels_frac": 0.75, "one_way_freq": 3}}
    trackers = []
    for c in cases:
        tr = Tracking(tcfg, c["K0"].cpu(), (480, 640))
        tr.setup()
        tr.update_kf_reference(([1.0], c["rgb"], torch.eye(4, device=device)[None],
                                torch.zeros(1, 2, 1, device=device), c["depth"]))
        trackers.append(tr)
    return trackers


def run_track_ours(args, rank, world, device, B, e2e=True):
    import como_b200.odom.frontend.photo_tracking as PT
    from como_b200.odom.frontend.photo_tracking import TrackBatchPlan, photo_tracking_pyr_batch

    probs, T0, a0, cases = build_track_problems(B, device, seed0=rank * B)
    plan = TrackBatchPlan(probs, TERM)   # descriptors built once; a step = copy initial poses + one launch
    iters_total = torch.zeros((), dtype=torch.int64, device=device)

    def step(i):
        T, aff, nit = plan.run(T0, a0)
        iters_total.add_(nit.sum())

    for i in range(max(args.warmup, 3)):
        step(i)
    iters_total.zero_()
    ms, clocks = timed_loop(step, args.steps, world, device)
    its = int(allreduce_sum(int(iters_total.item()), world, device))

    # algorithmic bytes of one launch on this rank: per problem, sum over its iterations of n_level * 52 B
    # (the stats record gives the level of each iteration) -- measured once outside the timed loop
    Ts, affs, stats, nit = photo_tracking_pyr_batch(T0, a0, probs, TERM, return_stats=True)
    torch.cuda.synchronize()
    stc = stats.cpu()
    alg_bytes = 0
    for b in range(B):
        lv = stc[b, :int(nit[b]), 0].long().tolist()
        msk = [int(m.sum()) for m in cases[b]["mask"]]
        alg_bytes += sum(msk[l] * TRACK_BYTES_PER_PX_ITER for l in lv)
    # the launch alone, >= 200 launches / >= 250 ms, CUDA events on the launching stream
    kernel_ms, reps = time_kernel(lambda: plan.run(T0, a0))
    peak, peak_src = measured_peaks()
    ach = alg_bytes / (kernel_ms * 1e-3) / 1e9
    res = {
        "metric": "GN-iterations/sec at 640x480 (tracking, 4-level pyramid)", "value": its / (ms * 1e-3),
        "unit": "GN-it/s", "n_gpus": world, "steps": args.steps, "ms_per_step": ms / args.steps,
        "scaling": "weak", "dtype": "f32", "config": track_config(B, world, [int(m.sum()) for m in cases[0]["mask"]]),
        "gn_iterations_per_step": its // max(args.steps * world, 1), "gpu_launches": args.steps,
        "roofline": {"kernel": "track_pyr_kernel", "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s",
                     "frac": ach / peak,
                     # dram__bytes_read + write of one launch at this exact shape (ncu --set full,
                     # profiles/r02_track_pyr_kernel_b592_full.txt): 24.78 GB + 1.86 GB; other batch sizes: not captured
                     "traffic": 26647081000 if B == 592 else None, "peak_source": peak_src,
                     "alg_bytes_per_launch": alg_bytes, "launch_ms": kernel_ms, "launches_timed": reps,
                     "bytes_per_px_iter": TRACK_BYTES_PER_PX_ITER},
        "clocks": clocks,
    }
    if not e2e:
        return res, cases

    # e2e through the public API: B `Tracking` objects (one per sequence); each step every tracker handles one new
    # frame: pinned host RGB -> device, gray pyramid, tracking, reprojection statistics, keyframe decision (the
    # reference's handle_frame), pose read back to the host.
    trackers = make_trackers(cases, device)
    rgb_host = [c["rgb2"].cpu().pin_memory() for c in cases]  # (1,3,480,640) fp32 each
    out_host = torch.empty(B, 16, dtype=torch.float32).pin_memory()
    e_it = torch.zeros((), dtype=torch.int64, device=device)

    def e2e_step(i):
        for b, tr in enumerate(trackers):
            tr.T_curr_kf = cases[b]["T_init"].clone()
            tr.aff_curr_kf = cases[b]["aff_init"].clone()
            rgb = rgb_host[b].to(device, non_blocking=True)
            viz, _ = tr.handle_frame((2.0, rgb))
            e_it.add_(PT.last_num_iters[0])
            out_host[b].copy_(viz[1].reshape(16), non_blocking=True)

    for i in range(2):
        e2e_step(i)
    e_it.zero_()
    e_ms, _ = timed_loop(e2e_step, args.steps, world, device, clock=False)
    e_its = int(allreduce_sum(int(e_it.item()), world, device))
    res["e2e"] = {"value": e_its / (e_ms * 1e-3), "unit": "GN-it/s", "ms_per_frame": e_ms / (args.steps * B),
                  "h2d_bytes_per_step": sum(int(r.numel()) * 4 for r in rgb_host),
                  "d2h_bytes_per_step": int(out_host.numel() * 4)}
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
def ba_alg_bytes(K, N, P, HW, M):
    """Algorithmic bytes of one BA iteration: predictor apply K*HW*M*8, pair kernels
    K*N*M*8 (one predictor row per reference pixel) + P*N*48."""
    return dict(predictor_apply=K * HW * M * 8, photo=K * N * M * 8 + P * N * 48)


def ba_config(K, R, world, shard):
    """The defining parameters of the workload -- printed identically by both arms."""
    return {"workload": "ba_window", "resolution": "640x480", "keyframes": K, "one_way_frames": R,
            "anchors_per_kf": 64, "pixels_per_kf": 19200, "pairs": 2 * (K - 1) + 2 * R, "pairwise_batch_size": 128,
            "l2": "inputs (5.0 GB predictor slabs) exceed the 126 MB L2; no flush",
            "parallelism": (f"pair blocks of one window sharded by reference keyframe over {world} GPUs"
                            if shard else f"replicas x{world} (one independent window per GPU)")}


def run_ba_ours(args, rank, world, device, shard=False, extras=True):
    from como_b200 import synth
    from como_b200.odom import mapping_core as MC

    K, R, H, W, M = args.kf, args.oneway, 480, 640, 64
    shard = bool(shard and world > 1)
    # sharded: every rank holds the SAME window (its pair blocks are split); replicas: one window per rank
    s = synth.make_ba_window(K, R, H, W, M=M, device=device, seed=0 if shard else rank)
    cfg = synth.ba_cfg()
    snap = snapshot_small(s) if (rank == 0 and extras) else None   # CPU baseline runs on the untouched initial state
    comm = MC.ShardComm(world, rank, device) if shard else None

    def step(i=0):
        MC.iterate(s, cfg, comm=comm)

    for i in range(max(args.warmup, 3)):
        step()
    torch.cuda.cudart().cudaProfilerStart()   # `ncu --profile-from-start off` then sees exactly the timed steps
    ms, clocks = timed_loop(step, args.steps, world, device)
    torch.cuda.cudart().cudaProfilerStop()
    kp, pp = MC.get_plans(s, cfg, device, rank if shard else 0, world if shard else 1)
    nwin = 1 if shard else world   # sharded: one window over all GPUs (strong); else one window per GPU (weak)
    peak, peak_src = measured_peaks()
    ab = ba_alg_bytes(K, kp.N, pp.P, H * W, M)
    res = {
        "metric": "GN-iterations/sec at 640x480, 32-keyframe window", "value": nwin * args.steps / (ms * 1e-3),
        "unit": "GN-it/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "strong" if shard else "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": ba_config(K, R, world, shard),
        "window": {"landmarks": kp.L, "pairs_on_this_rank": pp.P, "system_dim": 8 * (K + R) + 3 * kp.L},
        "gpu_launches": args.steps * MC.LAUNCHES_PER_ITERATION,
        "clocks": clocks,
    }
    res["finite"] = bool(torch.isfinite(s.kf_poses).all() and torch.isfinite(s.P_m).all())
    res["final_total_err"] = float(s.total_err_prev)
    if not extras:
        return res, (s, snap), cfg

    # per-kernel device time of the two kernels the metric names, each timed alone (>= 200 launches, >= 250 ms)
    launch = MC.kernel_launchers(s, cfg, device)
    pa_ms, pa_n = time_kernel(launch["predictor_stream"])
    rs_ms, rs_n = time_kernel(launch["ba_residual"])
    res["roofline"] = {
        "kernel": "predictor_stream_kernel", "bound": "hbm", "achieved": ab["predictor_apply"] / (pa_ms * 1e-3) / 1e9,
        "peak": peak, "unit": "GB/s", "frac": ab["predictor_apply"] / (pa_ms * 1e-3) / 1e9 / peak,
        # dram__bytes_read + write of one launch at this exact shape (ncu --set full,
        # profiles/r01_predictor_stream_full.txt): 5.034 GB + 0.081 GB; other shapes: not captured
        "traffic": 5115372864 if (K, H, W, M) == (32, 480, 640, 64) else None, "peak_source": peak_src,
        "alg_bytes_per_launch": ab["predictor_apply"], "launch_ms": pa_ms, "launches_timed": pa_n,
        "note": "the streaming pass that carries 92 % of the iteration's bytes (store_vars)"}
    res["roofline_warp"] = {
        "kernel": "ba_residual_kernel", "bound": "hbm", "achieved": ab["photo"] / (rs_ms * 1e-3) / 1e9, "peak": peak,
        "unit": "GB/s", "frac": ab["photo"] / (rs_ms * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": peak_src,
        "alg_bytes_per_launch": ab["photo"], "launch_ms": rs_ms, "launches_timed": rs_n,
        "note": "BA warp + residual pass: one gathered 512-byte predictor row per reference pixel + 48 B per (pair, pixel)"}
    res["roofline_step"] = {"bound": "hbm", "achieved": (ab["predictor_apply"] + ab["photo"]) / (ms / args.steps * 1e-3) / 1e9,
                            "peak": peak, "unit": "GB/s",
                            "frac": (ab["predictor_apply"] + ab["photo"]) / (ms / args.steps * 1e-3) / 1e9 / peak,
                            "note": "all algorithmic bytes of the iteration over the whole step time"}

    # e2e: a new one-way frame arrives on the host every step (pinned RGB) -> device -> gray + gradients ->
    # replaces the oldest one-way frame -> iterate -> poses + error back to the host.  The upload of frame i+1 runs
    # on a copy stream while iteration i computes (two device buffers, events both ways): every timed step still
    # issues one full H2D copy of its input and the D2H reads of its result.
    rgb_host = synth.make_rgb(H, W, seed=77, dtype=torch.float64).pin_memory()
    res_host = torch.empty((K + R) * 16 + 1, dtype=torch.float64).pin_memory()
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
    e_ms, _ = timed_loop(lambda i: e2e_step(i + 2), args.steps, world, device, clock=False)
    res["e2e"] = {"value": nwin * args.steps / (e_ms * 1e-3), "unit": "GN-it/s",
                  "h2d_bytes_per_step": int(rgb_host.numel() * 8), "d2h_bytes_per_step": int(res_host.numel() * 8)}
    return res, (s, snap), cfg


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


def cpu_ba_baseline(s_and_snap, cfg, budget_s=25.0, min_iters=5):
    """Oracle port of Mapping.iterate on the host cores (torch CPU, fp64), on the same window (initial state)."""
    from oracle import ba_oracle as BO

    s, sc = s_and_snap
    for k, v in s.__dict__.items():
        if k.startswith("_") or k in sc:
            continue
        sc[k] = v.detach().cpu() if isinstance(v, torch.Tensor) else v
    BO.iterate(sc, cfg)   # warm-up (thread pools, allocator)
    t0 = time.time()
    n = 0
    while n < min_iters or (time.time() - t0 < budget_s and n < 50):
        BO.iterate(sc, cfg)
        n += 1
        if time.time() - t0 > 4 * budget_s:
            break
    dt = time.time() - t0
    return {"value": n / dt, "unit": "GN-it/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{n} iterations of the same window after 1 warm-up ({dt:.1f} s)"}


def torch_cuda_reference_ops(K, R, M, N, H, W, P, dim, device, reps=3):
    """LOWER BOUND on one Mapping.iterate of the reference's PyTorch-CUDA path: the reference cannot run here (lietorch
    is not vendored, como_backends builds for sm_86 only), but its dominant tensor operations can -- stock torch CUDA
    ops (cuBLAS, cuSOLVER, ATen) at the reference's own shapes and dtypes (float64), on random data:
      store_vars            exp(Knm_Kmminv (K,H,W,M) @ logz_m) + per-keyframe median            Mapping.py:749-758
      setup_test_points     gather (K,N,M) predictor rows, materialise dPwn_dzm (K,N,3,M,1)       sparse_map.py:184-230
      create_photo_system   per-pair gathers: dPwn_dzm[ref] (P,N,3,M,1), images[target] (P,3,H,W) photo.py:262-347
      batch_photo_cost      grid_sample, global median, dIt_dzm = dIt_dPwn @ dPwn_dzm, the Gram
                            einsum (P,N,M)->(P,M,M), M->3M expansion, scatter_add into H         photo.py:83-233
      solve_system          cholesky_ex + cholesky_solve at dim                                   linear_system.py:101-112
    Everything else the reference does (projection Jacobians, priors, pose blocks, robust weights ...) is left out, so
    the real reference is slower than this sum."""
    f64 = torch.float64
    g = torch.Generator(device=device).manual_seed(0)
    rnd = lambda *sh: torch.rand(*sh, dtype=f64, device=device, generator=g)
    Knm = rnd(K, H, W, M)
    logzm = rnd(K, 1, M, 1)
    rows = torch.randint(0, H, (K, N), device=device, generator=g)
    cols = torch.randint(0, W, (K, N), device=device, generator=g)
    kidx = torch.arange(K, device=device)[:, None].expand(K, N)
    ref_ids = torch.randint(0, K, (P,), device=device, generator=g)
    tgt_ids = torch.randint(0, K, (P,), device=device, generator=g)
    imgs = rnd(K, 3, H, W)
    ray = rnd(K, N, 3, 1, 1)
    dIt_dPwn = rnd(P, N, 1, 3)
    grid = rnd(P, N, 1, 2) * 2 - 1
    Hm = torch.zeros(dim * dim, dtype=f64, device=device)
    idx = torch.randint(0, dim * dim, (P, 3 * M * 3 * M), device=device, generator=g)
    A = rnd(dim, dim)
    Hspd = A @ A.T + dim * torch.eye(dim, dtype=f64, device=device)
    gvec = rnd(dim, 1)
    dz = rnd(P, 3)
    out = {}

    def timed(name, fn):
        # a LOWER bound: the fastest of `reps` repetitions after a warm-up (a repetition that has to cudaMalloc a
        # multi-GB result -- the caching allocator's choice, not the operator's cost -- would otherwise inflate it)
        r = fn()
        torch.cuda.synchronize()
        best = float("inf")
        for _ in range(reps):
            r = None
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = fn()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        out[name] = best
        return r

    def store_vars():
        depth = torch.exp(Knm @ logzm)
        return torch.median(depth.view(K, -1), dim=1).values

    def test_points():
        Kt = Knm[kidx, rows, cols, :]                                   # (K,N,M)
        return (ray * Kt[:, :, None, :, None]).contiguous()            # dPwn_dzm (K,N,3,M,1)

    timed("store_vars", store_vars)
    dPwn_dzm = timed("setup_test_points", test_points)

    def gathers():
        return dPwn_dzm[ref_ids], imgs[tgt_ids]

    dP_b, img_b = timed("create_photo_system_gathers", gathers)

    def photo():
        v = torch.nn.functional.grid_sample(img_b, grid, mode="bilinear", padding_mode="zeros", align_corners=False)
        med = torch.median(v[:, 0].abs().reshape(-1))
        Jz = dIt_dPwn @ dP_b.view(P, N, 3, M)                          # (P,N,1,M)
        Hz = torch.einsum("bnck,bncl->bkl", Jz, Jz)                    # (P,M,M)
        HP = (dz[:, None, :, None, None] * Hz[:, :, None, :, None] * dz[:, None, None, None, :]).reshape(P, -1)
        Hm.scatter_add_(0, idx.reshape(-1), HP.reshape(-1))
        return med

    timed("batch_photo_cost", photo)
    del dP_b, img_b, dPwn_dzm

    def solve():
        Lc, _ = torch.linalg.cholesky_ex(Hspd, upper=False, check_errors=False)
        return torch.cholesky_solve(gvec, Lc, upper=False)

    timed("solve_system", solve)
    total = sum(out.values())
    return {"value": 1e3 / total, "unit": "GN-it/s", "ms_per_step": total, "kind": "torch-cuda-lower-bound",
            "ops_ms": {k: round(v, 3) for k, v in out.items()},
            "sample": "dominant tensor ops of the reference's Mapping.iterate at its own shapes (float64), stock torch "
                      "CUDA; a lower bound on the reference's PyTorch-CUDA time per iteration"}


def torch_cuda_ba_baseline(s, cfg, device, iters=5):
    """The oracle port with its tensors on the GPU: stock ATen / cuBLAS / cuSOLVER ops, no como_b200 kernel.  Stated
    proxy for the reference's PyTorch-CUDA path (the reference cannot run here: lietorch is not vendored and its
    como_backends builds for sm_86 only).  The port's fused rank-1 formulation does LESS work than the reference
    (no (b,N,3,M,1) Jacobian tensor), so it is a conservative (fast) stand-in."""
    from oracle import ba_oracle as BO

    sc = {}
    for k, v in s.__dict__.items():
        if k.startswith("_"):
            continue
        sc[k] = v.detach().clone() if (isinstance(v, torch.Tensor) and v.numel() < (1 << 24)) else v
    with torch.device(device):
        BO.iterate(sc, cfg)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            BO.iterate(sc, cfg)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    return {"value": 1e3 / ms, "unit": "GN-it/s", "ms_per_step": ms, "kind": "port-cuda",
            "sample": f"{iters} iterations of the same window after 1 warm-up; oracle port on stock torch CUDA ops"}


# --------------------------------------------------------------------------------------------- keyframe-creation workload
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
    from como_b200.odom.frontend.corr import track_and_init

    c = build_kfinit_case(device, seed=rank)
    H, W, M = c["H"], c["W"], c["M"]
    out = [None]

    def step(i=0, cov2=None, z_img=None):
        out[0] = track_and_init(c["pose1"], c["pose2"], c["coords_m1"], c["z_m1"], c["z_img1"] if z_img is None else z_img,
                                c["cov2"] if cov2 is None else cov2, c["K"], c["scale"], KFINIT_CORR, KFINIT_SAMP, (H, W))
        return out[0]

    for _ in range(max(args.warmup, 3)):
        step()
    ms, clocks = timed_loop(step, args.steps, world, device)

    # end to end: covariance + depth images arrive from pinned host memory, the new anchors go back
    cov_host = c["cov2"].cpu().pin_memory()
    z_host = c["z_img1"].cpu().pin_memory()
    out_host = torch.empty(M * 3 + M, dtype=torch.float64).pin_memory()

    def e2e_step(i):
        cov2 = cov_host.to(device, non_blocking=True)
        zi = z_host.to(device, non_blocking=True)
        c2, z2, mask, call, zall = step(0, cov2, zi)
        n = call.shape[1]
        out_host[:2 * n].copy_(call.reshape(-1), non_blocking=True)
        out_host[2 * M:2 * M + n].copy_(zall.reshape(-1), non_blocking=True)
        out_host[3 * M:3 * M + mask.numel()].copy_(mask.to(torch.float64), non_blocking=True)

    for i in range(2):
        e2e_step(i)
    e_ms, _ = timed_loop(e2e_step, args.steps, world, device, clock=False)

    # roofline of the dominant kernel (K-matrix / predictor rows with variance), timed alone
    n = H * W
    coords_n = torch.stack((torch.rand(n, device=device) * (H - 1), torch.rand(n, device=device) * (W - 1)), -1)[None].double()
    mask = torch.ones(n, dtype=torch.uint8, device=device)
    E_m = torch.empty(1, M, 4, dtype=torch.float64, device=device)
    K_mm = torch.empty(1, M, M, dtype=torch.float64, device=device)
    st = _lib.stream_ptr(device)
    _lib.kmat_kmm(_lib.ptr(c["cov2"]), 1, H, W, _lib.ptr(c["coords_m1"]), M, c["scale"], 0.0, _lib.ptr(E_m), _lib.ptr(K_mm), st)
    Kinv = torch.linalg.inv(K_mm).contiguous()
    rows = torch.empty(n, M, dtype=torch.float64, device=device)
    var = torch.empty(n, dtype=torch.float64, device=device)
    vmin = torch.empty(1, dtype=torch.float64, device=device)
    kernel_ms, reps = time_kernel(lambda: _lib.kmat_rows(
        _lib.ptr(c["cov2"]), 1, H, W, _lib.ptr(c["coords_m1"]), _lib.ptr(E_m), _lib.ptr(Kinv), M, c["scale"],
        _lib.ptr(coords_n), _lib.ptr(mask), n, _lib.ptr(rows), _lib.ptr(var), _lib.ptr(vmin), st))
    alg_flop = 2.0 * n * M * M + 30.0 * n * M   # GEMM 2 n m^2 + ~30 flop per kernel evaluation
    ach = alg_flop / (kernel_ms * 1e-3) * 1e-12
    o = out[0]
    res = {
        "metric": "keyframe creations/sec at 640x480 (track_and_init, 64 anchors)", "value": world * args.steps / (ms * 1e-3),
        "unit": "KF/s", "n_gpus": world, "steps": args.steps, "ms_per_step": ms / args.steps, "scaling": "weak", "dtype": "f64",
        "config": {"workload": "kf_init", "resolution": "640x480", "anchors": M, "dense_points": n,
                   "new_anchors": int(o[0].shape[1]), "correspondences": int(o[2].sum()),
                   "l2": "predictor rows (157 MB per pass, written once and read twice) exceed the 126 MB L2; no flush",
                   "parallelism": f"replicas x{world} (independent keyframes, no collective)"},
        "e2e": {"value": world * args.steps / (e_ms * 1e-3), "unit": "KF/s",
                "h2d_bytes_per_step": int(cov_host.numel() + z_host.numel()) * 8, "d2h_bytes_per_step": int(out_host.numel()) * 8},
        "roofline": {"kernel": "kmat_rows_kernel", "bound": "tensor", "achieved": ach, "peak": FP64_PEAK_TFLOPS,
                     "unit": "TFLOP/s", "frac": ach / FP64_PEAK_TFLOPS, "traffic": None,
                     "peak_source": "fp64 DFMA/DMMA peak measured with scripts/micro/dmma_bench.cu (MEASURED_PEAKS.json has no fp64 entry)",
                     "alg_flop_per_launch": alg_flop, "launch_ms": kernel_ms, "launches_timed": reps},
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



# --------------------------------------------------------------------------------------------- two-frame SfM workload
SFM_INIT = {"start_level": 0, "end_level": 3, "max_iter": 50, "delta_norm": 1.0e-4, "rel_tol": 1.0e-4,
            "kf_depth_motion_ratio": 0.04, "kf_num_pixels_frac": 0.75}    # config/como.yml:65-72


def build_sfm_case(device, seed=0, H=480, W=640, M=64, shift_px=2.5):
    """Synthetic bootstrap pair: frame 1 is frame 0 translated by 2.5 px; 3-level [I, gx, gy] pyramids in
    the mapper's dtype, anchors chosen by the sampler, the reference's identity / zero-log-depth initialisation."""
    import torch.nn.functional as F

    from como_b200 import synth
    from como_b200.depth_cov.core.samplers import sample_sparse_coords
    from como_b200.odom import mapping_core as MC
    from como_b200.odom.frontend import two_frame_sfm as SF

    tex = synth.make_rgb(H, W, seed=seed + 1, cell=16, extra_w=16, dtype=torch.float64, device=device)
    x1, a = int(shift_px), shift_px - int(shift_px)
    rgb0 = tex[..., 0:W].contiguous()
    rgb1 = ((1 - a) * tex[..., x1:x1 + W] + a * tex[..., x1 + 1:x1 + 1 + W]).contiguous()

    def pyr(rgb):   # coarsest first; the down-sampling filter is part of the synthetic input, not of the timed path
        out, cur = [], rgb
        for l in range(SFM_INIT["end_level"]):
            out.insert(0, MC.get_img_and_grads(cur))
            cur = F.avg_pool2d(cur, 2)
        return out

    iag0, iag1 = pyr(rgb0), pyr(rgb1)
    cov = synth.make_cov_image_wide(H, W, seed=seed).to(device)
    scale = 0.086
    coords_m, _ = sample_sparse_coords(cov, M, "greedy_conditional_entropy", 1e-2, border=3, dist_thresh=0.1,
                                       signal_var=scale, fixed_var=0.0)
    dims = torch.tensor([H, W], dtype=torch.float64, device=device)
    cm_norm = 2.0 * (1.0 / dims) * coords_m.double() + (1.0 / dims) - 1.0
    K = synth.make_intrinsics(H, W, dtype=torch.float64).to(device)
    vals, coords, Knm, sizes, Kp, dr, Hp = SF.setup_reference(iag0, cm_norm, scale, cov, K)
    T0 = torch.eye(4, dtype=torch.float64, device=device)[None]
    d0 = torch.zeros(1, coords_m.shape[1], 1, dtype=torch.float64, device=device)
    aff = torch.zeros(1, 2, 1, dtype=torch.float64, device=device)
    return dict(T0=T0, d0=d0, aff=aff, coords=coords, vals=vals, Knm=Knm, imgs=iag1, dr=dr, Hp=Hp, Kp=Kp, rgb1=rgb1,
                H=H, W=W, M=int(coords_m.shape[1]), pyr=pyr)


def run_sfm_ours(args, rank, world, device):
    from como_b200.odom.frontend import two_frame_sfm as SF

    c = build_sfm_case(device, seed=rank)
    its = [0]

    def step(i=0, imgs=None):
        out = SF.two_frame_sfm_pyr(c["T0"], c["d0"], c["aff"], c["coords"], c["vals"], c["Knm"], c["imgs"] if imgs is None else imgs,
                                   c["dr"], c["Hp"], c["Kp"], {"photo": 0.1}, None, SFM_INIT)
        its[0] += sum(SF.two_frame_sfm_pyr.last_iters)
        return out

    for _ in range(max(min(args.warmup, 3), 2)):
        step()
    its[0] = 0
    steps = max(2, min(args.steps, 5))     # one alignment is ~50 GN iterations over three levels
    ms, clocks = timed_loop(step, steps, world, device)
    n_it = its[0]
    # end to end: the new frame arrives as pinned host RGB; gray + gradients + pyramid on the device, pose and depths back
    rgb_host = c["rgb1"].cpu().pin_memory()
    out_host = torch.empty(16 + c["M"], dtype=torch.float64).pin_memory()
    its[0] = 0

    def e2e_step(i):
        rgb = rgb_host.to(device, non_blocking=True)
        out = step(0, c["pyr"](rgb))
        out_host[:16].copy_(out[0].reshape(-1), non_blocking=True)
        out_host[16:].copy_(out[1].reshape(-1), non_blocking=True)

    e2e_step(0)
    its[0] = 0
    e_ms, _ = timed_loop(e2e_step, steps, world, device, clock=False)
    e_it = its[0]
    px = [int(v.numel()) for v in c["vals"]]
    return {
        "metric": "GN-iterations/sec (two-frame SfM bootstrap, 640x480, 3 levels, 64 anchors)",
        "value": world * n_it / (ms * 1e-3), "unit": "GN-it/s", "n_gpus": world, "steps": steps,
        "ms_per_step": ms / steps, "ms_per_alignment": ms / steps, "gn_iterations_per_alignment": n_it // steps,
        "scaling": "weak", "dtype": "f64",
        "config": {"workload": "sfm", "resolution": "640x480", "anchors": c["M"], "px_per_level": px,
                   "l2": "predictor rows (157 MB at the finest level) exceed the 126 MB L2; no flush",
                   "parallelism": f"replicas x{world} (independent bootstraps, no collective)"},
        "e2e": {"value": world * e_it / (e_ms * 1e-3), "unit": "GN-it/s", "ms_per_alignment": e_ms / steps,
                "h2d_bytes_per_step": int(rgb_host.numel()) * 8, "d2h_bytes_per_step": int(out_host.numel()) * 8},
        "clocks": clocks,
    }


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


# --------------------------------------------------------------------------------------------- reference arm
def cpu_window(K, R, seed=0):
    """The synthetic 640x480 window of the GPU arm built by CPU code only (oracle generators): a farthest-point anchor
    picker instead of the 5 s/keyframe greedy sampler (anchor placement does not change the work of an iteration)
    and the oracle's K-matrix predictor.  Untimed set-up of the reference arm."""
    import numpy as np

    from como_b200 import synth
    from oracle import depthcov_oracle as DO

    rng = np.random.default_rng(1234 + seed)

    def sampler(cov, n, curr):
        H, W = cov.shape[-2:]
        gy, gx = np.meshgrid(np.linspace(8, H - 9, 20), np.linspace(8, W - 9, 26), indexing="ij")
        cand = np.round(np.stack((gy.ravel(), gx.ravel()), 1) + rng.uniform(-4, 4, (gy.size, 2)))
        have = curr[0].double().numpy() if curr is not None and curr.numel() else np.zeros((0, 2))
        pts = [p for p in have]
        chosen = []
        used = np.zeros(cand.shape[0], dtype=bool)
        for _ in range(max(n - have.shape[0], 0)):
            ref = np.array(pts) if pts else np.array([[-1e9, -1e9]])
            d = ((cand[:, None, :] - ref[None]) ** 2).sum(-1).min(1)
            d[used] = -1.0
            i = int(d.argmax())
            used[i] = True
            chosen.append(cand[i].copy())
            pts.append(cand[i].copy())
        return torch.tensor(np.array(chosen).reshape(-1, 2), dtype=torch.float64)[None]

    def predictor(cov, coords):
        return DO.prep_predictor(cov.double(), coords, 1.0)

    return synth.make_ba_window(K, R, 480, 640, M=64, device="cpu", seed=seed, sampler=sampler, predictor=predictor)


def run_reference(args, rank):
    """`--impl reference`: the port on the host cores (default) or on stock torch CUDA ops (--device cuda)."""
    if rank != 0:
        return
    from como_b200 import synth

    ncores = os.cpu_count() or 1
    torch.set_num_threads(ncores)   # torchrun exports OMP_NUM_THREADS=1: the CPU arm gets every host core explicitly
    on_cuda = args.device == "cuda"
    kind = "port-cuda" if on_cuda else "port"
    what = ("oracle port of the reference (oracle/*.py, pinned by tests/golden) on "
            + ("stock torch CUDA ops" if on_cuda else f"{ncores} host threads"))
    wl = "ba_window" if args.workload == "all" else args.workload
    dev = "cuda" if on_cuda else "cpu"
    if wl == "track640":
        from oracle import track_oracle as TO

        case = synth.make_tracking_case(480, 640, 4, seed=0, device=dev)

        def once():
            with torch.device(dev):
                return len(TO.track_pyr(case["T_init"], case["aff_init"], case["vals"], case["P"], case["dI_dT"],
                                        case["mask"], case["K"], case["img"], TERM)[2])

        for _ in range(max(args.warmup, 1)):
            once()
        t0 = time.time()
        its = sum(once() for _ in range(args.steps))
        if on_cuda:
            torch.cuda.synchronize()
        dt = time.time() - t0
        v, steps = its / dt, args.steps
        metric, cfgd, dtype = "GN-iterations/sec at 640x480 (tracking, 4-level pyramid)", track_config(1, 1), "f32"
        sample = f"{steps} x one 640x480 4-level tracking problem ({its} GN iterations)"
    else:
        from oracle import ba_oracle as BO

        if on_cuda:
            s = synth.make_ba_window(args.kf, args.oneway, 480, 640, M=64, device="cuda", seed=0)
        else:
            s = cpu_window(args.kf, args.oneway, seed=0)
        cfg = synth.ba_cfg()
        sc = {k: v for k, v in s.__dict__.items() if not k.startswith("_")}
        with torch.device(dev):
            for _ in range(min(max(args.warmup, 1), 2)):
                BO.iterate(sc, cfg)
            if on_cuda:
                torch.cuda.synchronize()
            t0 = time.time()
            n = 0
            for _ in range(args.steps):
                BO.iterate(sc, cfg)
                n += 1
                if time.time() - t0 > 150:   # bounded sample
                    break
            if on_cuda:
                torch.cuda.synchronize()
        dt = time.time() - t0
        v, steps = n / dt, n
        metric, dtype = "GN-iterations/sec at 640x480, 32-keyframe window", "f64"
        cfgd = ba_config(args.kf, args.oneway, args.gpus, False)
        sample = f"{n} iterations of one 640x480 K={args.kf} window after warm-up ({dt:.1f} s)"
    print(json.dumps({
        "impl": "reference", "metric": metric, "reference_is": what, "value": v, "unit": "GN-it/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": args.warmup, "ms_per_step": dt / max(steps, 1) * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": dtype, "data": "synthetic", "config": cfgd,
        "cpu_baseline": {"value": v, "unit": "GN-it/s", "cores": 0 if on_cuda else torch.get_num_threads(), "kind": kind,
                         "sample": sample},
        "e2e": {"value": v, "unit": "GN-it/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


# --------------------------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--device", default="cpu", choices=["cpu", "cuda"], help="--impl reference only: where the port runs")
    ap.add_argument("--workload", default="all", choices=["all", "ba_window", "track640", "kf_init", "sfm"])
    ap.add_argument("--batch", type=int, default=592, help="track640: independent sequences per launch (592 = 4 CTAs on each of the 148 SMs, one problem per CTA)")
    ap.add_argument("--kf", type=int, default=BA_K)
    ap.add_argument("--oneway", type=int, default=BA_R)
    ap.add_argument("--shard", type=int, default=0, help="1: headline = ONE window sharded over the GPUs (strong scaling)")
    ap.add_argument("--no-e2e", type=int, default=0, help="1: kernel-only sweep (tuning; no e2e / cpu_baseline legs)")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=device)

    def guarded(name, fn):
        try:
            return fn()
        except Exception as ex:   # the headline must still be printed
            return {"error": f"{name}: {repr(ex)[:300]}"}

    if args.workload == "track640":
        res, cases = run_track_ours(args, rank, world, device, args.batch, e2e=not args.no_e2e)
        res.update({"warmup": args.warmup, "higher_is_better": True, "vs_baseline": None, "data": "synthetic"})
        if rank == 0:
            res["cpu_baseline"] = cpu_track_baseline(cases) if (world == 1 and not args.no_e2e) else None
            print(json.dumps(res))
    elif args.workload == "sfm":
        res = run_sfm_ours(args, rank, world, device)
        res.update({"warmup": args.warmup, "higher_is_better": True, "vs_baseline": None, "data": "synthetic"})
        if rank == 0:
            print(json.dumps(res))
    elif args.workload == "kf_init":
        res, case = run_kfinit_ours(args, rank, world, device)
        res.update({"warmup": args.warmup, "higher_is_better": True, "vs_baseline": None, "data": "synthetic"})
        if rank == 0:
            res["cpu_baseline"] = guarded("cpu_baseline", lambda: cpu_kfinit_baseline(case)) if world == 1 else None
            print(json.dumps(res))
    else:
        res, s, cfg = run_ba_ours(args, rank, world, device, shard=bool(args.shard), extras=not args.no_e2e)
        if args.workload == "all" and not args.no_e2e:
            sec = {}
            if world == 1:
                sec["torch_cuda_port"] = guarded("torch_cuda_port", lambda: torch_cuda_ba_baseline(s[0], cfg, device, iters=2))
            if rank == 0:
                res["cpu_baseline"] = guarded("cpu_baseline", lambda: cpu_ba_baseline(s, cfg)) if world == 1 else None
            win = res["window"]
            del s
            torch.cuda.empty_cache()
            if world == 1:
                sec["torch_cuda_reference_ops"] = guarded("torch_cuda_reference_ops", lambda: torch_cuda_reference_ops(
                    args.kf, args.oneway, 64, 19200, 480, 640, win["pairs_on_this_rank"], win["system_dim"], device))
                torch.cuda.empty_cache()
            barrier(world)
            if world > 1 and not args.shard:
                sec["ba_shard"] = guarded("ba_shard", lambda: run_ba_ours(args, rank, world, device, shard=True, extras=False)[0])
                torch.cuda.empty_cache()

            def trk(B, e2e):
                r, cases = run_track_ours(args, rank, world, device, B, e2e=e2e)
                if rank == 0 and world == 1 and B > 1:
                    r["cpu_baseline"] = cpu_track_baseline(cases, budget_s=8.0)
                return r

            sec["track640"] = guarded("track640", lambda: trk(args.batch, True))
            torch.cuda.empty_cache()
            sec["track640_single"] = guarded("track640_single", lambda: trk(1, True))
            sec["kf_init"] = guarded("kf_init", lambda: run_kfinit_ours(args, rank, world, device)[0])
            torch.cuda.empty_cache()
            sec["sfm"] = guarded("sfm", lambda: run_sfm_ours(args, rank, world, device))
            res["secondary"] = sec
        elif rank == 0 and not args.no_e2e:
            res["cpu_baseline"] = guarded("cpu_baseline", lambda: cpu_ba_baseline(s, cfg)) if world == 1 else None
        if rank == 0:
            print(json.dumps(res))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
