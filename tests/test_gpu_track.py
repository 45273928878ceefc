"""GPU parity tests for the tracking path: CUDA (through the C ABI) vs the oracle and the
reference-generated golden vectors.  Tolerances: residual norm 1e-4 rel (+ the median's rank
quantisation 2/nvalid at tiny sizes), SE(3) log 1e-3 (we assert tighter)."""
import os

import numpy as np
import pytest
import torch

from oracle import track_oracle as TO
from test_oracle_track import level_inputs, load, se3_log_err, TERM

pytestmark = pytest.mark.gpu


def cuda_track(levels, T0, aff0, term, stats=True):
    from como_b200.odom.frontend.photo_tracking import photo_tracking_pyr

    dev = "cuda:0"
    c = lambda x: x.to(dev)
    return photo_tracking_pyr(c(T0), c(aff0), [c(v) for v in levels["vals"]], [c(v) for v in levels["P"]],
                              [c(v) for v in levels["dI_dT"]], [c(v) for v in levels["mask"]],
                              [v for v in levels["K"]], [c(v) for v in levels["img"]], 0.1, term, return_stats=stats)


def golden_levels(g):
    nl = int(g["num_levels"])
    t = lambda k: torch.from_numpy(g[k])
    return dict(vals=[t(f"vals_{l}") for l in range(nl)], P=[t(f"P_{l}") for l in range(nl)],
                dI_dT=[t(f"dI_dT_{l}") for l in range(nl)], mask=[t(f"mask_{l}") for l in range(nl)],
                K=[t(f"K_{l}") for l in range(nl)], img=[t(f"img_{l}") for l in range(nl)])


@pytest.mark.parametrize("name", ["track_80x60_l3", "track_80x60_l3_it1", "track_160x120_l4", "track_80x60_l3_rgb"])
def test_track_pyr_vs_reference_golden(golden_dir, name):
    g = load(golden_dir, name)
    term = dict(TERM, max_iter=int(g["max_iter"]))
    T, aff, stats = cuda_track(golden_levels(g), torch.from_numpy(g["T_init"]), torch.from_numpy(g["aff_init"]), term)
    assert stats.shape[0] == len(g["trace_mse"]), "iteration count differs from the reference"
    assert se3_log_err(T[0].cpu().numpy(), g["T_final"][0]) < 1e-4
    np.testing.assert_allclose(aff.cpu().numpy().ravel(), g["aff_final"].ravel(), atol=1e-4)


@pytest.mark.parametrize("name", ["track_80x60_l3", "track_160x120_l4", "track_80x60_l3_rgb"])
def test_tracking_iter_same_inputs_vs_golden(golden_dir, name):
    """Replay each recorded reference iteration from its own (T, aff): one level, max_iter 1.  The rgb golden
    (tracking.color: rgb, C = 3) exercises the channel-stacked tiles: valid counts are pixels, sums run over channels."""
    g = load(golden_dir, name)
    lv = golden_levels(g)
    counts = [int(torch.from_numpy(g[f"mask_{l}"]).sum()) for l in range(int(g["num_levels"]))]
    for i in range(len(g["trace_mse"])):
        l = counts.index(int(g["trace_n"][i]))
        one = {k: [v[l]] for k, v in lv.items()}
        T, aff, stats = cuda_track(one, torch.from_numpy(g["trace_T_in"][i]), torch.from_numpy(g["trace_aff_in"][i]),
                                   dict(TERM, max_iter=1))
        st = stats[0].cpu().numpy()
        nvalid = int(g["trace_nvalid"][i])
        assert abs(st[5] - nvalid) <= 2
        assert abs(st[1] - g["trace_mse"][i]) <= (1e-4 + 2.0 / nvalid) * g["trace_mse"][i]
        # |g| is what is left of sum w J r after cancellation; three channels leave three times the rounding noise
        assert abs(st[2] - g["trace_gnorm"][i]) <= (2e-3 if "rgb" not in name else 6e-3) * max(g["trace_gnorm"][i], 1.0)
        assert se3_log_err(T[0].cpu().numpy(), g["trace_T_out"][i, 0]) < 1e-5
        np.testing.assert_allclose(aff.cpu().numpy().ravel(), g["trace_aff_out"][i].ravel(), atol=1e-5)


def replay_check(levels, stats, T_final, aff_final, strict=True):
    """Every iteration the CUDA path ran, replayed by the oracle from the CUDA path's OWN iterate (stats[i, 8:26]):
    both sides evaluate the same (T, aff) on the same operands, so the bounds apply per iteration and do not depend
    on how either side's fp32 summation order steers the trajectory.

    Validity (1 <= x < w-1, 1 <= y < h-1) is a hard threshold on an fp32 projection.  When the pose is close to the
    identity -- the first iteration of every level in these cases -- a whole border row/column of reference pixels
    projects within an ulp of the threshold, and which side each lands on depends on the last bit (FMA contraction).
    The oracle's per-pixel data gives an exact bound for that: B = the pixels within 4 ulp of a threshold
    (|margin| <= 4 * 2^-23 * max(w, h)); their count bounds the difference in nvalid and in the median's rank, and the
    sum of their individual contributions bounds the difference in every accumulated quantity.  On top of that:
      sigma    2e-6 + 1e-5 sigma: sigma is an EXACT order statistic on both sides (test_exact_order_statistic_bitwise),
               i.e. ONE pixel's |r|, and that residual carries the fp32 noise of its warped coordinate (~1 ulp of x
               times the image gradient): up to ~2e-6 in intensity, it does not average out
      norm     north star, residual norms 1e-4 relative: the un-normalised robust norm sqrt(sum w r^2 / n)
               = sigma sqrt(mean_sq_err), + 3e-7 (2.5 ulp(1): the norm is only ~2e-3 at a converged coarse level)
      mse      the sigma^2-normalised value the reference reports: the same plus twice sigma's relative bound
      |g|      what is left of sum w J r after cancellation (3 of ~1e3 near convergence): 5e-3 of itself
      update   north star, SE(3) log 1e-3: asserted 1e-4 on (next iterate vs oracle update); affine 1e-4."""
    st = stats.cpu().numpy()
    masked = {}
    rows, bad = [], []
    for i in range(st.shape[0]):
        l = int(st[i, 0])
        if l not in masked:
            m = levels["mask"][l].reshape(-1).cpu()
            masked[l] = (levels["vals"][l].reshape(-1).cpu()[m], levels["P"][l].reshape(-1, 3).cpu()[m],
                         levels["dI_dT"][l].reshape(-1, 8).cpu()[m], levels["K"][l].cpu(), levels["img"][l][0, 0].cpu())
        vals, P, J, K, img = masked[l]
        T_in = torch.from_numpy(st[i, 8:24].reshape(4, 4).copy())
        aff_in = torch.from_numpy(st[i, 24:26].copy())
        d = {}
        Tn, affn, delta, mse, gn, H, g, sigma, nvalid = TO.tracking_iter(T_in, aff_in, vals, P, J, K, img, detail=d)
        # ---- pixels whose validity hangs on the last bits of the projection
        wd, hd = d["w"], d["h"]
        margin = torch.minimum(torch.minimum((d["x"] - 1).abs(), (d["x"] - (wd - 1)).abs()),
                               torch.minimum((d["y"] - 1).abs(), (d["y"] - (hd - 1)).abs()))
        near = (margin <= 4 * 2.0 ** -23 * max(wd, hd)) & (d["z"] > 0)
        nb = int(near.sum())
        rb = d["r"][near].abs().double()
        contrib = torch.minimum(rb * rb, 1.345 * sigma * rb)                      # w r^2 of each such pixel
        flip_sum = float(contrib.sum()) / max(nvalid, 1)                           # effect on mean(w r^2)
        flip_g = float((torch.minimum(rb, torch.full_like(rb, 1.345 * sigma)) * d["Jf"][near].double().norm(dim=1)).sum())
        q = 2.0 / max(nvalid, 1) if not strict else 0.0   # median rank quantisation, only matters at toy sizes
        dn = abs(st[i, 5] - nvalid)
        tol_s = 2e-6 + (1e-5 + q + 2.0 * nb / max(nvalid, 1)) * sigma
        norm_g, norm_o = float(st[i, 4]) * np.sqrt(float(st[i, 1])), sigma * np.sqrt(mse)
        tol_n = (1e-4 + q) * norm_o + 3e-7 + flip_sum / (2 * norm_o)
        T_next = st[i + 1, 8:24].reshape(4, 4) if i + 1 < st.shape[0] else T_final
        a_next = st[i + 1, 24:26] if i + 1 < st.shape[0] else aff_final
        chk = dict(nvalid=(dn, max(nb, 2 if not strict else 0)), sigma=(abs(st[i, 4] - sigma), tol_s),
                   norm=(abs(norm_g - norm_o), tol_n),
                   mse=(abs(st[i, 1] - mse), (2 * tol_n / norm_o + 2 * tol_s / sigma) * mse),
                   gnorm=(abs(st[i, 2] - gn), 5e-3 * max(gn, 1.0) + flip_g), se3=(se3_log_err(T_next, Tn.numpy()), 1e-4),
                   aff=(float(np.abs(a_next - affn.numpy()).max()), 1e-4))
        rows.append(f"it {i} L{l} n={nvalid} border={nb} sigma={sigma:.6g} " +
                    " ".join(f"{k}={v[0]:.3g}/{v[1]:.3g}" for k, v in chk.items()))
        bad += [(i, k) for k, v in chk.items() if not v[0] <= v[1]]
    assert not bad, f"out of bounds {bad}:\n" + "\n".join(rows)
    return st


@pytest.fixture(scope="module")
def full_case():
    from como_b200 import synth

    case = synth.make_tracking_case(480, 640, 4, seed=0)
    T, aff, stats = cuda_track(case, case["T_init"], case["aff_init"], TERM)
    return case, T, aff, stats


def test_full_size_640x480_every_iteration_replayed(full_case):
    """BASELINE config 2 shape (640x480, 4-level pyramid): each CUDA iteration vs the oracle on identical inputs."""
    case, T, aff, stats = full_case
    assert stats.shape[0] >= 4
    replay_check(case, stats, T[0].cpu().numpy(), aff.cpu().numpy().ravel())


def test_full_size_640x480_end_to_end_vs_oracle_trajectory(full_case):
    """Independent trajectories (the oracle's own fp32 summation order depends on the host thread count, so the
    iterates are not comparable one by one): the final pose must agree within the north-star SE(3)-log bound."""
    case, T, aff, stats = full_case
    To, affo, trace = TO.track_pyr(case["T_init"], case["aff_init"], case["vals"], case["P"], case["dI_dT"],
                                   case["mask"], case["K"], case["img"], TERM)
    assert abs(stats.shape[0] - len(trace)) <= 2
    assert se3_log_err(T[0].cpu().numpy(), To.numpy()) < 1e-3
    np.testing.assert_allclose(aff.cpu().numpy().ravel(), affo.numpy().ravel(), atol=1e-3)
    # property: converges back to identity from the 8d perturbation
    assert se3_log_err(T[0].cpu().numpy(), np.eye(4)) < 2e-3


def test_full_size_finest_level_single_iteration_strict():
    """N = 307200, one iteration from identical inputs: residual norm 1e-4, exact order statistic, update 1e-5."""
    from como_b200 import synth

    case = synth.make_tracking_case(480, 640, 4, seed=0)
    one = {k: [v[-1]] for k, v in case.items() if isinstance(v, list)}
    T1, aff1, st1 = cuda_track(one, case["T_init"], case["aff_init"], dict(TERM, max_iter=1))
    m = case["mask"][-1].reshape(-1)
    r = TO.tracking_iter(case["T_init"][0], case["aff_init"].reshape(2), case["vals"][-1].reshape(-1)[m],
                         case["P"][-1].reshape(-1, 3)[m], case["dI_dT"][-1].reshape(-1, 8)[m], case["K"][-1],
                         case["img"][-1][0, 0])
    s = st1[0].cpu().numpy()
    assert abs(s[5] - r[8]) <= 2
    assert abs(s[1] - r[3]) <= 1e-4 * r[3]
    assert abs(s[2] - r[4]) <= 1e-4 * r[4]
    assert abs(s[4] - r[7]) <= 3e-7 + 1e-5 * r[7]  # sigma: order statistic of 3e5 residuals (dense: noise averages)
    assert se3_log_err(T1[0].cpu().numpy(), r[0].numpy()) < 1e-5
    np.testing.assert_array_equal(s[8:24].reshape(4, 4), case["T_init"][0].numpy())


def test_exact_order_statistic_bitwise():
    """Inputs on which both sides compute every residual EXACTLY (identity intrinsics and pose, points on integer
    pixels, intensities on a 1/256 grid, a = b = 0): sigma must equal 1.4826f * torch.median(|r|) bit for bit, for
    the candidate-list finish and for the pure histogram-narrowing path, with ties and a ragged point count."""
    from como_b200 import _lib

    g = torch.Generator().manual_seed(5)
    h, w = 97, 131
    img = torch.randint(0, 256, (1, 1, h, w), generator=g).float() / 256.0
    ys, xs = torch.meshgrid(torch.arange(1, h - 2), torch.arange(1, w - 2), indexing="ij")
    P = torch.stack((xs.reshape(-1).float(), ys.reshape(-1).float(), torch.ones(xs.numel())), 1)[None]
    n = P.shape[1]
    vals = (torch.randint(0, 256, (1, n, 1), generator=g).float() / 256.0)
    J = torch.randn(1, n, 1, 8, generator=g)
    mask = torch.rand(1, n, generator=g) > 0.1
    lv = dict(vals=[vals], P=[P], dI_dT=[J], mask=[mask], K=[torch.eye(3)], img=[img])
    r = img[0, 0][ys.reshape(-1), xs.reshape(-1)] - vals.reshape(-1)
    ref = (torch.tensor(1.4826, dtype=torch.float32) * torch.median(r[mask.reshape(-1)].abs())).item()
    for cap in (2048, 0):
        _lib.track_debug_candidate_cap(cap)
        try:
            T, aff, st = cuda_track(lv, torch.eye(4)[None], torch.zeros(1, 2, 1), dict(TERM, max_iter=1))
        finally:
            _lib.track_debug_candidate_cap(2048)
        assert int(st[0, 5]) == int(mask.sum())
        assert float(st[0, 4]) == ref, (cap, float(st[0, 4]), ref)


def test_run_to_run_bitwise_deterministic(full_case):
    case, T, aff, stats = full_case
    T2, aff2, stats2 = cuda_track(case, case["T_init"], case["aff_init"], TERM)
    assert torch.equal(T, T2) and torch.equal(aff, aff2) and torch.equal(stats, stats2)


def test_median_paths_agree_bitwise(full_case):
    """The candidate-list finish and the pure histogram-narrowing path return the same exact order statistic."""
    from como_b200 import _lib

    case, T, aff, stats = full_case
    _lib.track_debug_candidate_cap(0)
    try:
        T2, aff2, stats2 = cuda_track(case, case["T_init"], case["aff_init"], TERM)
    finally:
        _lib.track_debug_candidate_cap(2048)
    assert torch.equal(stats[:, 4], stats2[:, 4])
    assert torch.equal(T, T2) and torch.equal(stats, stats2)


@pytest.mark.parametrize("G", [1, 3, 8, 40])
def test_group_sizes_replay(G, monkeypatch):
    """Different CTA-group sizes (slice boundaries, shared-memory vs L2 residual residency, single-CTA path):
    every iteration still matches the oracle from identical inputs."""
    from como_b200 import synth

    monkeypatch.setenv("COMO_B200_TRACK_G", str(G))
    case = synth.make_tracking_case(240, 320, 3, seed=3, cell=8)
    T, aff, stats = cuda_track(case, case["T_init"], case["aff_init"], TERM)
    replay_check(case, stats, T[0].cpu().numpy(), aff.cpu().numpy().ravel())


def test_two_ctas_per_sm_mode_replay(monkeypatch):
    from como_b200 import synth

    monkeypatch.setenv("COMO_B200_TRACK_OCC", "2")
    case = synth.make_tracking_case(480, 640, 4, seed=5)
    T, aff, stats = cuda_track(case, case["T_init"], case["aff_init"], TERM)
    replay_check(case, stats, T[0].cpu().numpy(), aff.cpu().numpy().ravel())


@pytest.mark.parametrize("G", [1, 4])
def test_four_ctas_per_sm_kernel_replay(G, monkeypatch):
    """The 96-register build of the kernel (batches of more than 3 x 148 sequences use it); G = 1: one CTA streams a whole
    problem through the packed tiles' residual slots (no shared-memory residency)."""
    from como_b200 import synth

    monkeypatch.setenv("COMO_B200_TRACK_OCC", "4")
    monkeypatch.setenv("COMO_B200_TRACK_G", str(G))
    monkeypatch.setenv("COMO_B200_TRACK_RCAP", "0")
    case = synth.make_tracking_case(240, 320, 3, seed=7, cell=8)
    T, aff, stats = cuda_track(case, case["T_init"], case["aff_init"], TERM)
    replay_check(case, stats, T[0].cpu().numpy(), aff.cpu().numpy().ravel())


def test_far_taps_on_the_image_border_stay_in_bounds():
    """A coordinate a hair below h-1 / w-1 can round onto it in the reference's fp32 normalise / unnormalise round trip:
    the far row / column then has weight 0 (grid_sample treats it as padding) and must not be fetched -- it lies past
    the image.  The image sits in the last bytes of its allocation so that such a fetch faults."""
    from como_b200 import synth

    case = synth.make_tracking_case(120, 160, 1, seed=11, cell=8)
    h, w = case["img"][0].shape[-2:]
    K = case["K"][0].double()
    Ti = torch.linalg.inv(case["T_init"][0].double())
    n = case["P"][0].reshape(-1, 3).shape[0]
    # points that project, under the initial pose, to y = h-1-eps (first half) or x = w-1-eps (second half)
    eps = torch.logspace(-7, -2, n, dtype=torch.float64)
    xs = torch.linspace(1.5, w - 2.5, n, dtype=torch.float64)
    ys = torch.linspace(1.5, h - 2.5, n, dtype=torch.float64)
    half = torch.arange(n) < n // 2
    px = torch.where(half, xs, (w - 1) - eps)
    py = torch.where(half, (h - 1) - eps, ys)
    z = torch.full((n,), 2.0, dtype=torch.float64)
    Pc = torch.stack(((px - K[0, 2]) / K[0, 0] * z, (py - K[1, 2]) / K[1, 1] * z, z), -1)
    case["P"][0] = ((Ti[:3, :3] @ Pc.T).T + Ti[:3, 3]).float().reshape(case["P"][0].shape)
    case["mask"][0] = torch.ones_like(case["mask"][0])
    buf = torch.empty(1 << 20, dtype=torch.float32, device="cuda")   # a whole allocator block: nothing mapped behind it
    img = buf[-h * w:].view(1, 1, h, w)
    img.copy_(case["img"][0])
    from como_b200.odom.frontend.photo_tracking import photo_tracking_pyr

    c = lambda x: x.cuda()
    T, aff, stats = photo_tracking_pyr(c(case["T_init"]), c(case["aff_init"]), [c(case["vals"][0])], [c(case["P"][0])],
                                       [c(case["dI_dT"][0])], [c(case["mask"][0])], [case["K"][0]], [img], 0.1,
                                       dict(TERM, max_iter=1), return_stats=True)
    torch.cuda.synchronize()
    assert stats.shape[0] == 1 and bool(torch.isfinite(stats[0, :6]).all())
    # most points sit within a few ulp of the validity threshold here, so only the aggregate is compared: the valid
    # count within the number of such points, the robust scale and error within a percent
    st = stats.cpu().numpy()
    m = case["mask"][0].reshape(-1)
    d = {}
    out = TO.tracking_iter(torch.from_numpy(st[0, 8:24].reshape(4, 4).copy()), torch.from_numpy(st[0, 24:26].copy()),
                           case["vals"][0].reshape(-1)[m], case["P"][0].reshape(-1, 3)[m],
                           case["dI_dT"][0].reshape(-1, 8)[m], case["K"][0], case["img"][0][0, 0], detail=d)
    mse, sigma, nvalid = out[3], out[7], out[8]
    margin = torch.minimum((d["x"] - (w - 1)).abs(), (d["y"] - (h - 1)).abs())
    nb = int(((margin <= 4 * 2.0 ** -23 * max(w, h)) & (d["z"] > 0)).sum())
    assert nb > 1000 and abs(st[0, 5] - nvalid) <= nb
    assert abs(st[0, 4] - sigma) <= 1e-2 * sigma and abs(st[0, 1] - mse) <= 2e-2 * mse


def test_all_zero_residuals_single_bin():
    """Every |r| is exactly 0 (black frames): one histogram bin holds every pixel -> the narrowing passes run down to
    a single key; sigma = 0 and the weights are NaN exactly as in the reference (r / 0).  Must terminate, not hang."""
    from como_b200 import synth

    case = synth.make_tracking_case(120, 160, 2, seed=4, cell=8)
    case["img"] = [torch.zeros_like(i) for i in case["img"]]
    case["vals"] = [torch.zeros_like(v) for v in case["vals"]]
    T, aff, stats = cuda_track(case, case["T_init"], case["aff_init"], dict(TERM, max_iter=2))
    st = stats.cpu().numpy()
    # the first update is NaN (r / 0); from then on no pixel is valid and every level stops after one iteration
    assert st.shape[0] == 3 and st[0, 4] == 0.0 and st[0, 5] > 0 and np.all(st[1:, 5] == 0)


@pytest.mark.parametrize("name", ["track_80x60_l3", "track_160x120_l4"])
def test_golden_inputs_replay(golden_dir, name):
    """Reference-generated operands (ragged level sizes: 300, 1200, 4800 points -> bulk-copy tails)."""
    g = load(golden_dir, name)
    lv = golden_levels(g)
    T, aff, stats = cuda_track(lv, torch.from_numpy(g["T_init"]), torch.from_numpy(g["aff_init"]),
                               dict(TERM, max_iter=int(g["max_iter"])))
    replay_check(lv, stats, T[0].cpu().numpy(), aff.cpu().numpy().ravel(), strict=False)


def test_rgb_every_iteration_vs_oracle(golden_dir):
    """C = 3: every iteration the kernel ran, replayed by the channel-aware oracle from the kernel's own iterate:
    valid pixels +-2, sigma (median over all valid (pixel, channel) residuals) and the robust error 1e-4 relative,
    update within 1e-4."""
    g = load(golden_dir, "track_80x60_l3_rgb")
    lv = golden_levels(g)
    T, aff, stats = cuda_track(lv, torch.from_numpy(g["T_init"]), torch.from_numpy(g["aff_init"]),
                               dict(TERM, max_iter=int(g["max_iter"])))
    st = stats.cpu().numpy()
    assert st.shape[0] == len(g["trace_mse"])
    for i in range(st.shape[0]):
        l = int(st[i, 0])
        m = lv["mask"][l].reshape(-1)
        out = TO.tracking_iter_multi(torch.from_numpy(st[i, 8:24].reshape(4, 4).copy()), torch.from_numpy(st[i, 24:26].copy()),
                                     lv["vals"][l].reshape(-1, 3)[m], lv["P"][l].reshape(-1, 3)[m],
                                     lv["dI_dT"][l].reshape(-1, 3, 8)[m], lv["K"][l], lv["img"][l][0])
        Tn, affn, delta, mse, gn, H, gr, sigma, nvalid = out
        assert abs(st[i, 5] - nvalid) <= 2
        q = 2.0 / max(3 * nvalid, 1)
        assert abs(st[i, 4] - sigma) <= 2e-6 + (1e-5 + q) * sigma
        assert abs(st[i, 1] - mse) <= (2e-4 + 4 * q) * mse
        T_next = st[i + 1, 8:24].reshape(4, 4) if i + 1 < st.shape[0] else T[0].cpu().numpy()
        assert se3_log_err(T_next, Tn.numpy()) < 1e-4


@pytest.mark.parametrize("name", ["track_80x60_l3", "track_80x60_l3_rgb"])
def test_precalc_jacobians_vs_golden(golden_dir, name):
    from como_b200.odom.frontend.photo_tracking import precalc_jacobians

    g = load(golden_dir, name)
    for l in range(int(g["num_levels"])):
        c = lambda k: torch.from_numpy(g[k]).cuda()
        J = precalc_jacobians(c(f"grads_{l}"), c(f"P_{l}"), c(f"vals_{l}"), torch.from_numpy(g[f"K_{l}"]))
        np.testing.assert_allclose(J.cpu().numpy(), g[f"dI_dT_{l}"], rtol=2e-5, atol=1e-6)


def test_all_points_masked_or_invalid_does_not_hang():
    from como_b200 import synth

    case = synth.make_tracking_case(60, 80, 2, seed=1, cell=4)
    case["mask"] = [torch.zeros_like(m) for m in case["mask"]]
    T, aff, stats = cuda_track(case, case["T_init"], case["aff_init"], TERM)
    assert stats.shape[0] == 2  # one (degenerate) iteration per level, flagged done
    assert torch.isnan(T).any() or torch.isfinite(T).all()


def test_empty_level_and_ragged_tile_counts():
    """A pyramid whose coarsest level holds no points at all (n = 0: skipped, no iteration recorded) and whose other
    levels have point counts that are not multiples of the 512-pixel tile (padding slots carry NaN points); the rest of
    the run still replays against the oracle."""
    from como_b200 import synth

    case = synth.make_tracking_case(120, 160, 3, seed=4, cell=8)
    keep = {1: 4801 - 512 - 7, 2: 19200 - 1}        # ragged: 4282 = 8 tiles + 186, 19199 = 37 tiles + 255
    for l in range(3):
        n = 0 if l == 0 else keep[l]
        case["vals"][l] = case["vals"][l][:, :n].contiguous()
        case["P"][l] = case["P"][l][:, :n].contiguous()
        case["dI_dT"][l] = case["dI_dT"][l][:, :n].contiguous()
        case["mask"][l] = case["mask"][l][:, :n].contiguous()
    T, aff, stats = cuda_track(case, case["T_init"], case["aff_init"], TERM)
    torch.cuda.synchronize()
    assert stats.shape[0] >= 2 and int(stats[:, 0].min()) == 1          # level 0 took no iteration
    sub = {k: (v[1:] if isinstance(v, list) else v) for k, v in case.items()}
    st = stats.clone()
    st[:, 0] -= 1                                                       # level indices of the two-level sub-problem
    replay_check(sub, st, T[0].cpu().numpy(), aff.cpu().numpy().ravel(), strict=False)


def test_config1_shape_320x240_one_iteration_vs_oracle():
    """BASELINE config 1: 2-frame 320x240 photometric tracking, 3 levels, 1 GN iteration per level."""
    from como_b200 import synth

    case = synth.make_tracking_case(240, 320, 3, seed=2, cell=8)
    term = dict(TERM, max_iter=1)
    T, aff, stats = cuda_track(case, case["T_init"], case["aff_init"], term)
    To, affo, trace = TO.track_pyr(case["T_init"], case["aff_init"], case["vals"], case["P"], case["dI_dT"],
                                   case["mask"], case["K"], case["img"], term)
    assert stats.shape[0] == len(trace) == 3
    st = stats.cpu().numpy()
    assert abs(st[0, 1] - trace[0]["mse"]) <= (1e-4 + 2.0 / trace[0]["nvalid"]) * trace[0]["mse"]
    assert se3_log_err(T[0].cpu().numpy(), To.numpy()) < 1e-4
    np.testing.assert_allclose(aff.cpu().numpy().ravel(), affo.numpy().ravel(), atol=1e-4)


def test_batched_independent_sequences_match_single_launches():
    """BASELINE config 5 shape: independent sequences batched in one launch give the per-sequence results."""
    from como_b200 import synth
    from como_b200.odom.frontend.photo_tracking import photo_tracking_pyr_batch

    cases = [synth.make_tracking_case(120, 160, 3, seed=s, cell=8, device="cuda") for s in range(5)]
    probs = [(c["vals"], c["P"], c["dI_dT"], c["mask"], c["K"], c["img"]) for c in cases]
    T0 = torch.cat([c["T_init"] for c in cases])
    a0 = torch.cat([c["aff_init"] for c in cases])
    Tb, ab, nit = photo_tracking_pyr_batch(T0, a0, probs, TERM)
    for i, c in enumerate(cases):
        T1, a1, st = cuda_track({k: v for k, v in c.items() if isinstance(v, list)}, c["T_init"], c["aff_init"], TERM)
        assert int(nit[i]) == st.shape[0]
        # different CTA counts change the summation order only
        assert se3_log_err(Tb[i].cpu().numpy(), T1[0].cpu().numpy()) < 1e-5
