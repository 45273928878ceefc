"""Which (reference keyframe, target frame) pairs get photometric factors.

Mirror of como/odom/backend/graph_pair_construction.py:5-182 for the configuration the reference ships
(`radius_thresh: 0.0`, `degrees_thresh: 0.0`, config/como.yml:40-41): consecutive keyframes in both
directions plus, for every one-way frame, the keyframe(s) that bracket it in time.  Pure host integer
logic -- the pair list is the sharding unit of the multi-GPU path and must be bit-exact.
"""


def get_forward_edges(B):
    return list(range(0, B - 1)), list(range(1, B))


def get_backward_edges(B):
    return list(range(1, B)), list(range(0, B - 1))


def get_one_way_temporal_neighbors(kf_timestamps, recent_timestamps):
    """Each one-way frame attaches to the keyframe behind it and, unless it is newer than the newest
    keyframe, to the keyframe ahead of it.  Timestamps are assumed ascending (as in the reference)."""
    nk, nr = len(kf_timestamps), len(recent_timestamps)
    kf_ids, ow_ids = [], []
    behind = -1
    while recent_timestamps[0] > kf_timestamps[behind + 1]:
        behind += 1
        if behind == nk - 1:
            break
    r = 0
    if behind < nk - 1:
        while r < nr:
            if recent_timestamps[r] > kf_timestamps[behind + 1]:
                behind += 1
            if behind >= nk - 1:
                break
            kf_ids.extend((behind, behind + 1))
            ow_ids.extend((r, r))
            r += 1
    for rr in range(r, nr):
        kf_ids.append(behind)
        ow_ids.append(rr)
    return kf_ids, ow_ids


def setup_photometric_pairs(poses, recent_poses, kf_timestamps, recent_timestamps, median_depths, cfg):
    """Same signature and return value as the reference (`poses`/`recent_poses` only supply the counts)."""
    if cfg.get("radius_thresh", 0.0) > 0.0 and cfg.get("degrees_thresh", 0.0) > 0.0:
        raise NotImplementedError("como_b200: radius-based pair construction is disabled in config/como.yml "
                                  "and not implemented")
    nk = int(poses.shape[0]) if hasattr(poses, "shape") else int(poses)
    nr = int(recent_poses.shape[0]) if (hasattr(recent_poses, "shape") and recent_poses.numel() > 0) else (
        int(recent_poses) if isinstance(recent_poses, int) else 0)
    rf, tf = get_forward_edges(nk)
    rb, tb = get_backward_edges(nk)
    if nr > 0:
        ow_kf, ow_t = get_one_way_temporal_neighbors(list(kf_timestamps), list(recent_timestamps))
    else:
        ow_kf, ow_t = [], []
    return rf + rb, tf + tb, ow_kf, ow_t
