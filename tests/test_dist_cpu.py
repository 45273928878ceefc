"""N>1 host logic on CPU (gloo, world_size 2, 127.0.0.1): the pair partition used to shard a BA window
over GPUs, and the algebra of the sharded iteration (global robust scale + all-reduced normal equations)
checked with the oracle."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, golden, q):
    import sys
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from como_b200.odom.backend.graph_pair_construction import setup_photometric_pairs
        from como_b200.odom.mapping_core import partition_pairs
        from oracle import ba_oracle as BO

        # 1. partition: disjoint, complete, order preserving, batches by GLOBAL pair index
        K, R, bs = 7, 9, 4
        kts = [1.0 + k for k in range(K)]
        rts = sorted(1.0 + (j % (K - 1)) + 0.5 + 0.001 * j for j in range(R))
        ref, tgt, owk, owt = setup_photometric_pairs(K, R, kts, rts, None, {})
        mine = partition_pairs(ref, tgt, owk, owt, K, bs, rank, world)
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        full = partition_pairs(ref, tgt, owk, owt, K, bs, 0, 1)
        merged = sorted((b, r, t) for g in gathered for r, t, b in zip(g[0], g[1], g[2]))
        assert merged == sorted((b, r, t) for r, t, b in zip(full[0], full[1], full[2]))
        assert sum(len(g[0]) for g in gathered) == len(full[0])
        assert all(r % world == rank for r in mine[0])
        assert all(g[3] == full[3] for g in gathered)

        # 2. sharded normal equations: sum over ranks == single-process result (oracle, fp64)
        g = np.load(golden)
        cfg = BO.cfg_from_golden(g)
        part = BO.iterate(BO.state_from_golden(g), cfg, rank=rank, world=world, photo_only=True)
        Hs, gs = part["H_photo"].clone(), part["g_photo"].clone()
        e = torch.tensor([part["photo_err"]], dtype=torch.float64)
        dist.all_reduce(Hs)
        dist.all_reduce(gs)
        dist.all_reduce(e)
        ref_full = BO.iterate(BO.state_from_golden(g), cfg, photo_only=True)
        assert float((Hs - ref_full["H_photo"]).abs().max()) <= 1e-9 * float(ref_full["H_photo"].abs().max())
        assert float((gs - ref_full["g_photo"]).abs().max()) <= 1e-9 * float(ref_full["g_photo"].abs().max())
        assert abs(float(e) - ref_full["photo_err"]) <= 1e-10 * ref_full["photo_err"]
        assert part["sigmas"] == ref_full["sigmas"]  # the robust scale is global, not per rank
        q.put((rank, "ok"))
    except Exception as ex:  # pragma: no cover
        import traceback
        q.put((rank, "FAIL " + traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_sharded_ba_host_logic_gloo_world2(golden_dir):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    golden = os.path.join(golden_dir, "ba_k4_notfull.npz")
    procs = [ctx.Process(target=_worker, args=(r, world, port, golden, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for r, msg in res:
        assert msg == "ok", f"rank {r}: {msg}"


@pytest.mark.parametrize("K,world", [(32, 2), (32, 4), (32, 8), (8, 3), (5, 8), (1, 2)])
def test_store_vars_keyframe_ranges_partition_the_window(K, world):
    """ShardComm.kf_range: the contiguous keyframe ranges the ranks stream in store_vars are disjoint, ordered and
    cover [0, K) -- also when K is not a multiple of the world size or smaller than it (empty ranges at the end)."""
    from como_b200.odom.mapping_core import ShardComm

    ranges = [ShardComm(world, r, torch.device("cpu")).kf_range(K) for r in range(world)]
    covered = []
    for lo, hi in ranges:
        assert 0 <= lo <= hi <= K
        covered += list(range(lo, hi))
    assert covered == list(range(K))
    sizes = [hi - lo for lo, hi in ranges]
    assert max(sizes) - min(s for s in sizes if s > 0) <= max(sizes)   # contiguous blocks of ceil(K / world)
    assert max(sizes) == -(-K // world)
