"""Inter-process hand-off, protocol test on CPU with two processes: the slot ring in shared memory,
the control records, the dtype conversion of transfer_data, FIFO and `latest wins` semantics, and the fallback to the
reference's transport when the consumer lags by more than the ring."""
import pytest
import torch
import torch.multiprocessing as mp

from handoff_workers import consumer, make_message


def run(mode, n, slots, maxsize, delay=0.0, dtype=torch.float32):
    from como_b200.utils.multiprocessing import TupleTensorQueue

    ctx = mp.get_context("spawn")
    q = TupleTensorQueue("cpu", dtype, maxsize=maxsize, slots=slots)
    out = ctx.Queue()
    got = ctx.Event()
    p = ctx.Process(target=consumer, args=(q, n, str(dtype).split(".")[1], mode, delay, out, got))
    p.start()
    for k in range(n):
        q.push(make_message(k, "cpu"))
    assert got.wait(timeout=120)
    q.push(("end",))
    res = out.get(timeout=120)
    p.join(timeout=30)
    assert res[0] == "ok", res[1]
    q.close()
    return res[1], q.stats


def test_fifo_all_messages_converted_and_intact():
    seen, stats = run("all", 12, slots=4, maxsize=2)
    assert seen == list(range(12))
    assert stats["rings"] == 1 and stats["slot"] + stats["fallback"] == 12 and stats["slot"] >= 8


def test_zero_copy_views_with_explicit_ack():
    seen, stats = run("zero_copy", 8, slots=3, maxsize=1, dtype=torch.float64)
    assert seen == list(range(8))


def test_latest_wins_and_lagging_consumer_falls_back():
    # unbounded queue, consumer asleep while 9 messages arrive in a ring of 3: the ring fills, the rest goes the
    # reference's way; pop_until_latest hands out only the newest one
    seen, stats = run("latest", 9, slots=3, maxsize=0, delay=1.5)
    assert seen == [8]
    assert stats["slot"] == 3 and stats["fallback"] == 6
