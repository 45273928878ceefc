"""Inter-process hand-off on the GPU: producer and consumer processes share ONE ring of device slots (CUDA IPC, opened
once) and one interprocess event per slot; a message is one pack launch (float64 -> float32 fused) + one event record."""
import time

import pytest
import torch
import torch.multiprocessing as mp

from handoff_workers import consumer, make_message

pytestmark = pytest.mark.gpu


def run(mode, n, slots, maxsize, delay=0.0, hw=(480, 640)):
    from como_b200.utils.multiprocessing import TupleTensorQueue

    ctx = mp.get_context("spawn")
    q = TupleTensorQueue("cuda:0", torch.float32, maxsize=maxsize, slots=slots)
    out = ctx.Queue()
    got = ctx.Event()
    p = ctx.Process(target=consumer, args=(q, n, "float32", mode, delay, out, got))
    p.start()
    msgs = [make_message(k, "cuda:0") for k in range(n)]
    torch.cuda.synchronize()
    t0 = time.time()
    for m in msgs:
        q.push(m)
    assert got.wait(timeout=180)
    q.push(("end",))
    res = out.get(timeout=180)
    dt = time.time() - t0
    p.join(timeout=60)
    assert res[0] == "ok", res[1]
    q.close()
    return res[1], q.stats, dt


def test_two_process_fifo_on_device():
    seen, stats, dt = run("all", 16, slots=4, maxsize=2)
    assert seen == list(range(16))
    assert stats["rings"] == 1 and stats["slot"] >= 12


def test_two_process_latest_wins_on_device():
    seen, stats, dt = run("latest", 7, slots=3, maxsize=0, delay=2.0)
    assert seen == [6]
    assert stats["slot"] == 3 and stats["fallback"] == 4


def test_pack_kernel_converts_like_transfer_data():
    from como_b200 import _lib

    src = [torch.rand(1000, dtype=torch.float64, device="cuda"), torch.arange(33, dtype=torch.int64, device="cuda"),
           (torch.arange(17, device="cuda") % 3 == 0), torch.rand(5, 7, device="cuda")]
    offs, off = [], 0
    for t in src:
        offs.append(off)
        off += (t.numel() * 4 + 255) // 256 * 256
    slot = torch.zeros(off, dtype=torch.uint8, device="cuda")
    arr = (_lib.PackItem * len(src))()
    for j, t in enumerate(src):
        arr[j].src, arr[j].dst_offset_bytes, arr[j].count = t.data_ptr(), offs[j], t.numel()
        arr[j].src_dtype, arr[j].dst_dtype = _lib.PACK_DTYPES[t.dtype], 0
    _lib.check(_lib.handoff_pack(arr, len(src), _lib.ptr(slot), _lib.stream_ptr()), "pack")
    for t, o in zip(src, offs):
        got = slot[o:o + t.numel() * 4].view(torch.float32).view(t.shape)
        assert torch.equal(got, t.to(torch.float32))
