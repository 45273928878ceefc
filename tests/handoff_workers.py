"""Worker processes of the hand-off tests (spawned: must be importable)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402


def make_message(k, device, dtype=torch.float64, hw=(48, 64)):
    g = torch.Generator().manual_seed(100 + k)
    rgb = torch.rand((2, 3) + hw, generator=g, dtype=dtype)
    pose = torch.eye(4, dtype=dtype)[None].repeat(2, 1, 1) * (k + 1)
    idx = torch.arange(7, dtype=torch.int64) + k
    mask = (torch.arange(9) % 2 == (k % 2))
    return ([1.0 + k, 2.0 + k], rgb.to(device), pose.to(device), "keyframe" if k % 2 else "one-way", idx.to(device), mask.to(device))


def check_message(k, msg, dtype):
    ref = make_message(k, "cpu")
    assert msg[0] == ref[0] and msg[3] == ref[3], (k, msg[0], msg[3])
    for i in (1, 2, 4, 5):
        got = msg[i].cpu()
        assert got.dtype == dtype, (i, got.dtype)                  # transfer_data converts every tensor to the queue dtype
        assert tuple(got.shape) == tuple(ref[i].shape)
        assert torch.equal(got, ref[i].to(dtype)), (k, i)


def consumer(q, n_msgs, dtype_name, mode, delay, out, got=None):
    try:
        dtype = getattr(torch, dtype_name)
        if torch.device(q.device).type == "cuda":
            torch.cuda.set_device(torch.device(q.device))
        seen = []
        if mode == "all":
            while len(seen) < n_msgs:
                m = q.pop(block=True, timeout=30)
                assert m is not None, "timeout"
                check_message(len(seen), m, dtype)
                seen.append(len(seen))
                if delay:
                    time.sleep(delay)
        elif mode == "zero_copy":
            for k in range(n_msgs):
                m = q.pop(block=True, timeout=30, zero_copy=True)
                check_message(k, m, dtype)
                q.ack()
                seen.append(k)
        else:  # latest: wait until the producer has queued everything, then take the newest only
            time.sleep(delay)
            m = q.pop_until_latest(block=True, timeout=30)
            k = int(round(m[0][0] - 1.0))
            check_message(k, m, dtype)
            seen.append(k)
        if got is not None:
            got.set()          # the producer sends "end" only now ("latest wins" would otherwise return it)
        end = q.pop(block=True, timeout=30)
        assert end == ("end",), end
        del m
        q.close()
        out.put(("ok", seen))
    except Exception:
        import traceback
        out.put(("FAIL", traceback.format_exc()))
