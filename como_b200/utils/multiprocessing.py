"""Inter-process hand-off of keyframe / frame tuples: same surface as the reference's
como/utils/multiprocessing.py -- `TupleTensorQueue(device, dtype, maxsize)` with push / pop / pop_until_latest /
qsize / empty / full, `transfer_data`, `release_data`, `init_gpu` -- with a different transport.

The reference pickles every tensor of every message through an mp.Queue: for CUDA tensors that is a fresh allocation,
a fresh cudaIpcGetMemHandle on the producer and a cudaIpcOpenMemHandle on the consumer per tensor per message
(multiprocessing.py:45-51), for dense depth / keyframe images at frame rate (Mapping.get_kf_ref_data,
como/odom/Mapping.py:498-511).  Here the producer owns a ring of persistent slots on the consumer's device that is
shared with the consumer process ONCE; a message is
    one kernel launch   all tensors of the tuple packed into the next free slot, converted to the consumer's dtype on
                        the way (csrc/handoff.cu: como_b200_handoff_pack) -- `transfer_data` fused into the copy
    one event record    an interprocess CUDA event per slot, shared once like the ring; the consumer's stream waits on
                        it (no host-side synchronisation between the processes)
    one small control record through the mp.Queue (slot, sequence number, shapes, the non-tensor items).
`pop` returns tensors that own their memory (one device-to-device copy out of the slot) unless zero_copy=True.  If the
consumer lags by more than the ring (unbounded "only want recent" queues), the message goes the reference's way.
For device "cpu" the ring lives in shared memory and there are no events (used by the CPU protocol test).
"""
import queue

import torch
import torch.multiprocessing as mp

_ALIGN = 256


def init_gpu(device):
    """Reference: multiprocessing.py:6-9 (warms the allocator of a freshly spawned process)."""
    for _ in range(4):
        t = torch.zeros((8, 192, 256), device="cpu").to(device)
        del t


def release_data(data):
    del data


def transfer_data(data, device, dtype):
    """Reference: multiprocessing.py:16-21."""
    return tuple(d.to(device=device, dtype=dtype, copy=False) if torch.is_tensor(d) else d for d in data)


def _nbytes(t, dtype):
    n = t.numel() * torch.empty((), dtype=dtype).element_size()
    return (n + _ALIGN - 1) // _ALIGN * _ALIGN


class TupleTensorQueue:
    def __init__(self, device, dtype, maxsize=0, slots=None, ctx=None):
        # no CUDA state here: the object is pickled into spawned processes (reference: Mapping.py:42).  CUDA needs the
        # "spawn" start method (the reference sets it globally in como_dataset.py); it is the default context here.
        ctx = ctx or mp.get_context("spawn")
        self.queue = ctx.Queue(maxsize=maxsize)
        self.device = device
        self.dtype = dtype
        self.num_slots = int(slots) if slots else max(3, (maxsize if maxsize > 0 else 6) + 2)
        self._popped = ctx.Value("q", 0)    # sequence number up to which the consumer is done with the slots
        self._seq = 0                       # producer: messages sent
        self._ring = None                   # producer: list of uint8 slot tensors on the consumer's device
        self._events = None
        self._slot_seq = None
        self._cap = 0
        self._cring = None                  # consumer: the same slots, opened once
        self._cevents = None
        self.stats = {"slot": 0, "fallback": 0, "rings": 0}

    # ------------------------------------------------------------------ queue surface
    def qsize(self):
        return self.queue.qsize()

    def empty(self):
        return self.queue.empty()

    def full(self):
        return self.queue.full()

    # ------------------------------------------------------------------ producer
    def _is_cuda(self):
        return torch.device(self.device).type == "cuda"

    def _make_ring(self, need):
        dev = torch.device(self.device)
        self._cap = max(int(need * 1.25), 1 << 20)
        if dev.type == "cuda":
            self._ring = [torch.empty(self._cap, dtype=torch.uint8, device=dev) for _ in range(self.num_slots)]
            self._events = [torch.cuda.Event(enable_timing=False, interprocess=True) for _ in range(self.num_slots)]
            for ev in self._events:
                ev.record(torch.cuda.current_stream(dev))
            handles = [ev.ipc_handle() for ev in self._events]
        else:
            self._ring = [torch.empty(self._cap, dtype=torch.uint8).share_memory_() for _ in range(self.num_slots)]
            self._events, handles = None, None
        self._slot_seq = [-1] * self.num_slots
        self.stats["rings"] += 1
        # shared ONCE: torch.multiprocessing turns the CUDA tensors into IPC handles here
        self.queue.put(("__ring__", self._ring, handles))

    def _free_slot(self):
        done = self._popped.value
        for s in range(self.num_slots):
            if self._slot_seq[s] < done:     # never used (-1) or already popped
                return s
        return None

    def _pack(self, slot_buf, items):
        """items: list of (tensor, offset).  Same device + supported dtypes: ONE launch; otherwise one copy_ each."""
        dev = torch.device(self.device)
        fused = dev.type == "cuda" and len(items) > 0
        if fused:
            from como_b200 import _lib

            fused = (self.dtype in (torch.float32, torch.float64)
                     and all(t.is_cuda and t.device == dev and t.dtype in _lib.PACK_DTYPES for t, _ in items))
        if fused:
            srcs = [t.contiguous() for t, _ in items]
            with torch.cuda.device(dev):
                for i in range(0, len(items), _lib.PACK_MAX_ITEMS):
                    chunk = list(zip(srcs[i:i + _lib.PACK_MAX_ITEMS], items[i:i + _lib.PACK_MAX_ITEMS]))
                    arr = (_lib.PackItem * len(chunk))()
                    for j, (src, (_, off)) in enumerate(chunk):
                        arr[j].src, arr[j].dst_offset_bytes, arr[j].count = src.data_ptr(), off, src.numel()
                        arr[j].src_dtype, arr[j].dst_dtype = _lib.PACK_DTYPES[src.dtype], _lib.PACK_DTYPES[self.dtype]
                    _lib.check(_lib.handoff_pack(arr, len(chunk), _lib.ptr(slot_buf), _lib.stream_ptr(dev)),
                               "como_b200_handoff_pack")
                for src in srcs:
                    src.record_stream(torch.cuda.current_stream(dev))
            return
        esz = torch.empty((), dtype=self.dtype).element_size()
        for t, off in items:
            slot_buf[off:off + t.numel() * esz].view(self.dtype).view(t.shape).copy_(t, non_blocking=True)
            if t.is_cuda and dev.type == "cuda" and t.device != dev:
                torch.cuda.current_stream(t.device).synchronize()   # cross-device copy is ordered on the source stream

    def push(self, data, block=True, timeout=None):
        tens = [(i, d) for i, d in enumerate(data) if torch.is_tensor(d)]
        if not tens:
            self.queue.put(("__plain__", tuple(data)), block=block, timeout=timeout)
            return
        need = sum(_nbytes(t, self.dtype) for _, t in tens)
        if self._ring is None or need > self._cap:
            self._make_ring(need)
        slot = self._free_slot()
        seq = self._seq
        self._seq += 1
        if slot is None:
            # consumer lags by a whole ring: the reference's transport for this message
            self.stats["fallback"] += 1
            self.queue.put(("__tensors__", seq, transfer_data(data, self.device, self.dtype)), block=block, timeout=timeout)
            return
        meta, items, off = [], [], 0
        for i, d in enumerate(data):
            if torch.is_tensor(d):
                meta.append(("t", off, tuple(d.shape)))
                items.append((d, off))
                off += _nbytes(d, self.dtype)
            else:
                meta.append(("o", d))
        self._pack(self._ring[slot], items)
        if self._events is not None:
            self._events[slot].record(torch.cuda.current_stream(torch.device(self.device)))
        self._slot_seq[slot] = seq
        self.stats["slot"] += 1
        self.queue.put(("__slot__", seq, slot, meta), block=block, timeout=timeout)

    # ------------------------------------------------------------------ consumer
    def _open(self, rec, zero_copy):
        kind = rec[0]
        if kind == "__plain__":
            return rec[1]
        if kind == "__tensors__":
            self._ack(rec[1])
            return rec[2]
        _, seq, slot, meta = rec
        dev = torch.device(self.device)
        buf = self._cring[slot]
        if self._cevents is not None:
            torch.cuda.current_stream(dev).wait_event(self._cevents[slot])
        esz = torch.empty((), dtype=self.dtype).element_size()
        out = []
        for m in meta:
            if m[0] == "o":
                out.append(m[1])
                continue
            _, off, shape = m
            n = 1
            for d in shape:
                n *= d
            v = buf[off:off + n * esz].view(self.dtype).view(shape)
            out.append(v if zero_copy else v.clone())
        if not zero_copy:
            if dev.type == "cuda":
                torch.cuda.current_stream(dev).synchronize()   # the copies out of the slot are done: it may be reused
            self._ack(seq)
        return tuple(out)

    def _ack(self, seq):
        with self._popped.get_lock():
            if seq + 1 > self._popped.value:
                self._popped.value = seq + 1

    def ack(self, n_messages_back=0):
        """zero_copy consumers: declare every message older than the last `n_messages_back` ones released."""
        with self._popped.get_lock():
            self._popped.value = max(self._popped.value, self._last_seq + 1 - n_messages_back)

    def _get(self, block, timeout):
        while True:
            rec = self.queue.get(block=block, timeout=timeout)
            if rec[0] == "__ring__":
                self._cring = rec[1]
                if rec[2] is not None:
                    dev = torch.device(self.device)
                    self._cevents = [torch.cuda.Event.from_ipc_handle(dev, h) for h in rec[2]]
                continue
            if rec[0] in ("__slot__", "__tensors__"):
                self._last_seq = rec[1]
            return rec

    def close(self):
        """Drop this process's references to the shared ring (consumer: before it exits, so the producer's memory is not
        released under it; producer: after the consumer has gone)."""
        self._cring = self._cevents = None
        self._ring = self._events = None
        if torch.device(self.device).type == "cuda" and torch.cuda.is_initialized():
            torch.cuda.ipc_collect()

    def pop(self, block=True, timeout=None, zero_copy=False):
        try:
            return self._open(self._get(block, timeout), zero_copy)
        except queue.Empty:
            return None

    def pop_until_latest(self, block=True, timeout=None, zero_copy=False):
        latest = None
        block_loop = block
        while True:
            try:
                rec = self._get(block_loop, timeout)
                if latest is not None and latest[0] in ("__slot__", "__tensors__"):
                    self._ack(latest[1])          # skipped without ever touching its slot
                latest = rec
                block_loop = False                # already got one message: no more blocking
            except queue.Empty:
                break
        return None if latest is None else self._open(latest, zero_copy)
