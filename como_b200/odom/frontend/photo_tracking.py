"""Drop-in replacements for the reference's tracking operators
(como/odom/frontend/photo_tracking.py:10-42 `photo_tracking_pyr`, :46-74 `precalc_jacobians`):
same names, argument order and meaning, same return values -- computed by the sm_100a kernels in
csrc/track.cu through the C ABI.  Tensors must live on a CUDA device; there is no CPU path.
"""
import ctypes as C
import weakref

import torch

from como_b200 import _lib

_ws_cache = {}
_k9_cache = {}
last_num_iters = None


def _k9_list(K):
    """3x3 intrinsics -> 9 python floats.  Device tensors are read back once and cached per tensor OBJECT
    (weak reference + version counter; never by address -- the caching allocator reuses addresses):
    a device->host copy per level and call would otherwise dominate the launch."""
    if not isinstance(K, torch.Tensor):
        return [float(v) for row in K for v in (row if hasattr(row, "__len__") else [row])]
    if not K.is_cuda:
        return K.detach().to(torch.float32).reshape(-1).tolist()
    key = id(K)
    ent = _k9_cache.get(key)
    if ent is not None and ent[0]() is K and ent[1] == K._version:
        return ent[2]
    v = K.detach().to("cpu", torch.float32).reshape(-1).tolist()
    _k9_cache[key] = (weakref.ref(K, lambda _r, key=key: _k9_cache.pop(key, None)), K._version, v)
    return v


def _workspace(nbytes, device):
    key = (device.index if device.index is not None else torch.cuda.current_device())
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
        _ws_cache[key] = ws
    return ws


def _aligned(t):
    """contiguous, and 16-byte aligned (the kernel streams the operands with bulk async copies)"""
    t = t.contiguous()
    return t if t.data_ptr() % 16 == 0 else t.clone(memory_format=torch.contiguous_format)


_pack_cache = {}


def _level_structs(vals_i, Pi, dI_dT, masks, intrinsics, img_j, keep, used=None):
    """Level descriptors of one problem.  The keyframe-side operands are re-laid out into the tracker's tile format
    (como_b200_track_pack) once per set of operand tensors: cached on identity + version counters of (vals, P, J, mask),
    dropped when the Jacobian tensor dies.  `used`: packs already claimed by other problems of the same launch -- the
    kernel keeps its residuals inside the pack, so a problem listed twice gets a private copy."""
    num_levels = len(vals_i)
    arr = (_lib.TrackLevel * num_levels)()
    max_n = 0
    for l in range(num_levels):
        v0, P0, J0, m0 = vals_i[l], Pi[l], dI_dT[l], masks[l]
        nch = int(v0.shape[-1])                       # image channels: 1 (gray) or 3 (rgb)
        img = img_j[l]
        if img.shape[0] != 1 or img.shape[1] != nch:
            raise RuntimeError(f"como_b200 tracking expects img_j[l] of shape (1,{nch},h,w), got {tuple(img.shape)}")
        img = img.contiguous().float()
        Kl = _k9_list(intrinsics[l])
        n = v0.numel() // max(nch, 1)
        a = arr[l]
        a.img = img.data_ptr()
        a.n, a.w, a.h, a.c = n, img.shape[-1], img.shape[-2], nch
        max_n = max(max_n, n * nch)
        for k in range(9):
            a.K[k] = Kl[k]
        dev = v0.device
        key = tuple(id(t) for t in (v0, P0, J0, m0))
        ent = _pack_cache.get(key)
        if ent is not None and all(r() is t for r, t in zip(ent[0], (v0, P0, J0, m0))) and \
                ent[1] == tuple(t._version for t in (v0, P0, J0, m0)):
            pack = ent[2]
        else:
            v = v0.reshape(-1).float().contiguous()                 # (n, c) entries
            P = P0.reshape(-1, 3).float().contiguous()
            J = _aligned(J0.reshape(-1, 8).float())                 # (n, c, 8) -> one row per entry
            m = m0.reshape(-1).contiguous()
            if m.dtype != torch.uint8:
                m = m.view(torch.uint8) if m.dtype == torch.bool else m.to(torch.uint8)
            a.vals, a.P, a.J, a.mask = v.data_ptr(), P.data_ptr(), J.data_ptr(), m.data_ptr()
            pack = torch.empty(max(int(_lib.track_pack_bytes(n, nch)), 128), dtype=torch.uint8, device=dev)
            a.pack = pack.data_ptr()
            if n > 0:
                _lib.check(_lib.track_pack(C.byref(a), _lib.stream_ptr(dev)), "como_b200_track_pack")
            if len(_pack_cache) > 1024:
                _pack_cache.clear()
            refs = (weakref.ref(v0), weakref.ref(P0), weakref.ref(J0, lambda _r, key=key: _pack_cache.pop(key, None)),
                    weakref.ref(m0))
            _pack_cache[key] = (refs, tuple(t._version for t in (v0, P0, J0, m0)), pack)
            a.vals = a.P = a.J = a.mask = None   # temporaries: the launch reads only pack and img
        if used is not None:
            if pack.data_ptr() in used:
                pack = pack.clone()
            used.add(pack.data_ptr())
        a.pack = pack.data_ptr()
        keep += [pack, img]
    return arr, max_n


def photo_tracking_pyr(Tji_init, aff_init, vals_i, Pi, dI_dT, masks, intrinsics, img_j, photo_sigma,
                       term_criteria, return_stats=False):
    """Coarse-to-fine inverse-compositional tracking; inputs are per-level lists (coarsest first); vals (1,N,C),
    dI_dT (1,N,C,8), img_j (1,C,h,w) with C = 1 (tracking.color: gray) or 3 (rgb).

    Like the reference, `photo_sigma` is accepted and ignored (the scale is 1.4826 * median |r|,
    photo_tracking.py:132-138).  Returns (Tji (1,4,4), aff (1,2,1)) [, stats (iters, 32): see COMO_B200_TRACK_STAT_STRIDE].
    """
    dev = _lib.require_cuda(Tji_init, aff_init, *vals_i, *Pi, *dI_dT, *masks, *img_j)
    num_levels = len(vals_i)
    keep = []
    with torch.cuda.device(dev):
        arr, max_n = _level_structs(vals_i, Pi, dI_dT, masks, intrinsics, img_j, keep)
        T = Tji_init.detach().reshape(1, 4, 4).float().clone().contiguous()
        aff = aff_init.detach().reshape(1, 2).float().clone().contiguous()
        term = _lib.TrackTerm(int(term_criteria["max_iter"]), float(term_criteria["delta_norm"]),
                              float(term_criteria["rel_tol"]), float(term_criteria["grad_norm"]))
        cap = num_levels * term.max_iter
        stats = torch.zeros(cap, _lib.TRACK_STAT_STRIDE, dtype=torch.float32, device=dev) if return_stats else None
        nit = torch.zeros(1, dtype=torch.int32, device=dev)
        nbytes = _lib.track_workspace_bytes(max_n, 1)
        ws = _workspace(nbytes, dev)
        st = _lib.track_pyr(arr, num_levels, 1, C.byref(term), _lib.ptr(T), _lib.ptr(aff), _lib.ptr(stats),
                            _lib.ptr(nit), _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev))
        _lib.check(st, "como_b200_track_pyr")
    global last_num_iters
    last_num_iters = nit   # device tensor (no sync): GN iterations of the most recent call, for instrumentation
    Tji = T.to(Tji_init.dtype)
    aff_out = aff.reshape(1, 2, 1).to(aff_init.dtype)
    if return_stats:
        n = int(nit.item())
        return Tji, aff_out, stats[:n]
    return Tji, aff_out


def photo_tracking_pyr_batch(Tji_init, aff_init, problems, term_criteria, return_stats=False):
    """B independent frame-to-keyframe problems in ONE cooperative launch (BASELINE config 5:
    independent sequences batched per GPU; no reference counterpart -- the reference loops in Python).

    Tji_init (B,4,4), aff_init (B,2,1); problems: list of B tuples
    (vals_i, Pi, dI_dT, masks, intrinsics, img_j), each as for photo_tracking_pyr (same num_levels).
    Returns Tji (B,4,4), aff (B,2,1) [, stats (B, L*max_iter, 32), num_iters (B,)]."""
    B = len(problems)
    num_levels = len(problems[0][0])
    dev = _lib.require_cuda(Tji_init, aff_init)
    keep = []
    with torch.cuda.device(dev):
        arr = (_lib.TrackLevel * (B * num_levels))()
        max_n = 0
        used = set()
        for p, (vals_i, Pi, dI_dT, masks, intrinsics, img_j) in enumerate(problems):
            a, mn = _level_structs(vals_i, Pi, dI_dT, masks, intrinsics, img_j, keep, used)
            max_n = max(max_n, mn)
            for l in range(num_levels):
                arr[p * num_levels + l] = a[l]
        T = Tji_init.detach().reshape(B, 4, 4).float().clone().contiguous()
        aff = aff_init.detach().reshape(B, 2).float().clone().contiguous()
        term = _lib.TrackTerm(int(term_criteria["max_iter"]), float(term_criteria["delta_norm"]),
                              float(term_criteria["rel_tol"]), float(term_criteria["grad_norm"]))
        cap = num_levels * term.max_iter
        stats = torch.zeros(B, cap, _lib.TRACK_STAT_STRIDE, dtype=torch.float32, device=dev) if return_stats else None
        nit = torch.zeros(B, dtype=torch.int32, device=dev)
        ws = _workspace(_lib.track_workspace_bytes(max_n, B), dev)
        st = _lib.track_pyr(arr, num_levels, B, C.byref(term), _lib.ptr(T), _lib.ptr(aff), _lib.ptr(stats),
                            _lib.ptr(nit), _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev))
        _lib.check(st, "como_b200_track_pyr")
    if return_stats:
        return T, aff.reshape(B, 2, 1), stats, nit
    return T, aff.reshape(B, 2, 1), nit


class TrackBatchPlan:
    """Pre-validated launch description of B independent tracking problems: the per-level descriptors are built
    once (the operand tensors must stay alive and in place), `run` only copies the initial poses and launches."""

    def __init__(self, problems, term_criteria):
        self.B = len(problems)
        self.num_levels = len(problems[0][0])
        self.dev = _lib.require_cuda(*problems[0][0])
        self.keep = []
        with torch.cuda.device(self.dev):
            self.arr = (_lib.TrackLevel * (self.B * self.num_levels))()
            max_n = 0
            used = set()
            for p, (vals_i, Pi, dI_dT, masks, intrinsics, img_j) in enumerate(problems):
                a, mn = _level_structs(vals_i, Pi, dI_dT, masks, intrinsics, img_j, self.keep, used)
                max_n = max(max_n, mn)
                for l in range(self.num_levels):
                    self.arr[p * self.num_levels + l] = a[l]
            self.term = _lib.TrackTerm(int(term_criteria["max_iter"]), float(term_criteria["delta_norm"]),
                                       float(term_criteria["rel_tol"]), float(term_criteria["grad_norm"]))
            self.ws = torch.empty(int(_lib.track_workspace_bytes(max_n, self.B)), dtype=torch.uint8, device=self.dev)
            self.T = torch.empty(self.B, 4, 4, dtype=torch.float32, device=self.dev)
            self.aff = torch.empty(self.B, 2, dtype=torch.float32, device=self.dev)
            self.nit = torch.zeros(self.B, dtype=torch.int32, device=self.dev)

    def run(self, Tji_init, aff_init):
        with torch.cuda.device(self.dev):
            self.T.copy_(Tji_init.reshape(self.B, 4, 4))
            self.aff.copy_(aff_init.reshape(self.B, 2))
            st = _lib.track_pyr(self.arr, self.num_levels, self.B, C.byref(self.term), _lib.ptr(self.T), _lib.ptr(self.aff),
                                None, _lib.ptr(self.nit), _lib.ptr(self.ws), self.ws.numel(), _lib.stream_ptr(self.dev))
            _lib.check(st, "como_b200_track_pyr")
        return self.T, self.aff.reshape(self.B, 2, 1), self.nit


def precalc_jacobians(dI_dw, P, vals, intrinsics):
    """dI_dw (B,N,C,2), P (B,N,3), vals (B,N,C), intrinsics (3,3) -> (B,N,C,8); C = 1 (gray) or 3 (rgb)."""
    dev = _lib.require_cuda(dI_dw, P, vals)
    c = int(vals.shape[2])
    b, n, _ = P.shape
    g = dI_dw.reshape(-1, 2).contiguous().float()
    Pf = P.reshape(-1, 3).contiguous().float()
    v = vals.reshape(-1).contiguous().float()
    J = torch.empty(b * n * c, 8, dtype=torch.float32, device=dev)
    K = (C.c_float * 9)(*_k9_list(intrinsics))
    with torch.cuda.device(dev):
        st = _lib.precalc_jacobians(_lib.ptr(g), _lib.ptr(Pf), _lib.ptr(v), K, b * n, c, _lib.ptr(J), _lib.stream_ptr(dev))
    _lib.check(st, "como_b200_precalc_jacobians")
    return J.reshape(b, n, c, 8).to(dI_dw.dtype)
