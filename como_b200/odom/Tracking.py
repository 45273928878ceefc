"""`Tracking`: same class surface as the reference tracker (como/odom/Tracking.py:21-379) --
`Tracking(cfg, intrinsics, img_size)`, `.setup()`, `.update_kf_reference(kf_data)`, `.handle_frame(data)` and the
state attributes its callers read -- with every per-pixel stage on the sm_100a kernels (csrc/frontend.cu,
csrc/track.cu, csrc/select.cu).  tracking.color must be "gray"; tensors live on cfg["device"] (CUDA).
Do not create CUDA tensors in __init__ (objects are pickled into spawned processes, Mapping.py:42).
"""
import ctypes as C

import torch

from como_b200 import _lib
import como_b200.odom.frontend.photo_tracking as _pt
from como_b200.odom.frontend.photo_tracking import TrackBatchPlan, photo_tracking_pyr  # noqa: F401


def _dtype(name):
    if name == "float":
        return torch.float
    if name == "double":
        return torch.double
    raise ValueError("Cannot convert : " + name + " to tensor type.")


def _inv_se3(T):
    Ti = torch.zeros_like(T)
    Rt = T[..., :3, :3].transpose(-1, -2)
    Ti[..., :3, :3] = Rt
    Ti[..., :3, 3:4] = -(Rt @ T[..., :3, 3:4])
    Ti[..., 3, 3] = 1.0
    return Ti


class Tracking:
    def __init__(self, cfg, intrinsics, img_size):
        self.cfg = cfg
        self.device = cfg["device"]
        self.dtype = _dtype(cfg["dtype"])
        self.intrinsics = intrinsics
        self.img_size = img_size
        self.mapping_init = False

    def track(self, data):
        raise NotImplementedError

    # ------------------------------------------------------------------ set-up
    def setup(self):
        if self.cfg["color"] != "gray":
            raise NotImplementedError("como_b200 Tracking supports tracking.color: gray only")
        if self.dtype != torch.float:
            raise NotImplementedError("como_b200 Tracking runs in float32 (tracking.dtype: float)")
        self.dev = torch.device(self.device)   # no process-wide set_device: every launch below runs under a device guard
        self.intrinsics = self.intrinsics.to(device=self.dev, dtype=self.dtype)
        self.start_level = int(self.cfg["pyr"]["start_level"])
        self.end_level = int(self.cfg["pyr"]["end_level"])
        if self.start_level != 0:
            raise NotImplementedError("como_b200 Tracking expects pyr.start_level: 0")
        if self.cfg["pyr"]["depth_interp_mode"] != "nearest_neighbor":
            raise NotImplementedError("como_b200 Tracking supports depth_interp_mode: nearest_neighbor")
        self.num_levels = self.end_level - self.start_level
        # IntrinsicsPyramidModule incl. the reference's resize_intrinsics (scale added to the principal point)
        self.intrinsics_pyr = []
        for i in range(self.start_level, self.end_level):
            sc = 2.0 ** (-i)
            Tm = torch.tensor([[sc, 0, sc], [0, sc, sc], [0, 0, 1.0]], device=self.dev, dtype=self.dtype)
            self.intrinsics_pyr.insert(0, Tm @ self.intrinsics)
        self._K9 = [(C.c_float * 9)(*k.detach().cpu().reshape(-1).tolist()) for k in self.intrinsics_pyr]
        H, W = int(self.img_size[-2]), int(self.img_size[-1])
        self.level_sizes = [(H, W)]
        for _ in range(self.num_levels - 1):
            h, w = self.level_sizes[0]
            self.level_sizes.insert(0, ((h + 1) // 2, (w + 1) // 2))
        self.init_kf_vars()
        self.reset_one_way_vars()
        self.T_w_rec_last = None
        self._winner = torch.empty(H * W, dtype=torch.int32, device=self.dev)
        self._reproj = torch.empty(H * W, dtype=torch.float32, device=self.dev)
        self._seg = torch.tensor([0, H * W], dtype=torch.int64, device=self.dev)
        self._med_ws = torch.empty(int(_lib.median_workspace_bytes(1, 4)), dtype=torch.uint8, device=self.dev)
        self._med = torch.empty(1, dtype=torch.float32, device=self.dev)
        self._cnt = torch.empty(1, dtype=torch.int64, device=self.dev)
        # per-frame fast path: persistent target pyramid (so that the launch descriptors of the tracker stay valid from
        # one frame to the next), the cached launch plan, and ONE pinned read-back for the keyframe decisions
        self._img_bufs = [torch.empty((1, 1, h, w), dtype=torch.float32, device=self.dev) for (h, w) in self.level_sizes]
        self._plan = None
        self._host = torch.empty(18, dtype=torch.float32).pin_memory()

    def reset_one_way_vars(self):
        self.num_one_way_since_kf = 0
        self.last_one_way_empty_pixels = 0
        self.last_flow_rmse = 0.0
        self.last_flow_wo_rot_rmse = 0.0

    def init_kf_vars(self):
        self.T_curr_kf = torch.eye(4, device=self.dev, dtype=self.dtype).unsqueeze(0)
        self.aff_curr_kf = torch.zeros((1, 2, 1), device=self.dev, dtype=self.dtype)
        self.last_one_way_num_pixels = self.img_size[-1] * self.img_size[-2]
        self.last_kf_sent_ts = torch.zeros(1, device=self.dev, dtype=self.dtype)
        self.kf_received_ts = torch.zeros(1, device=self.dev, dtype=self.dtype)

    # ------------------------------------------------------------------ small pose / affine algebra
    def get_curr_world_pose(self):
        return self.T_w_kf @ _inv_se3(self.T_curr_kf)

    def get_curr_world_aff(self):
        a = self.aff_w_kf.clone()
        a[:, 0, :] += self.aff_curr_kf[:, 0, :]
        a[:, 1, :] += self.aff_curr_kf[:, 1, :] * torch.exp(self.aff_curr_kf[:, 0, :])
        return a

    # ------------------------------------------------------------------ images
    def _pyramid(self, rgb, with_grads, out=None):
        """One fused launch per image: gray + all pyramid levels (+ Scharr gradients laid out [I, gx, gy] per level,
        the layout kf_reference_level gathers from).  Pyramids deeper than 4 levels use the per-level kernels."""
        rgb = rgb.to(device=self.dev, dtype=torch.float32).contiguous()
        b = rgb.shape[0]
        nl = self.num_levels
        H, W = self.level_sizes[-1]
        ch = 3 if with_grads else 1
        if out is None:
            out = [torch.empty((b, ch, h, w), dtype=torch.float32, device=self.dev) for (h, w) in self.level_sizes]
        with torch.cuda.device(self.dev):
            stream = _lib.stream_ptr(self.dev)
            for i in range(b):
                ptrs = (C.c_void_p * nl)(*[o[i, 0].data_ptr() for o in out])
                if nl <= 4:
                    gx = (C.c_void_p * nl)(*[o[i, 1].data_ptr() for o in out]) if with_grads else None
                    gy = (C.c_void_p * nl)(*[o[i, 2].data_ptr() for o in out]) if with_grads else None
                    st = _lib.image_pyramid_fused(_lib.ptr(rgb[i]), H, W, nl, ptrs, gx, gy, stream)
                    _lib.check(st, "como_b200_image_pyramid_fused")
                else:
                    st = _lib.gray_pyramid(_lib.ptr(rgb[i]), H, W, nl, ptrs, stream)
                    _lib.check(st, "como_b200_gray_pyramid")
                    if with_grads:
                        for o in out:
                            st = _lib.image_gradients(_lib.ptr(o[i, 0]), o.shape[-2], o.shape[-1], _lib.ptr(o[i, 1]),
                                                      _lib.ptr(o[i, 2]), stream)
                            _lib.check(st, "como_b200_image_gradients")
        return out

    def prep_tracking_img(self, rgb):
        """rgb (1,3,H,W) -> list of (1,1,h,w) gray levels, coarsest first."""
        return self._pyramid(rgb, False)

    def get_img_gradients(self, img_pyr):
        """Gradients of an existing pyramid (reference signature, Tracking.py:96-102) -> list of (b,3,h,w) [I, gx, gy]."""
        res = []
        with torch.cuda.device(self.dev):
            for lvl in img_pyr:
                o = torch.empty((lvl.shape[0], 3, lvl.shape[-2], lvl.shape[-1]), dtype=torch.float32, device=self.dev)
                o[:, 0:1] = lvl
                for i in range(lvl.shape[0]):
                    st = _lib.image_gradients(_lib.ptr(o[i, 0]), lvl.shape[-2], lvl.shape[-1], _lib.ptr(o[i, 1]),
                                              _lib.ptr(o[i, 2]), _lib.stream_ptr(self.dev))
                    _lib.check(st, "como_b200_image_gradients")
                res.append(o)
        return res

    # ------------------------------------------------------------------ keyframe / one-way decisions
    def check_keyframe(self, median_depth, num_reproj_depth, T_curr_kf):
        new_kf = False
        num_kf_pixels = self.vals_pyr[-1].shape[1]
        if self.last_kf_sent_ts <= self.kf_received_ts:
            kf_dist = torch.linalg.norm(T_curr_kf[:, :3, 3])
            if kf_dist > self.cfg["keyframing"]["kf_depth_motion_ratio"] * median_depth:
                new_kf = True
            elif self.cfg["keyframing"]["kf_num_pixels_frac"] > num_reproj_depth / num_kf_pixels:
                new_kf = True
        return new_kf

    def check_one_way_frame(self, median_depth, num_reproj_depth, T_curr_kf, T_w_curr):
        new_one_way_frame = False
        extra_count = 1 if self.last_kf_sent_ts > self.kf_received_ts else 0
        thresh_scale_kf = (1.0 + self.num_one_way_since_kf + extra_count) / (1.0 + self.cfg["keyframing"]["one_way_freq"])
        dist_thresh = self.cfg["keyframing"]["kf_depth_motion_ratio"] * median_depth
        num_kf_pixels = self.vals_pyr[-1].shape[1]
        pixel_thresh = (1 - self.cfg["keyframing"]["kf_num_pixels_frac"]) * num_kf_pixels
        num_empty_pixels = num_kf_pixels - num_reproj_depth
        kf_dist = torch.linalg.norm(T_curr_kf[:, :3, 3])
        if kf_dist > thresh_scale_kf * dist_thresh:
            new_one_way_frame = True
        elif num_empty_pixels > thresh_scale_kf * pixel_thresh:
            new_one_way_frame = True
        if new_one_way_frame:
            self.last_one_way_empty_pixels = num_empty_pixels
            self.T_w_rec_last = T_w_curr
        return new_one_way_frame

    def get_reproj_last_kf(self, T_curr_kf):
        """(1,H,W) depth image of the last keyframe's finest cloud seen from the current frame (NaN = empty)."""
        P_last = self.P_pyr[-1][-1].contiguous()
        n = P_last.shape[0]
        H, W = self.level_sizes[-1]
        T = T_curr_kf.reshape(4, 4).to(torch.float32).contiguous()
        with torch.cuda.device(self.dev):
            st = _lib.reproj_depth(_lib.ptr(P_last), n, _lib.ptr(T), self._K9[-1], H, W, _lib.ptr(self._winner),
                                   _lib.ptr(self._reproj), _lib.stream_ptr(self.dev))
        _lib.check(st, "como_b200_reproj_depth")
        return self._reproj.view(1, H, W)

    def _reproj_stats(self, T_curr_kf):
        self.get_reproj_last_kf(T_curr_kf)
        H, W = self.level_sizes[-1]
        with torch.cuda.device(self.dev):
            st = _lib.median_f32(_lib.ptr(self._reproj), _lib.ptr(self._seg), 1, H * W, 1.0, _lib.ptr(self._med),
                                 _lib.ptr(self._cnt), _lib.ptr(self._med_ws), self._med_ws.numel(), _lib.stream_ptr(self.dev))
        _lib.check(st, "como_b200_median_f32")
        return self._med[0], self._cnt[0]

    # ------------------------------------------------------------------ keyframe reference
    def update_kf_reference(self, kf_data):
        timestamps, kf_rgb, kf_pose, kf_aff, depth = kf_data
        kf_pose = kf_pose.to(device=self.dev, dtype=self.dtype)
        kf_aff = kf_aff.to(device=self.dev, dtype=self.dtype)
        depth = depth.to(device=self.dev, dtype=torch.float32).contiguous()
        if timestamps[-1] > self.kf_received_ts and self.mapping_init:
            num_kf = kf_pose.shape[0]
            self.T_w_f = self.get_curr_world_pose()
            self.T_curr_kf = _inv_se3(self.T_w_f) @ kf_pose[num_kf - 1:num_kf]
            self.aff_w_f = self.get_curr_world_aff()
            last_aff = kf_aff[num_kf - 1:num_kf]
            rel = torch.empty_like(self.aff_w_f)
            rel[:, 0, :] = self.aff_w_f[:, 0, :] - last_aff[:, 0, :]
            rel[:, 1, :] = torch.exp(-rel[:, 0, :]) * (self.aff_w_f[:, 1, :] - last_aff[:, 1, :])
            self.aff_curr_kf = rel
            self.reset_one_way_vars()
        elif not self.mapping_init:
            self.mapping_init = True
            self.last_kf_sent_ts = timestamps[-1]

        if timestamps[-1] != self.kf_received_ts:
            self._kf_img_grads = self._pyramid(kf_rgb, True)   # gray + pyramid + Scharr of every level: one launch
            self._kf_img_pyr = [g[:, 0:1] for g in self._kf_img_grads]
        num_kf = kf_pose.shape[0]
        rel_poses = (_inv_se3(kf_pose[num_kf - 1:num_kf]) @ kf_pose).to(torch.float32).contiguous()  # (B,4,4)
        Hf, Wf = int(depth.shape[-2]), int(depth.shape[-1])
        if (Hf, Wf) != tuple(self.level_sizes[-1]) or depth.shape[0] != kf_pose.shape[0]:
            raise RuntimeError(f"update_kf_reference: depth {tuple(depth.shape)} does not match {kf_pose.shape[0]} keyframes "
                               f"of size {tuple(self.level_sizes[-1])}")
        self.vals_pyr, self.img_grads_pyr, self.coords_pyr = [], [], []
        self.P_pyr, self.dI_dT_pyr, self.mask_pyr = [], [], []
        for l, (h, w) in enumerate(self.level_sizes):
            n = h * w
            sub = 2 ** (self.num_levels - 1 - l)
            vals = torch.empty((num_kf, n, 1), dtype=torch.float32, device=self.dev)
            grads = torch.empty((num_kf, n, 1, 2), dtype=torch.float32, device=self.dev)
            P = torch.empty((num_kf, n, 3), dtype=torch.float32, device=self.dev)
            J = torch.empty((num_kf, n, 1, 8), dtype=torch.float32, device=self.dev)
            mask = torch.empty((num_kf, n), dtype=torch.uint8, device=self.dev)
            ig = self._kf_img_grads[l]
            for b in range(num_kf):
                with torch.cuda.device(self.dev):
                    st = _lib.kf_reference_level(
                        _lib.ptr(ig[b, 0]), _lib.ptr(ig[b, 1]), _lib.ptr(ig[b, 2]), _lib.ptr(depth[b, 0]), Hf, Wf, sub, h, w,
                        self._K9[l], _lib.ptr(rel_poses[b]), 50.0, 1e-4, _lib.ptr(vals[b]), _lib.ptr(grads[b]), _lib.ptr(P[b]),
                        _lib.ptr(J[b]), _lib.ptr(mask[b]), _lib.stream_ptr(self.dev))
                _lib.check(st, "como_b200_kf_reference_level")
            self.vals_pyr.append(vals)
            self.img_grads_pyr.append(grads)
            self.P_pyr.append(P)
            self.dI_dT_pyr.append(J)
            self.mask_pyr.append(mask.view(torch.bool))
            rr, cc = torch.meshgrid(torch.arange(h, device=self.dev), torch.arange(w, device=self.dev), indexing="ij")
            self.coords_pyr.append(torch.stack((rr.reshape(-1), cc.reshape(-1)), 1)[None].repeat(num_kf, 1, 1))
        self._plan = None   # new reference operands: the launch plan is rebuilt by the next handle_frame
        self.kf_received_ts = timestamps[-1]
        self.T_w_kf = kf_pose[num_kf - 1:num_kf]
        self.aff_w_kf = kf_aff[num_kf - 1:num_kf]

    # ------------------------------------------------------------------ per-frame tracking
    def handle_frame(self, data):
        timestamp, rgb = data
        if rgb.shape[0] == 1 and self.vals_pyr[0].shape[0] == 1:
            # one fused front-end launch into the persistent pyramid, one tracker launch from the cached plan
            self._pyramid(rgb, False, out=self._img_bufs)
            if self._plan is None:
                self._plan = TrackBatchPlan([(self.vals_pyr, self.P_pyr, self.dI_dT_pyr, self.mask_pyr, self.intrinsics_pyr,
                                              self._img_bufs)], self.cfg["term_criteria"])
            T, aff, nit = self._plan.run(self.T_curr_kf, self.aff_curr_kf)
            self.T_curr_kf, self.aff_curr_kf = T.clone(), aff.clone()
            _pt.last_num_iters = nit
        else:
            img_pyr = self.prep_tracking_img(rgb)
            self.T_curr_kf, self.aff_curr_kf = photo_tracking_pyr(
                self.T_curr_kf, self.aff_curr_kf, self.vals_pyr, self.P_pyr, self.dI_dT_pyr, self.mask_pyr,
                self.intrinsics_pyr, img_pyr, self.cfg["sigmas"]["photo"], self.cfg["term_criteria"])
        T_w_curr = self.get_curr_world_pose()
        track_data_viz = (timestamp, T_w_curr.clone())
        track_data_map = None
        median_depth, num_valid = self._reproj_stats(self.T_curr_kf)
        # the decisions need three numbers on the host: one pinned read-back (one synchronisation) instead of one
        # implicit device->host read per comparison
        self._host.copy_(torch.cat((self.T_curr_kf.reshape(16), median_depth.reshape(1), num_valid.reshape(1).float())),
                         non_blocking=True)
        torch.cuda.current_stream(self.dev).synchronize()
        T_host = self._host[:16].reshape(1, 4, 4).clone()
        med_h, cnt_h = float(self._host[16]), float(self._host[17])
        new_kf = self.check_keyframe(med_h, cnt_h, T_host)
        if new_kf:
            track_data_map = ("keyframe", rgb.clone(), self.T_curr_kf, self.aff_curr_kf, self.kf_received_ts, timestamp)
            self.last_kf_sent_ts = timestamp
        else:
            if self.check_one_way_frame(med_h, cnt_h, T_host, T_w_curr):
                track_data_map = ("one-way", rgb.clone(), self.T_curr_kf, self.aff_curr_kf, self.kf_received_ts, timestamp)
                self.last_rec_sent_ts = timestamp
                self.num_one_way_since_kf += 1
        return track_data_viz, track_data_map
