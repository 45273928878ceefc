"""Drop-in installer: routes the reference package's hot path to the B200 kernels without touching
como_dataset.py, the GUI, the orchestration or the data loaders.

    import como_b200.patch as p; p.install()       # before `from como.odom...` objects are created
    python como/como_dataset.py ...                 # unchanged entry point

What is replaced (reference name -> como_b200 implementation):
  sys.modules["como_backends"]                                  -> como_b200.como_backends
  como.depth_cov.core.samplers.sample_sparse_coords             -> como_b200.depth_cov.core.samplers.sample_sparse_coords
  como.odom.frontend.photo_tracking.photo_tracking_pyr / precalc_jacobians (also as imported by como.odom.Tracking)
  como.odom.Mapping.Mapping.iterate / store_vars / prep_predictor / get_img_and_grads
  como.odom.backend.linear_system.solve_system
  como.odom.frontend.corr.track_and_init (also as imported by como.odom.Mapping) and
  como.depth_cov.core.distill_depth.distill_depth_from_scratch / distill_conditional_depth_from_scratch
  como.odom.frontend.two_frame_sfm.two_frame_sfm_pyr / two_frame_sfm / setup_reference (also as imported by
  como.odom.frontend.TwoFrameSfm)
  como.utils.multiprocessing.TupleTensorQueue / transfer_data / release_data / init_gpu (also as imported by
  como.odom.multiprocessing.ComoMp when that module is importable): slot ring shared once over CUDA IPC

Aliasing contract (the reference's update_vars rebinds fresh tensors; mapping_core.iterate updates kf_poses,
kf_aff_params, recent_*, P_m IN PLACE): holders of slices of those tensors see the update.  The reference's own
consumers copy on the way out (get_kf_viz_data clones; get_kf_ref_data goes through transfer_data to another dtype /
the queue, and como_b200's TupleTensorQueue.pop returns tensors that own their memory), so nothing in the reference's
flow observes the difference; code that keeps raw slices across iterate() calls must clone them.
Everything else (UNet, orchestration) keeps running the reference's own Python on the
same device.  Requires tracking.dtype float / mapping.dtype double, color gray (the tracking operators themselves also
take rgb; the Tracking class front-end and the mapper do not).
"""
import sys


def install():
    import como_b200.como_backends as cb

    sys.modules["como_backends"] = cb  # must happen before como.depth_cov.core.samplers is imported
    import como.depth_cov.core.samplers as ref_samplers
    import como.odom.backend.linear_system as ref_ls
    import como.odom.frontend.photo_tracking as ref_pt
    import como.odom.Mapping as ref_mapping
    import como.odom.Tracking as ref_tracking

    from como_b200.depth_cov.core import predictor as b_pred
    from como_b200.depth_cov.core import samplers as b_samplers
    from como_b200.odom import mapping_core as mc
    from como_b200.odom.frontend import photo_tracking as b_pt

    ref_samplers.como_backends = cb
    ref_samplers.sample_sparse_coords = b_samplers.sample_sparse_coords
    for mod in list(sys.modules.values()):
        if mod is not None and getattr(mod, "__name__", "").startswith("como.") and hasattr(mod, "sample_sparse_coords"):
            mod.sample_sparse_coords = b_samplers.sample_sparse_coords
    ref_pt.photo_tracking_pyr = b_pt.photo_tracking_pyr
    ref_pt.precalc_jacobians = b_pt.precalc_jacobians
    ref_tracking.photo_tracking_pyr = b_pt.photo_tracking_pyr
    ref_tracking.precalc_jacobians = b_pt.precalc_jacobians
    ref_ls.solve_system = mc.solve_system

    def iterate(self):
        return mc.iterate(self, self.cfg)

    def store_vars(self, pm, logzm, Knm_Kmminv):
        return mc.store_vars(self, pm, logzm, Knm_Kmminv)

    def prep_predictor(self, cov_params_img, coords_m):
        scale = float(self.model.cov_modules[-1].get_scale())
        return b_pred.prep_predictor(cov_params_img, coords_m, scale, photo_img_size=self.kf_img_and_grads.shape[-2:])

    import como.depth_cov.core.distill_depth as ref_dd
    import como.odom.frontend.corr as ref_corr

    from como_b200.depth_cov.core import distill_depth as b_dd
    from como_b200.odom.frontend import corr as b_corr

    ref_corr.track_and_init = b_corr.track_and_init
    ref_mapping.track_and_init = b_corr.track_and_init
    for name in ("distill_depth_from_scratch", "distill_conditional_depth_from_scratch"):
        setattr(ref_dd, name, getattr(b_dd, name))
        setattr(ref_corr, name, getattr(b_dd, name))
    def get_img_and_grads(self, rgb):
        if self.cfg["color"] != "gray":
            raise NotImplementedError("como_b200 implements color: gray only")
        return mc.get_img_and_grads(rgb.to(self.dtype)).to(self.dtype)

    ref_mapping.Mapping.get_img_and_grads = get_img_and_grads
    import como.odom.frontend.two_frame_sfm as ref_sfm
    import como.odom.frontend.TwoFrameSfm as ref_sfm_cls

    from como_b200.odom.frontend import two_frame_sfm as b_sfm

    for name in ("two_frame_sfm_pyr", "two_frame_sfm", "setup_reference"):
        setattr(ref_sfm, name, getattr(b_sfm, name))
    ref_sfm_cls.two_frame_sfm_pyr = b_sfm.two_frame_sfm_pyr
    ref_sfm_cls.setup_reference = b_sfm.setup_reference
    import como.utils.multiprocessing as ref_mp

    from como_b200.utils import multiprocessing as b_mp

    for name in ("TupleTensorQueue", "transfer_data", "release_data", "init_gpu"):
        setattr(ref_mp, name, getattr(b_mp, name))
    comp = sys.modules.get("como.odom.multiprocessing.ComoMp")   # needs open3d/glfw to import: patched only if loaded
    if comp is not None:
        for name in ("TupleTensorQueue", "release_data", "init_gpu"):
            if hasattr(comp, name):
                setattr(comp, name, getattr(b_mp, name))
    ref_mapping.Mapping.iterate = iterate
    ref_mapping.Mapping.store_vars = store_vars
    ref_mapping.Mapping.prep_predictor = prep_predictor
    return {"patched": ["como_backends", "sample_sparse_coords", "photo_tracking_pyr", "precalc_jacobians",
                        "Mapping.iterate", "Mapping.store_vars", "Mapping.prep_predictor", "Mapping.get_img_and_grads", "solve_system",
                        "track_and_init", "distill_depth_from_scratch", "distill_conditional_depth_from_scratch",
                        "two_frame_sfm_pyr", "setup_reference", "TupleTensorQueue"]}
