"""CPU-only: como_b200.patch.install() rebinds the reference's hot-path names (needs the reference tree,
so it only runs in the authoring container)."""
import sys

import pytest

from oracle import ref_harness


@pytest.mark.skipif(not ref_harness.available(), reason="reference tree not present")
def test_install_rebinds_reference_names():
    ref_harness.load_reference()
    saved = sys.modules.get("como_backends")
    try:
        import como_b200.patch as P

        info = P.install()
        import como.odom.Mapping as RM
        import como.odom.Tracking as RT
        import como.depth_cov.core.samplers as RS
        import como_b200.odom.frontend.photo_tracking as BPT

        assert RT.photo_tracking_pyr is BPT.photo_tracking_pyr
        assert RT.precalc_jacobians is BPT.precalc_jacobians
        assert RM.Mapping.iterate.__module__ == "como_b200.patch"
        assert RS.sample_sparse_coords.__module__ == "como_b200.depth_cov.core.samplers"
        assert sys.modules["como_backends"].__name__ == "como_b200.como_backends"
        assert "Mapping.iterate" in info["patched"]
        assert RM.track_and_init.__module__ == "como_b200.odom.frontend.corr"
        import como.odom.frontend.TwoFrameSfm as RSC
        assert RSC.two_frame_sfm_pyr.__module__ == "como_b200.odom.frontend.two_frame_sfm"
        import como.odom.frontend.corr as RC
        assert RC.distill_depth_from_scratch.__module__ == "como_b200.depth_cov.core.distill_depth"
        import como.utils.multiprocessing as RMP
        assert RMP.TupleTensorQueue.__module__ == "como_b200.utils.multiprocessing"
        # no CPU fallback: the patched operators refuse CPU tensors loudly
        import torch
        with pytest.raises(RuntimeError, match="same device"):
            sys.modules["como_backends"].cross_covariance(torch.zeros(1, 1, 2), torch.zeros(1, 1, 2, 2),
                                                          torch.zeros(1, 1, 2), torch.zeros(1, 1, 2, 2), 1.0)
    finally:
        if saved is not None:
            sys.modules["como_backends"] = saved
