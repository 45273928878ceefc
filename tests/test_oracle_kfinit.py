"""Oracle (CPU restatement of the keyframe-creation path) vs goldens from the unmodified reference."""
import os

import numpy as np
import pytest
import torch

from oracle import kfinit_oracle as KO


def _case(g, ci):
    pre = f"c{ci}_"
    ins = [torch.from_numpy(g[pre + k]) for k in ("pose1", "pose2", "coords_m1", "z_m1", "z_img1", "cov_params_img2", "K")]
    corr = {k[5:]: g[k].item() for k in g.files if k.startswith("corr_")}
    samp = {k[5:]: g[k].item() for k in g.files if k.startswith("samp_")}
    return pre, ins, corr, samp


@pytest.mark.parametrize("name,ci", [("kfinit_64x48", 0), ("kfinit_64x48", 1), ("kfinit_64x48", 2), ("kfinit_96x72", 0),
                                     ("kfinit_96x72", 1)])
def test_track_and_init_oracle(golden_dir, name, ci):
    g = np.load(os.path.join(golden_dir, name + ".npz"), allow_pickle=True)
    pre, ins, corr, samp = _case(g, ci)
    dbg = {}
    c2, z2, mask, call, zall = KO.track_and_init(*ins, float(g["gp_scale"]), corr, samp, debug=dbg)
    np.testing.assert_allclose(dbg["dd_coords_m"].numpy(), g[pre + "dd_coords_m"], rtol=1e-12)
    assert dbg["dd_n"] == int(g[pre + "dd_n"])
    np.testing.assert_allclose(dbg["dd_logz_m"].numpy(), g[pre + "dd_logz_m"], rtol=0, atol=1e-8)
    np.testing.assert_allclose(dbg["dd_res_std"], float(g[pre + "dd_res_std"]), rtol=1e-7)
    np.testing.assert_array_equal(dbg["ss0_inds"].numpy(), g[pre + "ss0_inds"])
    np.testing.assert_array_equal(dbg["ss1_inds"].numpy(), g[pre + "ss1_inds"])
    np.testing.assert_array_equal(mask.numpy(), g[pre + "corr_mask"])
    np.testing.assert_allclose(c2.numpy(), g[pre + "coords_2"], rtol=0, atol=0)
    np.testing.assert_allclose(call.numpy(), g[pre + "coords_all"], rtol=1e-12)
    np.testing.assert_allclose(dbg["dc_logz_2"].numpy(), g[pre + "dc_logz_2"], rtol=0, atol=1e-8)
    np.testing.assert_allclose(z2.numpy(), g[pre + "z2"], rtol=1e-8)
    np.testing.assert_allclose(zall.numpy(), g[pre + "z_all"], rtol=1e-8)
