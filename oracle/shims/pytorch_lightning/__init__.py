"""TEST INFRASTRUCTURE ONLY -- stand-in for pytorch_lightning (absent here; no arithmetic).

The reference uses it only as a base class + checkpoint loader
(como/depth_cov/core/DepthCovModule.py:15, como/odom/Mapping.py:402).
"""
import torch


class LightningModule(torch.nn.Module):
    @classmethod
    def load_from_checkpoint(cls, path, **kw):
        m = cls()
        sd = torch.load(path, map_location="cpu", weights_only=False)["state_dict"]
        m.load_state_dict(sd)
        return m
