"""Import the UNMODIFIED reference from /root/reference (build container only -- the GPU box has no
/root/reference).  `timm` is absent from the image; the reference imports it only for DropPath, which is
never called in eval mode (MODEL:17,123-128), so an identity stub is installed in sys.modules."""
import sys
import types

import torch.nn as nn

REFERENCE_ROOT = "/root/reference"


def _install_timm_stub():
    if "timm" in sys.modules:
        return
    timm = types.ModuleType("timm")
    models = types.ModuleType("timm.models")
    layers = types.ModuleType("timm.models.layers")

    class DropPath(nn.Module):
        def __init__(self, drop_prob=0.0):
            super().__init__()
            self.drop_prob = drop_prob

        def forward(self, x):
            return x

    layers.DropPath = DropPath
    timm.models = models
    models.layers = layers
    sys.modules.update({"timm": timm, "timm.models": models, "timm.models.layers": layers})


def load_reference():
    """Returns (ModelClass, GaussianDiffusionClass) of the reference."""
    _install_timm_stub()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from common.nets.load_net import HPE_model
    from common.conditional_diffusion_ddim_normal_directPredict_variableLoss_both_crossFrames import GaussianDiffusion
    return HPE_model("ConditionalDiffusionMixSTES2SGRANDLinLift"), GaussianDiffusion
