"""diff3dhpe_b200: B200-native (sm_100a) DDIM / MixSTE-s2s sampler behind the reference's module API.

    from diff3dhpe_b200 import HPE_model, GaussianDiffusion     # drop-in for common.nets.load_net / the s2s
                                                                 # conditional_diffusion module of the reference
"""
from .diffusion import GaussianDiffusion
from .load_net import HPE_model
from .model import ConditionalDiffusionMixSTES2SGRANDLinLift

__all__ = ["GaussianDiffusion", "HPE_model", "ConditionalDiffusionMixSTES2SGRANDLinLift"]
