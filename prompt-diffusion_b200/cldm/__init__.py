"""Mirrors of the reference's ``cldm`` package for the denoising path."""
from .cldm import ControlLDM, ControlNet, ControlledUnetModel  # noqa: F401
from .ddim_hacked import DDIMSampler  # noqa: F401
