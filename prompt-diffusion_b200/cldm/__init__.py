"""Mirrors of the reference's ``cldm`` package for the denoising path."""
from .cldm import ControlLDM, ControlNet, ControlledUnetModel  # noqa: F401
from .ddim_hacked import DDIMSampler  # noqa: F401
from .model import add_control, create_model, get_state_dict, load_state_dict  # noqa: F401
