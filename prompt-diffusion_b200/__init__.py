"""B200-native Prompt-Diffusion denoising hot path (ControlLDM.apply_model inside
the DDIM loop) behind the reference's own Python signatures.

Import as ``prompt_diffusion_b200`` (alias of this directory).  Heavy pieces
(the CUDA C-ABI library) are loaded on first use by ``prompt_diffusion_b200._lib``
and fail loudly if the in-tree ``libpd_b200.so`` is missing.
"""
__version__ = "0.1.0"

from .config import CLDM_V15, CLDMConfig  # noqa: E402,F401


def __getattr__(name):
    # heavy, CUDA-library-backed symbols are imported lazily
    if name in ("ControlLDM", "ControlNet", "ControlledUnetModel", "DDIMSampler"):
        from . import cldm as _cldm
        return getattr(_cldm, name)
    if name in ("PromptDiffusionPipeline", "DDIMScheduler", "UNet2DConditionShim"):
        from . import pipeline_prompt_diffusion as _ppd
        return getattr(_ppd, name)
    if name == "FrozenCLIPTextEncoder":
        from .clip_text import FrozenCLIPTextEncoder
        return FrozenCLIPTextEncoder
    if name == "AutoencoderKLDecoder":
        from .autoencoder import AutoencoderKLDecoder
        return AutoencoderKLDecoder
    if name == "PromptDiffusionControlNetModel":
        from .promptdiffusioncontrolnet import PromptDiffusionControlNetModel
        return PromptDiffusionControlNetModel
    raise AttributeError(name)
