"""diffusers-signature front end of the prompt-pair ControlNet: same call contract as the reference's
``PromptDiffusionControlNetModel.forward`` (promptdiffusioncontrolnet.py:188-391; called from
pipeline_prompt_diffusion.py:1237-1246), executed by the same CUDA kernels as ``cldm.ControlNet``.

The reference class inherits all arithmetic from ``diffusers.ControlNetModel`` (not vendored, not installed
here — parity of this front end is therefore pinned only THROUGH the ldm path: a diffusers-layout state dict
is mapped key-by-key onto the ldm layout, whose kernels are checked against the reference's golden vectors).
The math is identical: ``ControlNetConditioningEmbedding`` is the 16-16-32-32-96-96-256 conv
stack of cldm.py:147-181, ``flip_sin_to_cos=True, freq_shift=0`` is the [cos, sin] order of util.py:169,
ResNet GroupNorm eps 1e-5, transformer GroupNorm eps 1e-6, 8 heads.
"""
from __future__ import annotations

import re
from dataclasses import dataclass
from typing import Any, Dict, List, Mapping, Optional, Tuple, Union

import torch

from . import ops
from .cldm.cldm import ControlNet
from .config import CLDM_V15, CLDMConfig


@dataclass
class ControlNetOutput:
    """Field-compatible stand-in for ``diffusers.models.controlnet.ControlNetOutput``."""
    down_block_res_samples: Tuple[torch.Tensor, ...]
    mid_block_res_sample: torch.Tensor

    def __iter__(self):
        return iter((self.down_block_res_samples, self.mid_block_res_sample))


_RESNET = {"norm1": "in_layers.0", "conv1": "in_layers.2", "time_emb_proj": "emb_layers.1",
           "norm2": "out_layers.0", "conv2": "out_layers.3", "conv_shortcut": "skip_connection"}
_HINT = {"conv_in": 0, "blocks.0": 2, "blocks.1": 4, "blocks.2": 6, "blocks.3": 8, "blocks.4": 10, "blocks.5": 12,
         "conv_out": 14}


def diffusers_key_to_ldm(key: str, layers_per_block: int = 2) -> Optional[str]:
    """Map one diffusers ``ControlNetModel``-layout key to the ldm ``ControlNet`` layout (no prefix).
    Returns None for keys that have no counterpart on this path."""
    lpb = layers_per_block
    m = re.match(r"time_embedding\.linear_(\d)\.(.+)", key)
    if m:
        return f"time_embed.{0 if m.group(1) == '1' else 2}.{m.group(2)}"
    m = re.match(r"conv_in\.(.+)", key)
    if m:
        return f"input_blocks.0.0.{m.group(1)}"
    m = re.match(r"down_blocks\.(\d+)\.resnets\.(\d+)\.([a-z0-9_]+)\.(.+)", key)
    if m:
        i, j = int(m.group(1)), int(m.group(2))
        return f"input_blocks.{(lpb + 1) * i + j + 1}.0.{_RESNET[m.group(3)]}.{m.group(4)}"
    m = re.match(r"down_blocks\.(\d+)\.attentions\.(\d+)\.(.+)", key)
    if m:
        i, j = int(m.group(1)), int(m.group(2))
        return f"input_blocks.{(lpb + 1) * i + j + 1}.1.{m.group(3)}"
    m = re.match(r"down_blocks\.(\d+)\.downsamplers\.0\.conv\.(.+)", key)
    if m:
        return f"input_blocks.{(lpb + 1) * (int(m.group(1)) + 1)}.0.op.{m.group(2)}"
    m = re.match(r"mid_block\.resnets\.(\d)\.([a-z0-9_]+)\.(.+)", key)
    if m:
        return f"middle_block.{0 if m.group(1) == '0' else 2}.{_RESNET[m.group(2)]}.{m.group(3)}"
    m = re.match(r"mid_block\.attentions\.0\.(.+)", key)
    if m:
        return f"middle_block.1.{m.group(1)}"
    m = re.match(r"controlnet_down_blocks\.(\d+)\.(.+)", key)
    if m:
        return f"zero_convs.{m.group(1)}.0.{m.group(2)}"
    m = re.match(r"controlnet_mid_block\.(.+)", key)
    if m:
        return f"middle_block_out.0.{m.group(1)}"
    m = re.match(r"controlnet_(query_)?cond_embedding\.(conv_in|conv_out|blocks\.\d)\.(.+)", key)
    if m:
        stem = "input_cond_block" if m.group(1) else "input_hint_block"
        return f"{stem}.{_HINT[m.group(2)]}.{m.group(3)}"
    return None


def ldm_key_to_diffusers(key: str, layers_per_block: int = 2) -> Optional[str]:
    """Inverse of :func:`diffusers_key_to_ldm` (used to build diffusers-layout fixtures)."""
    lpb = layers_per_block
    inv_res = {v: k for k, v in _RESNET.items()}
    inv_hint = {v: k for k, v in _HINT.items()}
    m = re.match(r"time_embed\.(\d)\.(.+)", key)
    if m:
        return f"time_embedding.linear_{1 if m.group(1) == '0' else 2}.{m.group(2)}"
    m = re.match(r"input_blocks\.0\.0\.(.+)", key)
    if m:
        return f"conv_in.{m.group(1)}"
    m = re.match(r"input_blocks\.(\d+)\.0\.op\.(.+)", key)
    if m:
        return f"down_blocks.{int(m.group(1)) // (lpb + 1) - 1}.downsamplers.0.conv.{m.group(2)}"
    m = re.match(r"(input_blocks\.(\d+)|middle_block)\.(\d)\.(in_layers\.\d|out_layers\.\d|emb_layers\.1|skip_connection)\.(.+)", key)
    if m:
        name = inv_res[m.group(4)]
        if m.group(1) == "middle_block":
            return f"mid_block.resnets.{0 if m.group(3) == '0' else 1}.{name}.{m.group(5)}"
        n = int(m.group(2)) - 1
        return f"down_blocks.{n // (lpb + 1)}.resnets.{n % (lpb + 1)}.{name}.{m.group(5)}"
    m = re.match(r"input_blocks\.(\d+)\.1\.(.+)", key)
    if m:
        n = int(m.group(1)) - 1
        return f"down_blocks.{n // (lpb + 1)}.attentions.{n % (lpb + 1)}.{m.group(2)}"
    m = re.match(r"middle_block\.1\.(.+)", key)
    if m:
        return f"mid_block.attentions.0.{m.group(1)}"
    m = re.match(r"zero_convs\.(\d+)\.0\.(.+)", key)
    if m:
        return f"controlnet_down_blocks.{m.group(1)}.{m.group(2)}"
    m = re.match(r"middle_block_out\.0\.(.+)", key)
    if m:
        return f"controlnet_mid_block.{m.group(1)}"
    m = re.match(r"input_(hint|cond)_block\.(\d+)\.(.+)", key)
    if m:
        pre = "controlnet_cond_embedding" if m.group(1) == "hint" else "controlnet_query_cond_embedding"
        return f"{pre}.{inv_hint[int(m.group(2))]}.{m.group(3)}"
    return None


class PromptDiffusionControlNetModel:
    def __init__(self, cfg: CLDMConfig = CLDM_V15, mode: str = "bf16", device="cuda",
                 controlnet_conditioning_channel_order: str = "rgb", global_pool_conditions: bool = False,
                 class_embed_type: Optional[str] = None, num_class_embeds: Optional[int] = None,
                 addition_embed_type: Optional[str] = None, net: Optional[ControlNet] = None):
        """``net``: wrap an existing ``cldm.ControlNet`` (e.g. ``ControlLDM.control_model``) instead of building one —
        the pipeline then runs the fused step over the shared buffer pool."""
        if class_embed_type is not None or num_class_embeds is not None or addition_embed_type is not None:
            raise NotImplementedError("class / additional embeddings (promptdiffusioncontrolnet.py:288-320) are not "
                                      "part of the SD1.5 prompt-diffusion configuration")
        self.net = net if net is not None else ControlNet(cfg, mode, device)
        self.config = type("Config", (), {})()
        self.config.class_embed_type, self.config.num_class_embeds = class_embed_type, num_class_embeds
        self.config.addition_embed_type = addition_embed_type
        self.class_embedding = None
        self.config.controlnet_conditioning_channel_order = controlnet_conditioning_channel_order
        self.config.global_pool_conditions = global_pool_conditions
        self.config.in_channels = cfg.in_channels
        self.config.cross_attention_dim = cfg.context_dim
        self.device = self.net.device
        self.dtype = torch.float32

    def load_state_dict(self, sd: Mapping[str, torch.Tensor], strict: bool = True):
        """Accepts a diffusers-layout state dict (``conv_in.weight``, ``down_blocks.0.resnets.0...``,
        ``controlnet_cond_embedding...``) or an ldm-layout one (``control_model.*`` / bare ldm keys)."""
        if any(k.startswith("control_model.") for k in sd):
            self.net.load_state_dict(sd)
            return self
        if any(k.startswith("input_blocks.") for k in sd):
            self.net.load_state_dict(sd, prefix="")
            return self
        mapped, unknown = {}, []
        for k, v in sd.items():
            lk = diffusers_key_to_ldm(k)
            if lk is None:
                unknown.append(k)
            else:
                mapped[lk] = v
        if strict and unknown:
            raise KeyError(f"unexpected diffusers keys (first 5): {unknown[:5]}")
        self.net.load_state_dict(mapped, prefix="")
        return self

    @torch.no_grad()
    @ops.on_device
    def forward(self, sample: torch.Tensor, timestep: Union[torch.Tensor, float, int],
                encoder_hidden_states: torch.Tensor, controlnet_cond: torch.Tensor,
                controlnet_query_cond: torch.Tensor, conditioning_scale: float = 1.0,
                class_labels: Optional[torch.Tensor] = None, timestep_cond: Optional[torch.Tensor] = None,
                attention_mask: Optional[torch.Tensor] = None,
                added_cond_kwargs: Optional[Dict[str, torch.Tensor]] = None,
                cross_attention_kwargs: Optional[Dict[str, Any]] = None, guess_mode: bool = False,
                return_dict: bool = True):
        order = self.config.controlnet_conditioning_channel_order
        if order == "bgr":
            controlnet_cond = torch.flip(controlnet_cond, dims=[1])
        elif order != "rgb":
            raise ValueError(f"unknown `controlnet_conditioning_channel_order`: {order}")
        # `class_labels` and `added_cond_kwargs` are read only behind `self.class_embedding is not None` /
        # `config.addition_embed_type is not None` (:288-320): the SD1.5 prompt-diffusion config has neither, so the
        # reference accepts and ignores them — mirrored.  (A config that asks for those embeddings is refused by the
        # constructor.)  The rest feeds arithmetic this path does not have:
        #   timestep_cond          -> TimestepEmbedding.cond_proj (None without `time_cond_proj_dim`: the reference fails);
        #   attention_mask         -> a [B, 1, keys] bias on every SELF-attention (:257-259), which cannot broadcast over
        #                             the 4096 / 1024 / 256 / 64-token levels of this UNet and is never passed by the
        #                             pipeline (pipeline_prompt_diffusion.py:1237-1246);
        #   cross_attention_kwargs -> attention processors (LoRA scale), not part of this model.
        for name, val in (("timestep_cond", timestep_cond), ("attention_mask", attention_mask),
                          ("cross_attention_kwargs", cross_attention_kwargs)):
            if val is not None:
                raise NotImplementedError(f"{name} is not part of the SD1.5 prompt-diffusion path")

        # 1. time (promptdiffusioncontrolnet.py:262-276)
        timesteps = timestep
        if not torch.is_tensor(timesteps):
            timesteps = torch.tensor([timesteps], dtype=torch.float64 if isinstance(timestep, float) else torch.int64,
                                     device=sample.device)
        elif timesteps.dim() == 0:
            timesteps = timesteps[None].to(sample.device)
        timesteps = timesteps.expand(sample.shape[0])
        if timesteps.is_floating_point():
            if not bool((timesteps == timesteps.round()).all()):
                raise NotImplementedError("fractional timesteps are not supported by the int64 embedding kernel")
        timesteps = timesteps.to(torch.int64)

        outs = self.net.forward(sample, timesteps, controlnet_cond, controlnet_query_cond, encoder_hidden_states)
        down, mid = outs[:-1], outs[-1]

        # 6. scaling (:371-384)
        if guess_mode and not self.config.global_pool_conditions:
            scales = torch.logspace(-1, 0, len(down) + 1, device=mid.device) * conditioning_scale
            down = [s * sc for s, sc in zip(down, scales)]
            mid = mid * scales[-1]
        else:
            down = [s * conditioning_scale for s in down]
            mid = mid * conditioning_scale
        if self.config.global_pool_conditions:
            down = [torch.mean(s, dim=(2, 3), keepdim=True) for s in down]
            mid = torch.mean(mid, dim=(2, 3), keepdim=True)
        if not return_dict:
            return (down, mid)
        return ControlNetOutput(down_block_res_samples=down, mid_block_res_sample=mid)

    __call__ = forward
