"""Checkpoint ingest for the B200 path: mirrors of the reference's ``cldm/model.py`` helpers
(``get_state_dict`` :8-9, ``load_state_dict`` :12-21, ``create_model`` :24-28) and of the key scheme of
``tool_add_control.py`` (:17-48), so the notebook's two set-up lines keep working unchanged::

    model = create_model('./models/cldm_v15.yaml')
    model.load_state_dict(load_state_dict('./models/prompt_diffusion.ckpt', location='cuda'))

``create_model`` returns the B200 ``ControlLDM`` (CUDA only — there is no CPU model to ``.cpu()`` into; the
method exists and returns ``self``).  ``load_state_dict`` reads ``.ckpt`` / ``.pth`` (``torch.load``) and
``.safetensors`` files and unwraps a ``{'state_dict': ...}`` envelope exactly like the reference.  The loaded
dict may carry the full Lightning checkpoint (``first_stage_model.*``, ``cond_stage_model.*``, schedule buffers,
``model_ema.*``): ``ControlLDM.load_state_dict`` picks ``control_model.*`` and ``model.diffusion_model.*`` and
ignores the rest (VAE / CLIP are out of scope of this path).
"""
from __future__ import annotations

import os
from typing import Dict, Iterable, List, Mapping, Tuple

import torch

from ..config import CLDMConfig


def get_state_dict(d):
    """cldm/model.py:8-9."""
    return d.get("state_dict", d)


def load_state_dict(ckpt_path: str, location: str = "cpu", allow_pickle: bool = False) -> Dict[str, torch.Tensor]:
    """cldm/model.py:12-21 (same extension dispatch, same envelope handling, same log line).

    ``.safetensors`` is the preferred format.  ``.ckpt`` / ``.pth`` files are read with ``weights_only=True`` (tensors
    and plain containers only); a checkpoint that needs full unpickling — which can execute arbitrary code from the
    file — is loaded only on explicit opt-in (``allow_pickle=True`` or ``PD_B200_ALLOW_PICKLE=1``), with a warning."""
    _, extension = os.path.splitext(ckpt_path)
    if extension.lower() == ".safetensors":
        import safetensors.torch
        state_dict = safetensors.torch.load_file(ckpt_path, device=location)
    else:
        try:
            raw = torch.load(ckpt_path, map_location=torch.device(location), weights_only=True)
        except Exception as e:
            if not (allow_pickle or os.environ.get("PD_B200_ALLOW_PICKLE") == "1"):
                raise RuntimeError(
                    f"{ckpt_path} cannot be read with weights_only=True ({type(e).__name__}: {e}). Full unpickling can "
                    "run arbitrary code from the file: convert the checkpoint to .safetensors, or opt in with "
                    "allow_pickle=True / PD_B200_ALLOW_PICKLE=1 if you trust its source.") from e
            import warnings
            warnings.warn(f"loading {ckpt_path} with full unpickling (weights_only=False): only do this for trusted files")
            raw = torch.load(ckpt_path, map_location=torch.device(location), weights_only=False)
        state_dict = get_state_dict(raw)
    state_dict = get_state_dict(state_dict)
    print(f"Loaded state_dict from [{ckpt_path}]")
    return state_dict


def create_model(config_path: str, mode: str = "bf16", device: str = "cuda"):
    """cldm/model.py:24-28: the yaml names the reference classes; this path instantiates its own ``ControlLDM``
    with the topology read from the same file (``CLDMConfig.from_yaml`` rejects configs it does not implement)."""
    from .cldm import ControlLDM
    cfg = CLDMConfig.from_yaml(config_path)
    model = ControlLDM(cfg, mode=mode, device=device)
    print(f"Loaded model config from [{config_path}]")
    return model


def get_node_name(name: str, parent_name: str) -> Tuple[bool, str]:
    """tool_add_control.py:17-23."""
    if len(name) <= len(parent_name):
        return False, ""
    p = name[:len(parent_name)]
    if p != parent_name:
        return False, ""
    return True, name[len(parent_name):]


def add_control(pretrained_weights: Mapping[str, torch.Tensor], scratch_dict: Mapping[str, torch.Tensor],
                verbose: bool = True) -> Tuple[Dict[str, torch.Tensor], List[str]]:
    """Key scheme of ``tool_add_control.py:28-48``: build a ControlLDM state dict from a Stable-Diffusion one.

    Every key ``k`` of ``scratch_dict`` (the freshly initialised ControlLDM) is filled from the pretrained
    checkpoint: ``control_<rest>`` copies ``model.diffusion_<rest>`` (so ``control_model.input_blocks.*`` starts as
    a copy of the UNet encoder), any other key copies itself; keys the checkpoint does not have (hint encoders,
    zero convs, ``middle_block_out``) keep their scratch values.  Returns ``(target_dict, newly_added_keys)``."""
    pretrained_weights = get_state_dict(pretrained_weights)
    target: Dict[str, torch.Tensor] = {}
    added: List[str] = []
    for k in scratch_dict.keys():
        is_control, name = get_node_name(k, "control_")
        copy_k = "model.diffusion_" + name if is_control else k
        if copy_k in pretrained_weights:
            target[k] = pretrained_weights[copy_k].clone()
        else:
            target[k] = scratch_dict[k].clone()
            added.append(k)
            if verbose:
                print(f"These weights are newly added: {k}")
    return target, added


def path_keys(sd: Mapping[str, torch.Tensor]) -> Iterable[str]:
    """The checkpoint entries this path consumes (everything else in a full Lightning checkpoint is ignored)."""
    return (k for k in sd if k.startswith("control_model.") or k.startswith("model.diffusion_model."))
