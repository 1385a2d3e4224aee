"""ORACLE — test infrastructure, not product code.

Plain-PyTorch fp32 restatement of the reference's first-stage DECODE
(``LatentDiffusion.decode_first_stage`` -> ``AutoencoderKL.decode`` -> ``Decoder.forward``), written
functionally over a reference-format state dict (``first_stage_model.*``).  Only ``tests/`` and ``bench.py``'s
CPU-baseline leg may import it.

Parity status: PINNED — ``tests/golden/make_golden_vae.py`` instantiates the reference's own ``Decoder``
(ldm/modules/diffusionmodules/model.py) in the dev container with the same procedural checkpoint and stores its
output in ``tests/golden/vae_decoder_golden.npz``; ``tests/test_oracle_golden.py`` holds this file to it.
"""
from __future__ import annotations

from typing import Mapping

import torch
import torch.nn.functional as F

VAE = "first_stage_model."


def _gn(sd, key, x):
    """Normalize(): GroupNorm(32, eps=1e-6, affine) — model.py:42-43."""
    return F.group_norm(x, 32, sd[key + ".weight"], sd[key + ".bias"], eps=1e-6)


def _conv(sd, key, x, padding):
    return F.conv2d(x, sd[key + ".weight"], sd[key + ".bias"], padding=padding)


def resnet_block(sd, key, x):
    """ResnetBlock.forward with temb=None, dropout 0 — model.py:123-145 (nin_shortcut when cin != cout :112-121)."""
    h = _conv(sd, key + ".conv1", F.silu(_gn(sd, key + ".norm1", x)), 1)
    h = _conv(sd, key + ".conv2", F.silu(_gn(sd, key + ".norm2", h)), 1)
    if (key + ".nin_shortcut.weight") in sd:
        x = _conv(sd, key + ".nin_shortcut", x, 0)
    return x + h


def attn_block(sd, key, x):
    """AttnBlock.forward — model.py:176-203: single head over h*w tokens, scale c^-0.5."""
    h_ = _gn(sd, key + ".norm", x)
    q, k, v = (_conv(sd, f"{key}.{n}", h_, 0) for n in ("q", "k", "v"))
    b, c, h, w = q.shape
    q = q.reshape(b, c, h * w).permute(0, 2, 1)
    k = k.reshape(b, c, h * w)
    w_ = torch.bmm(q, k) * (int(c) ** (-0.5))
    w_ = F.softmax(w_, dim=2)
    v = v.reshape(b, c, h * w)
    h_ = torch.bmm(v, w_.permute(0, 2, 1)).reshape(b, c, h, w)
    return x + _conv(sd, key + ".proj_out", h_, 0)


def decoder_forward(sd: Mapping[str, torch.Tensor], z: torch.Tensor, prefix: str = VAE + "decoder",
                    ch_mult=(1, 2, 4, 4), num_res_blocks: int = 2) -> torch.Tensor:
    """Decoder.forward — model.py:618-653 (no attention in the up path for cldm_v15: attn_resolutions [])."""
    sd = {k: v.to(z.device, torch.float32) for k, v in sd.items() if k.startswith(prefix)}
    h = _conv(sd, prefix + ".conv_in", z, 1)
    h = resnet_block(sd, prefix + ".mid.block_1", h)
    h = attn_block(sd, prefix + ".mid.attn_1", h)
    h = resnet_block(sd, prefix + ".mid.block_2", h)
    for i_level in reversed(range(len(ch_mult))):
        for i_block in range(num_res_blocks + 1):
            h = resnet_block(sd, f"{prefix}.up.{i_level}.block.{i_block}", h)
        if i_level != 0:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")            # Upsample.forward model.py:60-64
            h = _conv(sd, f"{prefix}.up.{i_level}.upsample.conv", h, 1)
    h = F.silu(_gn(sd, prefix + ".norm_out", h))
    return _conv(sd, prefix + ".conv_out", h, 1)


def decode(sd: Mapping[str, torch.Tensor], z: torch.Tensor) -> torch.Tensor:
    """AutoencoderKL.decode — ldm/models/autoencoder.py:88-91."""
    w, b = sd[VAE + "post_quant_conv.weight"], sd[VAE + "post_quant_conv.bias"]
    z = F.conv2d(z.float(), w.to(z.device, torch.float32), b.to(z.device, torch.float32))
    return decoder_forward(sd, z)


def decode_first_stage(sd: Mapping[str, torch.Tensor], z: torch.Tensor, scale_factor: float = 0.18215) -> torch.Tensor:
    """LatentDiffusion.decode_first_stage — ldm/models/diffusion/ddpm.py:820-828 (predict_cids False)."""
    return decode(sd, 1.0 / scale_factor * z)
