"""ORACLE — test infrastructure, not product code.

Plain-PyTorch fp32 restatement of the text-conditioning step the reference runs right before the denoising loop:
``FrozenCLIPEmbedder.forward`` (ldm/modules/encoders/modules.py:117-128, ``layer="last"``) from the token ids on, i.e.
``CLIPTextModel(input_ids).last_hidden_state`` of the ViT-L/14 text tower.

The arithmetic lives in a THIRD-PARTY dependency that is not vendored in the reference: Hugging Face ``transformers``
(pinned ``transformers==4.19.2`` in the reference's environment.yaml:23; 5.5.0 is what this image has).  Restated from
its published algorithm (``CLIPTextTransformer``: token + position embeddings; 12 pre-LayerNorm blocks of causal
multi-head self-attention (12 heads, d = 64, scale d^-1/2 applied to q) and a 768-3072-768 MLP with quick-GELU
``x * sigmoid(1.702 x)``; final LayerNorm; eps 1e-5).  Parity status: PINNED to the library itself —
``tests/golden/make_golden_clip.py`` runs ``transformers.CLIPTextModel`` (random-init config of
openai/clip-vit-large-patch14, procedural weights) and stores ``last_hidden_state`` in
``tests/golden/clip_text_golden.npz``.  The tokenizer (needs the vocabulary files) is out of scope: inputs are token ids.
"""
from __future__ import annotations

from typing import Mapping

import torch
import torch.nn.functional as F

CLIP = "cond_stage_model.transformer."


def clip_text_forward(sd: Mapping[str, torch.Tensor], tokens: torch.Tensor, heads: int = 12, prefix: str = CLIP) -> torch.Tensor:
    """tokens int64 [B, L] -> last_hidden_state fp32 [B, L, C]."""
    t = prefix + "text_model."
    dev = tokens.device
    w = lambda k: sd[t + k].to(dev, torch.float32)
    B, L = tokens.shape
    x = w("embeddings.token_embedding.weight")[tokens] + w("embeddings.position_embedding.weight")[:L][None]
    C = x.shape[-1]
    d = C // heads
    mask = torch.full((L, L), float("-inf"), device=dev).triu(1)           # causal: key j visible to query i iff j <= i
    i = 0
    while (t + f"encoder.layers.{i}.layer_norm1.weight") in sd:
        p = f"encoder.layers.{i}."
        h = F.layer_norm(x, (C,), w(p + "layer_norm1.weight"), w(p + "layer_norm1.bias"), 1e-5)
        q = F.linear(h, w(p + "self_attn.q_proj.weight"), w(p + "self_attn.q_proj.bias")) * d ** -0.5
        k = F.linear(h, w(p + "self_attn.k_proj.weight"), w(p + "self_attn.k_proj.bias"))
        v = F.linear(h, w(p + "self_attn.v_proj.weight"), w(p + "self_attn.v_proj.bias"))
        sp = lambda z: z.reshape(B, L, heads, d).permute(0, 2, 1, 3)
        a = torch.softmax(sp(q) @ sp(k).transpose(-1, -2) + mask, dim=-1) @ sp(v)
        a = a.permute(0, 2, 1, 3).reshape(B, L, C)
        x = x + F.linear(a, w(p + "self_attn.out_proj.weight"), w(p + "self_attn.out_proj.bias"))
        h = F.layer_norm(x, (C,), w(p + "layer_norm2.weight"), w(p + "layer_norm2.bias"), 1e-5)
        h = F.linear(h, w(p + "mlp.fc1.weight"), w(p + "mlp.fc1.bias"))
        h = h * torch.sigmoid(1.702 * h)                                    # quick_gelu
        x = x + F.linear(h, w(p + "mlp.fc2.weight"), w(p + "mlp.fc2.bias"))
        i += 1
    return F.layer_norm(x, (C,), w("final_layer_norm.weight"), w("final_layer_norm.bias"), 1e-5)
