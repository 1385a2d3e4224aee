"""Golden vectors of the text-conditioning step from the library the reference calls (dev container only).

    python tests/golden/make_golden_clip.py [--out tests/golden]

``FrozenCLIPEmbedder`` (ldm/modules/encoders/modules.py:88-128) is a thin wrapper around Hugging Face
``CLIPTextModel.from_pretrained("openai/clip-vit-large-patch14")``; there is no network for the weights, so the same
architecture is built from its config, loaded (strict) with the procedural checkpoint of ``prompt_diffusion_b200.synth``
and run on tokenizer-shaped synthetic ids.  Stores ``last_hidden_state`` (what ``layer="last"`` returns, :122-123).
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=HERE)
    a = ap.parse_args()
    import transformers
    from transformers import CLIPTextConfig, CLIPTextModel
    from prompt_diffusion_b200.synth import CLIP_PREFIX, synthetic_clip_state_dict, synthetic_tokens
    cfg = CLIPTextConfig(vocab_size=49408, hidden_size=768, intermediate_size=3072, num_hidden_layers=12,
                         num_attention_heads=12, max_position_embeddings=77, hidden_act="quick_gelu",
                         layer_norm_eps=1e-5, projection_dim=768, pad_token_id=1, bos_token_id=49406, eos_token_id=49407)
    model = CLIPTextModel(cfg).eval()
    sd = synthetic_clip_state_dict(seed=0)
    missing, unexpected = model.load_state_dict({k[len(CLIP_PREFIX):]: v for k, v in sd.items()}, strict=False)
    assert not unexpected and all("position_ids" in m for m in missing), (missing, unexpected)
    tokens = synthetic_tokens(3, seed=2)
    with torch.no_grad():
        z = model(input_ids=tokens).last_hidden_state
    print("transformers", transformers.__version__, tuple(z.shape), float(z.abs().mean()))
    np.savez_compressed(os.path.join(a.out, "clip_text_golden.npz"), tokens=tokens.numpy(), z=z.numpy(),
                        n_params=np.int64(sum(v.numel() for v in sd.values())),
                        transformers_version=np.array(transformers.__version__))
    print("wrote", os.path.join(a.out, "clip_text_golden.npz"))


if __name__ == "__main__":
    main()
