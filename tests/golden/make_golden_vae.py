"""Golden vectors of the first-stage DECODE from the UNMODIFIED reference (dev container only).

    python tests/golden/make_golden_vae.py [--out tests/golden]

Builds the reference's own ``AutoencoderKL``-decode pieces — ``torch.nn.Conv2d`` post_quant_conv
(ldm/models/autoencoder.py:44-45) + ``Decoder`` (ldm/modules/diffusionmodules/model.py:546-653) with the ddconfig
of ``models/cldm_v15.yaml`` — loads the procedural checkpoint of ``prompt_diffusion_b200.synth`` (strict), runs
``decoder(post_quant_conv(z / scale_factor))`` exactly as ``LatentDiffusion.decode_first_stage`` /
``AutoencoderKL.decode`` do (ddpm.py:827-828, autoencoder.py:88-91) and stores the images.  Only numeric outputs of
the reference enter the repo.
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("PD_REFERENCE_ROOT", "/root/reference")
sys.path.insert(0, HERE)
from make_golden import install_stubs  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=HERE)
    a = ap.parse_args()
    import yaml
    install_stubs()
    sys.path.insert(0, REF)
    sys.path.insert(0, REPO)
    from ldm.modules.diffusionmodules.model import Decoder  # noqa: reference class
    from prompt_diffusion_b200.synth import VAE_PREFIX, synthetic_vae_state_dict
    with open(os.path.join(REF, "models", "cldm_v15.yaml")) as f:
        y = yaml.safe_load(f)
    p = y["model"]["params"]
    dd = p["first_stage_config"]["params"]["ddconfig"]
    embed_dim = p["first_stage_config"]["params"]["embed_dim"]
    scale_factor = float(p["scale_factor"])
    torch.manual_seed(0)
    dec = Decoder(**dd).eval()
    pq = torch.nn.Conv2d(embed_dim, dd["z_channels"], 1).eval()
    sd = synthetic_vae_state_dict(seed=0)
    dec.load_state_dict({k[len(VAE_PREFIX + "decoder."):]: v for k, v in sd.items() if k.startswith(VAE_PREFIX + "decoder.")},
                        strict=True)
    pq.load_state_dict({k[len(VAE_PREFIX + "post_quant_conv."):]: v for k, v in sd.items()
                        if k.startswith(VAE_PREFIX + "post_quant_conv.")}, strict=True)
    out = {"scale_factor": np.float32(scale_factor), "n_params": np.int64(sum(v.numel() for v in sd.values()))}
    with torch.no_grad():
        for name, (b, h, w) in {"z16": (2, 16, 16), "z8x24": (1, 8, 24)}.items():
            g = torch.Generator().manual_seed(11)
            z = torch.randn((b, dd["z_channels"], h, w), generator=g)
            img = dec(pq(1.0 / scale_factor * z))
            out[name + "_z"] = z.numpy()
            out[name + "_img"] = img.numpy()
            print(name, tuple(img.shape), float(img.abs().mean()))
    np.savez_compressed(os.path.join(a.out, "vae_decoder_golden.npz"), **out)
    print("wrote", os.path.join(a.out, "vae_decoder_golden.npz"))


if __name__ == "__main__":
    main()
