"""Synthetic checkpoint + synthetic inputs for the ControlLDM denoising path.

There is no network for real checkpoints, so weights are procedural: every
tensor of the reference ``state_dict`` (key grammar: SURVEY.md appendix B;
produced by tool_add_control.py:36-48 as ``control_model.*`` /
``model.diffusion_model.*``) is drawn from its own ``torch.Generator`` seeded by
(seed, crc32(key)).  The same call therefore yields the same checkpoint in
the dev container, on the GPU box, in the reference (golden generation) and in
this package — independent of module construction order.

All reference ``zero_module`` sites (openaimodel.py:228,729; attention.py:312;
cldm.py:162,180,300) get non-zero values, otherwise eps == 0 and parity would be
vacuous.
"""
from __future__ import annotations

import math
import zlib
from collections import OrderedDict
from typing import Dict, Iterator, List, Tuple

import torch

from .config import CLDMConfig, Conv, Down, HINT_STACK, Res, ST, Up, build_topology

UNET_PREFIX = "model.diffusion_model."
CTRL_PREFIX = "control_model."

# kind: "w" weight (uniform +-1/sqrt(fan_in)), "zw" reference-zero-init weight
# (normal * 0.5/sqrt(fan_in)), "b" bias, "nw" norm scale, "nb" norm shift
Spec = Tuple[str, Tuple[int, ...], str, int]


def _conv_specs(key: str, cin: int, cout: int, k: int, zero: bool = False) -> List[Spec]:
    fan = cin * k * k
    return [(f"{key}.weight", (cout, cin, k, k), "zw" if zero else "w", fan),
            (f"{key}.bias", (cout,), "b", fan)]


def _lin_specs(key: str, cin: int, cout: int, bias: bool = True) -> List[Spec]:
    out = [(f"{key}.weight", (cout, cin), "w", cin)]
    if bias:
        out.append((f"{key}.bias", (cout,), "b", cin))
    return out


def _norm_specs(key: str, ch: int) -> List[Spec]:
    return [(f"{key}.weight", (ch,), "nw", ch), (f"{key}.bias", (ch,), "nb", ch)]


def _layer_specs(layer, cfg: CLDMConfig) -> List[Spec]:
    ted = cfg.time_embed_dim
    if isinstance(layer, Conv):
        return _conv_specs(layer.key, layer.cin, layer.cout, layer.ksize)
    if isinstance(layer, Res):
        k = layer.key
        s = _norm_specs(f"{k}.in_layers.0", layer.cin)
        s += _conv_specs(f"{k}.in_layers.2", layer.cin, layer.cout, 3)
        s += _lin_specs(f"{k}.emb_layers.1", ted, layer.cout)
        s += _norm_specs(f"{k}.out_layers.0", layer.cout)
        s += _conv_specs(f"{k}.out_layers.3", layer.cout, layer.cout, 3, zero=True)
        if layer.cin != layer.cout:
            s += _conv_specs(f"{k}.skip_connection", layer.cin, layer.cout, 1)
        return s
    if isinstance(layer, ST):
        k, c = layer.key, layer.ch
        tb = f"{k}.transformer_blocks.0"
        s = _norm_specs(f"{k}.norm", c)
        s += _conv_specs(f"{k}.proj_in", c, c, 1)
        for n in ("to_q", "to_k", "to_v"):
            s += _lin_specs(f"{tb}.attn1.{n}", c, c, bias=False)
        s += _lin_specs(f"{tb}.attn1.to_out.0", c, c)
        s += _lin_specs(f"{tb}.ff.net.0.proj", c, 8 * c)
        s += _lin_specs(f"{tb}.ff.net.2", 4 * c, c)
        s += _lin_specs(f"{tb}.attn2.to_q", c, c, bias=False)
        s += _lin_specs(f"{tb}.attn2.to_k", cfg.context_dim, c, bias=False)
        s += _lin_specs(f"{tb}.attn2.to_v", cfg.context_dim, c, bias=False)
        s += _lin_specs(f"{tb}.attn2.to_out.0", c, c)
        for n in ("norm1", "norm2", "norm3"):
            s += _norm_specs(f"{tb}.{n}", c)
        s += _conv_specs(f"{k}.proj_out", c, c, 1, zero=True)
        return s
    if isinstance(layer, Down):
        return _conv_specs(f"{layer.key}.op", layer.ch, layer.ch, 3)
    if isinstance(layer, Up):
        return _conv_specs(f"{layer.key}.conv", layer.ch, layer.ch, 3)
    raise TypeError(layer)


def _hint_specs(prefix: str, cin: int, cfg: CLDMConfig) -> List[Spec]:
    s, c = [], cin
    for i, (cout, _stride) in enumerate(HINT_STACK):
        s += _conv_specs(f"{prefix}.{2 * i}", c, cout, 3)
        c = cout
    s += _conv_specs(f"{prefix}.{2 * len(HINT_STACK)}", c, cfg.model_channels, 3, zero=True)
    return s


def param_specs(cfg: CLDMConfig) -> List[Spec]:
    """Every (key, shape, kind, fan_in) of the denoising path's state dict."""
    mc, ted = cfg.model_channels, cfg.time_embed_dim
    specs: List[Spec] = []
    for prefix, decoder in ((UNET_PREFIX, True), (CTRL_PREFIX, False)):
        topo = build_topology(cfg, with_decoder=decoder)
        s = _lin_specs("time_embed.0", mc, ted) + _lin_specs("time_embed.2", ted, ted)
        for blk in topo.input_blocks:
            for layer in blk:
                s += _layer_specs(layer, cfg)
        for layer in topo.middle:
            s += _layer_specs(layer, cfg)
        if decoder:
            for blk in topo.output_blocks:
                for layer in blk:
                    s += _layer_specs(layer, cfg)
            s += _norm_specs("out.0", mc)
            s += _conv_specs("out.2", mc, cfg.out_channels, 3, zero=True)
        else:
            for i, ch in enumerate(topo.input_chans):
                s += _conv_specs(f"zero_convs.{i}.0", ch, ch, 1, zero=True)
            s += _conv_specs("middle_block_out.0", topo.mid_ch, topo.mid_ch, 1, zero=True)
            s += _hint_specs("input_hint_block", cfg.hint_channels, cfg)
            s += _hint_specs("input_cond_block", cfg.query_channels, cfg)
        specs += [(prefix + k, shp, kind, fan) for (k, shp, kind, fan) in s]
    return specs


def _gen(seed: int, key: str, device) -> torch.Generator:
    g = torch.Generator(device=device)
    g.manual_seed((seed * 1000003 + zlib.crc32(key.encode())) & 0x7FFFFFFFFFFF)
    return g


def synth_tensor(key: str, shape, kind: str, fan: int, seed: int = 0, device="cpu") -> torch.Tensor:
    g = _gen(seed, key, device)
    if kind == "w":
        bound = 1.0 / math.sqrt(fan)
        return (torch.rand(shape, generator=g, device=device) * 2 - 1) * bound
    if kind == "zw":
        return torch.randn(shape, generator=g, device=device) * (0.5 / math.sqrt(fan))
    if kind == "b":
        bound = 1.0 / math.sqrt(fan)
        return (torch.rand(shape, generator=g, device=device) * 2 - 1) * bound
    if kind == "nw":
        return 1.0 + 0.1 * torch.randn(shape, generator=g, device=device)
    if kind == "nb":
        return 0.1 * torch.randn(shape, generator=g, device=device)
    if kind == "e":                                       # embedding table (fan = 1 / std)
        return torch.randn(shape, generator=g, device=device) / float(fan)
    raise ValueError(kind)


def iter_synthetic_state_dict(cfg: CLDMConfig, seed: int = 0, device="cpu",
                              prefixes=(UNET_PREFIX, CTRL_PREFIX)) -> Iterator[Tuple[str, torch.Tensor]]:
    for key, shape, kind, fan in param_specs(cfg):
        if key.startswith(tuple(prefixes)):
            yield key, synth_tensor(key, shape, kind, fan, seed, device)


def synthetic_state_dict(cfg: CLDMConfig, seed: int = 0, device="cpu",
                         prefixes=(UNET_PREFIX, CTRL_PREFIX)) -> "OrderedDict[str, torch.Tensor]":
    """fp32 state dict with the reference's key names (CPU generators are
    bit-reproducible across machines for a fixed torch build)."""
    return OrderedDict(iter_synthetic_state_dict(cfg, seed, device, prefixes))


def synthetic_inputs(cfg: CLDMConfig, batch: int, height: int, width: int, seed: int = 2,
                     device="cpu") -> Dict[str, torch.Tensor]:
    """SURVEY.md 8(d) inputs: x_T ~ N(0,1) [B,4,H/8,W/8]; cond/uncond context
    ~ N(0,1) [B,77,ctx]; example_pair ~ U(0,1) [B,6,H,W]; query ~ U(0,1) [B,3,H,W]."""
    def g(name):
        return _gen(seed, name, device)
    h8, w8 = height // 8, width // 8
    return {
        "x_T": torch.randn((batch, cfg.in_channels, h8, w8), generator=g("x_T"), device=device),
        "c_crossattn": torch.randn((batch, 77, cfg.context_dim), generator=g("c"), device=device),
        "uc_crossattn": torch.randn((batch, 77, cfg.context_dim), generator=g("uc"), device=device),
        "example_pair": torch.rand((batch, cfg.hint_channels, height, width), generator=g("pair"),
                                   device=device),
        "query": torch.rand((batch, cfg.query_channels, height, width), generator=g("query"),
                            device=device),
    }


def make_conds(inp: Dict[str, torch.Tensor]):
    """cond / un_cond dicts in the notebook's format (run_prompt_diffusion.ipynb cell 5:27-33)."""
    cond = {"c_crossattn": [inp["c_crossattn"]], "example_pair": [inp["example_pair"]],
            "query": [inp["query"]]}
    un_cond = {"c_crossattn": [inp["uc_crossattn"]], "example_pair": [inp["example_pair"]],
               "query": [inp["query"]]}
    return cond, un_cond


# ---- first-stage decoder: AutoencoderKL.decode = post_quant_conv -> Decoder ------------------
VAE_PREFIX = "first_stage_model."


def vae_decoder_specs(ch: int = 128, ch_mult=(1, 2, 4, 4), num_res_blocks: int = 2, z_channels: int = 4,
                      embed_dim: int = 4, out_ch: int = 3) -> List[Spec]:
    """(key, shape, kind, fan_in) of ``post_quant_conv`` + ``decoder.*`` exactly as
    ldm/modules/diffusionmodules/model.py:546-616 and ldm/models/autoencoder.py:44-45 create them
    (cldm_v15.yaml first_stage_config.ddconfig: ch 128, ch_mult [1,2,4,4], num_res_blocks 2, no up-path attention)."""
    def res(key, cin, cout):
        s = _norm_specs(f"{key}.norm1", cin) + _conv_specs(f"{key}.conv1", cin, cout, 3)
        s += _norm_specs(f"{key}.norm2", cout) + _conv_specs(f"{key}.conv2", cout, cout, 3)
        if cin != cout:
            s += _conv_specs(f"{key}.nin_shortcut", cin, cout, 1)
        return s
    s = _conv_specs("post_quant_conv", embed_dim, z_channels, 1)
    block_in = ch * ch_mult[-1]
    s += _conv_specs("decoder.conv_in", z_channels, block_in, 3)
    s += res("decoder.mid.block_1", block_in, block_in)
    s += _norm_specs("decoder.mid.attn_1.norm", block_in)
    for n in ("q", "k", "v", "proj_out"):
        s += _conv_specs(f"decoder.mid.attn_1.{n}", block_in, block_in, 1)
    s += res("decoder.mid.block_2", block_in, block_in)
    for i_level in reversed(range(len(ch_mult))):
        block_out = ch * ch_mult[i_level]
        for i_block in range(num_res_blocks + 1):
            s += res(f"decoder.up.{i_level}.block.{i_block}", block_in, block_out)
            block_in = block_out
        if i_level != 0:
            s += _conv_specs(f"decoder.up.{i_level}.upsample.conv", block_in, block_in, 3)
    s += _norm_specs("decoder.norm_out", block_in)
    s += _conv_specs("decoder.conv_out", block_in, out_ch, 3)
    return [(VAE_PREFIX + k, shp, kind, fan) for (k, shp, kind, fan) in s]


def synthetic_vae_state_dict(seed: int = 0, device="cpu", **kw) -> "OrderedDict[str, torch.Tensor]":
    """Procedural ``first_stage_model.{post_quant_conv,decoder}.*`` checkpoint (same per-key generators as above)."""
    return OrderedDict((key, synth_tensor(key, shape, kind, fan, seed, device))
                       for key, shape, kind, fan in vae_decoder_specs(**kw))


# ---- CLIP text encoder: FrozenCLIPEmbedder's CLIPTextModel, ViT-L/14 text tower ---------------
CLIP_PREFIX = "cond_stage_model.transformer."


def clip_text_specs(vocab: int = 49408, width: int = 768, layers: int = 12, mlp: int = 3072, max_len: int = 77) -> List[Spec]:
    """(key, shape, kind, fan) of ``CLIPTextModel("openai/clip-vit-large-patch14")`` as the LDM checkpoint stores it under
    ``cond_stage_model.transformer.`` (ldm/modules/encoders/modules.py:96-97): 196 tensors, 123 060 480 parameters."""
    t = "text_model"
    s: List[Spec] = [(f"{t}.embeddings.token_embedding.weight", (vocab, width), "e", 50),
                     (f"{t}.embeddings.position_embedding.weight", (max_len, width), "e", 100)]
    for i in range(layers):
        L = f"{t}.encoder.layers.{i}"
        for n in ("k_proj", "v_proj", "q_proj", "out_proj"):
            s += _lin_specs(f"{L}.self_attn.{n}", width, width)
        s += _norm_specs(f"{L}.layer_norm1", width)
        s += _lin_specs(f"{L}.mlp.fc1", width, mlp) + _lin_specs(f"{L}.mlp.fc2", mlp, width)
        s += _norm_specs(f"{L}.layer_norm2", width)
    s += _norm_specs(f"{t}.final_layer_norm", width)
    return [(CLIP_PREFIX + k, shp, kind, fan) for (k, shp, kind, fan) in s]


def synthetic_clip_state_dict(seed: int = 0, device="cpu", **kw) -> "OrderedDict[str, torch.Tensor]":
    return OrderedDict((key, synth_tensor(key, shape, kind, fan, seed, device))
                       for key, shape, kind, fan in clip_text_specs(**kw))


def synthetic_tokens(batch: int, seed: int = 2, max_len: int = 77, vocab: int = 49408) -> torch.Tensor:
    """Token ids shaped like the CLIP tokenizer's output (modules.py:118-120): BOS, a random-length prompt, EOS padding."""
    g = _gen(seed, "tokens", "cpu")
    ids = torch.randint(0, vocab - 2, (batch, max_len), generator=g)
    lens = torch.randint(3, max_len - 1, (batch,), generator=g)
    ids[:, 0] = vocab - 2                                   # <|startoftext|> = 49406
    for b in range(batch):
        ids[b, int(lens[b]):] = vocab - 1                   # <|endoftext|> = 49407 (also the pad token of this tokenizer)
    return ids
