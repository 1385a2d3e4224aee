"""Model configuration + block topology of the ControlLDM denoising path.

The numbers mirror ``models/cldm_v15.yaml`` of the reference (control_stage_config
/ unet_config, lines 30-62) and the way ``UNetModel.__init__``
(ldm/modules/diffusionmodules/openaimodel.py:542-730) and ``ControlNet.__init__``
(cldm/cldm.py:138-297) expand them into blocks.  Only the structure is derived
here; no reference module is instantiated.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, NamedTuple, Tuple


@dataclass(frozen=True)
class CLDMConfig:
    in_channels: int = 4
    out_channels: int = 4
    model_channels: int = 320
    hint_channels: int = 6          # example pair: condition-of-example || example image
    query_channels: int = 3         # Prompt-Diffusion's second hint encoder (cldm.py:165-181)
    channel_mult: Tuple[int, ...] = (1, 2, 4, 4)
    num_res_blocks: int = 2
    attention_resolutions: Tuple[int, ...] = (4, 2, 1)
    num_heads: int = 8
    context_dim: int = 768
    timesteps: int = 1000
    linear_start: float = 0.00085
    linear_end: float = 0.0120
    parameterization: str = "eps"
    scale_factor: float = 0.18215   # latent scaling of decode_first_stage (cldm_v15.yaml:17, ddpm.py:827)

    @property
    def time_embed_dim(self) -> int:
        return self.model_channels * 4

    @classmethod
    def from_yaml(cls, path: str) -> "CLDMConfig":
        """Parse a reference-style yaml (models/cldm_v15.yaml) with PyYAML."""
        import yaml
        with open(path) as f:
            y = yaml.safe_load(f)
        p = y["model"]["params"]
        c = p["control_stage_config"]["params"]
        u = p["unet_config"]["params"]
        for k in ("model_channels", "channel_mult", "num_res_blocks", "attention_resolutions",
                  "num_heads", "context_dim"):
            if c[k] != u[k]:
                raise ValueError(f"control/unet config mismatch on {k}")
        if not u.get("use_spatial_transformer", False) or u.get("legacy", True):
            raise ValueError("only use_spatial_transformer=True, legacy=False is supported")
        return cls(in_channels=u["in_channels"], out_channels=u["out_channels"],
                   model_channels=u["model_channels"], hint_channels=c["hint_channels"],
                   channel_mult=tuple(u["channel_mult"]), num_res_blocks=u["num_res_blocks"],
                   attention_resolutions=tuple(u["attention_resolutions"]),
                   num_heads=u["num_heads"], context_dim=u["context_dim"],
                   timesteps=p.get("timesteps", 1000), linear_start=p["linear_start"],
                   linear_end=p["linear_end"], scale_factor=float(p.get("scale_factor", 1.0)))


CLDM_V15 = CLDMConfig()


class Conv(NamedTuple):
    key: str
    cin: int
    cout: int
    ksize: int
    stride: int


class Res(NamedTuple):
    key: str
    cin: int
    cout: int


class ST(NamedTuple):        # SpatialTransformer with one BasicTransformerBlock
    key: str
    ch: int
    heads: int
    d_head: int


class Down(NamedTuple):      # conv3x3 stride 2, weights at key + ".op"
    key: str
    ch: int


class Up(NamedTuple):        # nearest x2 then conv3x3, weights at key + ".conv"
    key: str
    ch: int


HINT_STACK = ((16, 1), (16, 1), (32, 2), (32, 1), (96, 2), (96, 1), (256, 2))  # (cout, stride)


@dataclass
class Topology:
    input_blocks: List[list] = field(default_factory=list)
    input_chans: List[int] = field(default_factory=list)    # channels of each hs entry
    middle: list = field(default_factory=list)
    output_blocks: List[list] = field(default_factory=list)  # UNet only
    mid_ch: int = 0


def build_topology(cfg: CLDMConfig, with_decoder: bool = True) -> Topology:
    """Encoder/middle(/decoder) block list; same loop as openaimodel.py:553-724."""
    topo = Topology()
    mc = cfg.model_channels
    topo.input_blocks.append([Conv("input_blocks.0.0", cfg.in_channels, mc, 3, 1)])
    topo.input_chans.append(mc)
    ch, ds, idx = mc, 1, 1
    for level, mult in enumerate(cfg.channel_mult):
        for _ in range(cfg.num_res_blocks):
            layers = [Res(f"input_blocks.{idx}.0", ch, mult * mc)]
            ch = mult * mc
            if ds in cfg.attention_resolutions:
                layers.append(ST(f"input_blocks.{idx}.1", ch, cfg.num_heads, ch // cfg.num_heads))
            topo.input_blocks.append(layers)
            topo.input_chans.append(ch)
            idx += 1
        if level != len(cfg.channel_mult) - 1:
            topo.input_blocks.append([Down(f"input_blocks.{idx}.0", ch)])
            topo.input_chans.append(ch)
            idx += 1
            ds *= 2
    topo.mid_ch = ch
    topo.middle = [Res("middle_block.0", ch, ch),
                   ST("middle_block.1", ch, cfg.num_heads, ch // cfg.num_heads),
                   Res("middle_block.2", ch, ch)]
    if with_decoder:
        chans = list(topo.input_chans)
        oidx = 0
        for level, mult in list(enumerate(cfg.channel_mult))[::-1]:
            for i in range(cfg.num_res_blocks + 1):
                ich = chans.pop()
                layers = [Res(f"output_blocks.{oidx}.0", ch + ich, mc * mult)]
                ch = mc * mult
                if ds in cfg.attention_resolutions:
                    layers.append(ST(f"output_blocks.{oidx}.{len(layers)}", ch, cfg.num_heads,
                                     ch // cfg.num_heads))
                if level and i == cfg.num_res_blocks:
                    layers.append(Up(f"output_blocks.{oidx}.{len(layers)}", ch))
                    ds //= 2
                topo.output_blocks.append(layers)
                oidx += 1
    return topo
