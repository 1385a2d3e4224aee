"""One-time checkpoint ingest: reference ``state_dict`` (OIHW / [out,in] fp32, key grammar
of SURVEY.md appendix B, written by tool_add_control.py:36-48 and read by
cldm/model.py:12-21) -> K-major tap-major device weights in the compute dtype, with the
load-time fusions the kernels rely on:

* ResBlock ``out_layers.3`` 3x3 conv and ``skip_connection`` 1x1 conv share one weight
  matrix ``[Cout, 9*Cout + Cin]`` (second K segment) and one summed bias;
* attn1 ``to_q|to_k|to_v`` -> one ``[3C, C]`` matrix; attn2 ``to_k|to_v`` -> ``[2C, ctx]``;
* all ``emb_layers.1`` of a net -> one ``[sum Cout, 4*mc]`` matrix (one GEMM per step);
* bf16 mode folds each LayerNorm of BasicTransformerBlock into the linear layer that consumes it (``ops.fold_layer_norm``:
  gamma-scaled weights, beta pushed through into the bias, column sums for the mean correction): the normalised
  tensor is never written, the GEMM reads the raw residual stream and a per-row (mean, rstd) pair;
* bf16 mode pads input channels that are >= 16 but not a multiple of 64 (hint stacks:
  16, 32, 96) up to 64/128 so those convs also run on the tcgen05 engine.
"""
from __future__ import annotations

from typing import Dict, List, Mapping, Optional

import torch

from . import ops
from .config import CLDMConfig, Conv, Down, HINT_STACK, Res, ST, Up, build_topology


# k-block-major weight layout (ops.block_weight): supported by the tcgen05 engine, measured on B200 at BASELINE
# config 2 and found ~3 % SLOWER than plain [Cout][K] rows (18.2 vs 17.6 ms of GEMM time per denoising step: the
# strided 128-byte pieces spread over more HBM channels than one contiguous 12-32 KiB run), so it stays off.
BLOCK_WEIGHTS = False

# Split-precision entry and exit of the UNet / ControlNet (bf16 mode).  The oracle's operand-rounding hook shows that
# rounding the operands of just TWO tiny convs to bf16 — conv_in (4 -> 320, both nets) and the UNet's `out` conv
# (320 -> 4) — accounts for 2.7e-3 and 2.4e-3 of the 7.0e-3 eps error that bf16 GEMM operands cost in total (their
# inputs / outputs ARE the latent and eps, nothing averages the rounding out).  Both run here as hi/lo bf16 pairs
# (x = hi + lo, w = hi + lo; hi*hi + lo*hi + hi*lo) inside padding the tcgen05 engine needs anyway: conv_in's K is
# zero-padded from 4 to 64 channels per tap, `out`'s N from 4 to 8 rows; only `out` pays (its K doubles: +~60 us / step).
SPLIT_IN_CONV = True
SPLIT_OUT_CONV = True

# Upsample (nearest x2 + conv3x3) as four 2x2 phase convolutions of the low-resolution tensor (bf16 mode)
PHASE_UPSAMPLE = True

# Fold every LayerNorm of the transformer blocks into the linear layer behind it (bf16 mode).  Off = separate
# pd_layer_norm passes (the fp32 mode always uses those).
FOLD_LAYER_NORM = True


def pad_channels(c: int, dt: torch.dtype) -> int:
    if dt == torch.bfloat16 and c >= 16 and c % 64 != 0:
        return (c + 63) // 64 * 64
    return c


class PConv:
    __slots__ = ("w", "bias", "cin", "cin_pad", "cout", "cout_pad", "ksize", "stride", "c2", "role", "colsum",
                 "split", "phases")

    def __init__(self, w, bias, cin, cin_pad, cout, ksize, stride, c2=0, role="conv"):
        self.w, self.bias, self.cin, self.cin_pad, self.cout = w, bias, cin, cin_pad, cout
        self.cout_pad = cout
        self.colsum = None          # set when a LayerNorm is folded into this layer (ops.fold_layer_norm)
        self.ksize, self.stride, self.c2, self.role = ksize, stride, c2, role
        self.split = None           # "in": activation columns [hi | lo | hi], "out": [hi | lo] and rows [W_hi | W_lo]
        self.phases = None          # Upsample: four 2x2 phase convs [(py, px, PConv)] replacing the 3x3 on the 4x tensor


class PNorm:
    __slots__ = ("gamma", "beta")

    def __init__(self, gamma, beta):
        self.gamma, self.beta = gamma, beta


class PRes:
    __slots__ = ("key", "cin", "cout", "gn1", "conv1", "gn2", "conv2", "emb_off", "has_skip")


class PST:
    __slots__ = ("key", "ch", "heads", "d", "gn", "proj_in", "ln1", "wqkv", "out1", "ln2", "wq2", "wkv2",
                 "out2", "ln3", "ff1", "ff1_geglu", "ff2", "proj_out", "ln_folded")


class Packer:
    def __init__(self, sd: Mapping[str, torch.Tensor], prefix: str, dt: torch.dtype, device):
        self.sd, self.prefix, self.dt, self.device = sd, prefix, dt, device

    def t(self, key: str) -> torch.Tensor:
        try:
            v = self.sd[self.prefix + key]
        except KeyError:
            raise KeyError(f"checkpoint is missing '{self.prefix + key}'") from None
        return v.to(device=self.device, dtype=torch.float32, non_blocking=True)

    def vec(self, key: str) -> torch.Tensor:
        return self.t(key).contiguous()

    def norm(self, key: str) -> PNorm:
        return PNorm(self.vec(key + ".weight"), self.vec(key + ".bias"))

    def split_in_conv(self, key: str) -> PConv:
        """conv_in (4 -> 320) with the latent in split precision: per tap the 64-channel K block holds
        [w_hi (cin) | w_hi (cin) | w_lo (cin) | 0 ...] against activations [x_hi | x_lo | x_hi | 0 ...]."""
        w = self.t(key + ".weight")
        cout, cin, kh, kw = w.shape
        w_hi = w.to(torch.bfloat16).float()
        w3 = torch.cat([w_hi, w_hi, w - w_hi], dim=1).contiguous()
        out = torch.empty((cout, kh * kw * 64), dtype=self.dt, device=self.device)
        ops.repack_conv_weight(w3, out, cin_pad=64)
        pc = PConv(self.block(out), self.vec(key + ".bias"), cin, 64, cout, kh, 1)
        pc.split = "in"
        return pc

    def split_out_conv(self, key: str) -> PConv:
        """`out` conv (320 -> 4) over split-precision activations [a_hi (C) | a_lo (C)]: rows [0, cout) hold w_hi for
        both halves, rows [cout, 2 cout) hold w_lo for the hi half; the two row groups are summed after the GEMM."""
        w = self.t(key + ".weight")
        cout, cin, kh, kw = w.shape
        assert 2 * cout <= 8
        w_hi = w.to(torch.bfloat16).float()
        w2 = torch.zeros((8, 2 * cin, kh, kw), dtype=torch.float32, device=self.device)
        w2[:cout, :cin] = w_hi
        w2[:cout, cin:] = w_hi
        w2[cout:2 * cout, :cin] = w - w_hi
        out = torch.empty((8, kh * kw * 2 * cin), dtype=self.dt, device=self.device)
        ops.repack_conv_weight(w2, out, cin_pad=2 * cin)
        bias = torch.zeros(8, dtype=torch.float32, device=self.device)
        bias[:cout] = self.vec(key + ".bias")
        pc = PConv(self.block(out), bias.contiguous(), 2 * cin, 2 * cin, cout, kh, 1)
        pc.cout_pad = 8
        pc.split = "out"
        return pc

    def up_phases(self, key: str):
        """Upsample.forward = nearest x2 then conv3x3 (openaimodel.py:108-118).  Output pixel (2y + py, 2x + px) only
        ever sees source rows {y - 1, y} (py = 0) or {y, y + 1} (py = 1), likewise in x: four 2x2 convolutions of the
        LOW-resolution tensor whose taps are the sums of the 3x3 taps that land on the same source pixel
        (rows: py = 0 -> [k0, k1 + k2], py = 1 -> [k0 + k1, k2]).  Summed in fp32, then rounded once."""
        w = self.t(key + ".weight")
        cout, cin, kh, kw = w.shape
        assert kh == 3 and kw == 3
        sets = {(0, 0): (0,), (0, 1): (1, 2), (1, 0): (0, 1), (1, 1): (2,)}
        bias = self.vec(key + ".bias")
        phases = []
        for py in (0, 1):
            for px in (0, 1):
                wp = torch.zeros((cout, cin, 2, 2), dtype=torch.float32, device=self.device)
                for dy in (0, 1):
                    for dx in (0, 1):
                        for ky in sets[(py, dy)]:
                            for kx in sets[(px, dx)]:
                                wp[:, :, dy, dx] += w[:, :, ky, kx]
                out = torch.empty((cout, 4 * cin), dtype=self.dt, device=self.device)
                ops.repack_conv_weight(wp, out)
                phases.append((py, px, PConv(self.block(out), bias, cin, cin, cout, 2, 1)))
        return phases

    def conv(self, key: str, stride: int = 1, skip_key: Optional[str] = None, pad: bool = True,
             tc_small: bool = False) -> PConv:
        """``tc_small`` (bf16 mode only): also pad a tiny channel count so that the layer runs on the tcgen05
        engine — input channels up to 64 (conv_in: 4 -> 64, zero columns in the activation buffer) and output
        channels up to 8 (out conv: 4 -> 8, zero weight rows; the caller slices the result)."""
        w = self.t(key + ".weight")
        if w.dim() == 2:
            w = w[:, :, None, None]
        cout, cin, kh, kw = w.shape
        cin_pad = pad_channels(cin, self.dt) if pad else cin
        cout_pad = cout
        if tc_small and self.dt == torch.bfloat16:
            cin_pad = (cin + 63) // 64 * 64
            cout_pad = (cout + 7) // 8 * 8
        ktot = kh * kw * cin_pad
        c2 = 0
        if skip_key is not None:
            ws = self.t(skip_key + ".weight")
            c2 = ws.shape[1]
            ktot += c2
        out = (torch.zeros if cout_pad != cout else torch.empty)((cout_pad, ktot), dtype=self.dt, device=self.device)
        ops.repack_conv_weight(w, out[:cout], cin_pad=cin_pad, k_offset=0)
        bias = self.vec(key + ".bias") if (self.prefix + key + ".bias") in self.sd else None
        if skip_key is not None:
            ops.repack_conv_weight(ws, out, k_offset=kh * kw * cin_pad)
            bias = (bias + self.vec(skip_key + ".bias")).contiguous()
        if cout_pad != cout and bias is not None:
            bias = torch.cat([bias, torch.zeros(cout_pad - cout, dtype=bias.dtype, device=bias.device)]).contiguous()
        pc = PConv(self.block(out), bias, cin, cin_pad, cout, kh, stride, c2)
        pc.cout_pad = cout_pad
        return pc

    def block(self, w: torch.Tensor) -> torch.Tensor:
        """Optional k-block-major weight layout for the layers the tcgen05 engine runs (see BLOCK_WEIGHTS)."""
        if BLOCK_WEIGHTS and self.dt == torch.bfloat16 and w.shape[1] % 64 == 0 and w.shape[0] % 8 == 0:
            return ops.block_weight(w)
        return w

    def unblocked(self, key: str) -> torch.Tensor:
        """[Cout, K] K-major weight of a linear layer in the compute dtype (before any blocking)."""
        w = self.t(key + ".weight")
        out = torch.empty((w.shape[0], w.shape[1]), dtype=self.dt, device=self.device)
        ops.repack_conv_weight(w, out)
        return out

    def stacked_linear(self, keys: List[str], with_bias: bool) -> PConv:
        ws = [self.t(k + ".weight") for k in keys]
        cin = ws[0].shape[1]
        cout = sum(w.shape[0] for w in ws)
        out = torch.empty((cout, cin), dtype=self.dt, device=self.device)
        r = 0
        for w in ws:
            ops.repack_conv_weight(w, out[r:r + w.shape[0]])
            r += w.shape[0]
        bias = torch.cat([self.vec(k + ".bias") for k in keys]).contiguous() if with_bias else None
        return PConv(self.block(out), bias, cin, cin, cout, 1, 1)

    def res(self, layer: Res) -> PRes:
        k = layer.key
        r = PRes()
        r.key, r.cin, r.cout = k, layer.cin, layer.cout
        r.gn1 = self.norm(k + ".in_layers.0")
        r.conv1 = self.conv(k + ".in_layers.2")
        r.gn2 = self.norm(k + ".out_layers.0")
        r.has_skip = layer.cin != layer.cout
        r.conv2 = self.conv(k + ".out_layers.3", skip_key=(k + ".skip_connection") if r.has_skip else None)
        r.emb_off = -1
        return r

    def st(self, layer: ST) -> PST:
        k = layer.key
        tb = k + ".transformer_blocks.0"
        s = PST()
        s.key, s.ch, s.heads, s.d = k, layer.ch, layer.heads, layer.d_head
        s.gn = self.norm(k + ".norm")
        s.proj_in = self.conv(k + ".proj_in")
        s.ln1, s.ln2, s.ln3 = self.norm(tb + ".norm1"), self.norm(tb + ".norm2"), self.norm(tb + ".norm3")
        s.wqkv = self.stacked_linear([tb + ".attn1.to_q", tb + ".attn1.to_k", tb + ".attn1.to_v"], False)
        s.out1 = self.conv(tb + ".attn1.to_out.0")
        s.wq2 = self.conv(tb + ".attn2.to_q")
        s.wkv2 = self.stacked_linear([tb + ".attn2.to_k", tb + ".attn2.to_v"], False)
        s.out2 = self.conv(tb + ".attn2.to_out.0")
        s.ff1 = self.conv(tb + ".ff.net.0.proj")
        # bf16 mode: GEGLU runs in ff1's GEMM epilogue -> rows interleaved (32 values | 32 gates), see ops.geglu_interleave
        s.ff1_geglu = None
        if self.dt == torch.bfloat16 and (4 * layer.ch) % 64 == 0:
            s.ff1_geglu = PConv(self.block(ops.geglu_interleave(self.unblocked(tb + ".ff.net.0.proj"))),
                                ops.geglu_interleave(s.ff1.bias), s.ff1.cin, s.ff1.cin_pad, s.ff1.cout, 1, 1)
            s.ff1 = None            # the plain layout is not needed (saves 0.4 GB of weights)
        s.ff2 = self.conv(tb + ".ff.net.2")
        s.proj_out = self.conv(k + ".proj_out")
        # bf16 mode: LayerNorm folded into its consumer (same condition as the tcgen05 engine: C % 64 == 0)
        s.ln_folded = False
        if FOLD_LAYER_NORM and s.ff1_geglu is not None and layer.ch % 64 == 0:
            wq, wk, wv = (self.t(tb + f".attn1.to_{n}.weight") for n in "qkv")
            s.wqkv = self.folded(torch.cat([wq, wk, wv], 0), None, s.ln1)
            s.wq2 = self.folded(self.t(tb + ".attn2.to_q.weight"), None, s.ln2)
            s.ff1_geglu = self.folded(self.t(tb + ".ff.net.0.proj.weight"), self.vec(tb + ".ff.net.0.proj.bias"), s.ln3,
                                      interleave=True)
            s.ln_folded = True
        return s

    def folded(self, w: torch.Tensor, bias: Optional[torch.Tensor], ln: PNorm, interleave: bool = False) -> PConv:
        """Linear layer with the LayerNorm ``ln`` that feeds it folded in (bf16 weights, fp32 bias' and colsum)."""
        ws, b, cs = ops.fold_layer_norm(w, bias, ln.gamma, ln.beta, self.dt)
        if interleave:
            ws, b, cs = ops.geglu_interleave(ws), ops.geglu_interleave(b), ops.geglu_interleave(cs)
        pc = PConv(self.block(ws.contiguous()), b.contiguous(), w.shape[1], w.shape[1], w.shape[0], 1, 1)
        pc.colsum = cs.contiguous()
        return pc

    def layer(self, layer):
        if isinstance(layer, Res):
            return self.res(layer)
        if isinstance(layer, ST):
            return self.st(layer)
        if isinstance(layer, Conv):                                               # input_blocks.0.0 (4 -> 320)
            if SPLIT_IN_CONV and self.dt == torch.bfloat16 and 3 * self.t(layer.key + ".weight").shape[1] <= 64:
                return self.split_in_conv(layer.key)
            return self.conv(layer.key, stride=layer.stride, tc_small=True)
        if isinstance(layer, Down):
            pc = self.conv(layer.key + ".op", stride=2)
            pc.role = "down"
            return pc
        if isinstance(layer, Up):
            pc = self.conv(layer.key + ".conv")
            pc.role = "up"
            if PHASE_UPSAMPLE and self.dt == torch.bfloat16 and pc.cin % 64 == 0 and pc.cout % 8 == 0:
                pc.phases = self.up_phases(layer.key + ".conv")
                pc.w = None                                   # the 3x3 weight on the 4x tensor is never used
            return pc
        raise TypeError(layer)


class PackedNet:
    """Packed weights of one net (UNet or ControlNet)."""

    def __init__(self, cfg: CLDMConfig, sd: Mapping[str, torch.Tensor], prefix: str, decoder: bool,
                 dt: torch.dtype, device):
        pk = Packer(sd, prefix, dt, device)
        topo = build_topology(cfg, with_decoder=decoder)
        self.topo = topo
        self.te0 = pk.conv("time_embed.0")
        self.te2 = pk.conv("time_embed.2")
        self.input_blocks = [[pk.layer(l) for l in blk] for blk in topo.input_blocks]
        self.middle = [pk.layer(l) for l in topo.middle]
        self.output_blocks = [[pk.layer(l) for l in blk] for blk in topo.output_blocks] if decoder else []
        # batch every ResBlock's emb_layers.1 into one GEMM
        res_layers = [l for blk in (self.input_blocks + [self.middle] + self.output_blocks) for l in blk
                      if isinstance(l, PRes)]
        off = 0
        for r in res_layers:
            r.emb_off = off
            off += r.cout
        self.emb_all = pk.stacked_linear([r.key + ".emb_layers.1" for r in res_layers], True)
        self.emb_total = off
        if decoder:
            self.out_norm = pk.norm("out.0")
            if SPLIT_OUT_CONV and dt == torch.bfloat16 and cfg.model_channels % 32 == 0 and 2 * cfg.out_channels <= 8:
                self.out_conv = pk.split_out_conv("out.2")                          # 320 -> 4, hi / lo operands
                self.out_conv_plain = pk.conv("out.2", tc_small=True)               # latents without 64-pixel records
            else:
                self.out_conv = pk.conv("out.2", tc_small=True)                     # 320 -> 4
        else:
            self.zero_convs = [pk.conv(f"zero_convs.{i}.0") for i in range(len(topo.input_blocks))]
            self.middle_out = pk.conv("middle_block_out.0")
            self.hint_pair = self._hint(pk, "input_hint_block", cfg)
            self.hint_query = self._hint(pk, "input_cond_block", cfg)

    @staticmethod
    def _hint(pk: Packer, stem: str, cfg: CLDMConfig) -> List[PConv]:
        convs = []
        for i, (_cout, stride) in enumerate(HINT_STACK):
            convs.append(pk.conv(f"{stem}.{2 * i}", stride=stride))
        convs.append(pk.conv(f"{stem}.{2 * len(HINT_STACK)}"))
        return convs
