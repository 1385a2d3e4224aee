"""First-stage DECODE on the B200 kernels — the step right after the denoising loop.

Mirrors, for inference only:

* ``AutoencoderKL.decode(z)``  = ``decoder(post_quant_conv(z))``            (ldm/models/autoencoder.py:88-91)
* ``Decoder.forward(z)``                                                      (ldm/modules/diffusionmodules/model.py:618-653)
* ``LatentDiffusion.decode_first_stage(z)`` = ``decode(z / scale_factor)``    (ldm/models/diffusion/ddpm.py:820-828),
  exposed as ``ControlLDM.decode_first_stage`` once ``first_stage_model.*`` weights are loaded.

Every layer runs on the same C-ABI ops as the UNet: ``pd_group_norm`` (+SiLU, 32 groups, eps 1e-6), ``pd_conv2d``
(tcgen05 implicit GEMM; ``nin_shortcut`` fused as the second K segment of ``conv2``, residual adds in the epilogue),
``pd_upsample2x``.  The single 512-channel, single-head attention of the middle block (``AttnBlock.forward``
:176-203) has d = 512, beyond the streaming-softmax kernels, and is compute-dense enough to run as three GEMMs per
image around ``pd_softmax_rows``: S = q k^T, P = softmax(S c^-1/2), O = P v — with v produced already transposed
(``v^T = W_v g^T``, a GEMM whose "activation" is the weight matrix) and v's bias folded through ``proj_out``
(softmax rows sum to one).  Activations are pixel-major [B*H*W, C] in the compute dtype; there is no PyTorch compute
fallback.
"""
from __future__ import annotations

from typing import Dict, Mapping, Optional

import torch

from . import ops
from ._lib import PD_ACT_NONE, PD_ACT_SILU
from .packing import PConv, PNorm, Packer

_MODES = {"bf16": torch.bfloat16, "fp32": torch.float32}
VAE_PREFIX = "first_stage_model."


class _VRes:
    __slots__ = ("cin", "cout", "gn1", "conv1", "gn2", "conv2", "has_skip")


class AutoencoderKLDecoder:
    """``first_stage_model`` restricted to what sampling needs: ``decode``."""

    def __init__(self, mode: str = "bf16", device="cuda", ch: int = 128, ch_mult=(1, 2, 4, 4),
                 num_res_blocks: int = 2, z_channels: int = 4, embed_dim: int = 4, out_ch: int = 3):
        if mode not in _MODES:
            raise ValueError(f"mode must be one of {list(_MODES)}")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("prompt_diffusion_b200.AutoencoderKLDecoder runs on CUDA only (no CPU fallback)")
        self.mode, self.dt = mode, _MODES[mode]
        self.ch, self.ch_mult, self.num_res_blocks = ch, tuple(ch_mult), num_res_blocks
        self.z_channels, self.embed_dim, self.out_ch = z_channels, embed_dim, out_ch
        self.bufs: Dict[tuple, torch.Tensor] = {}
        self.loaded = False

    # ---- buffers -----------------------------------------------------------------------------------------------
    def buf(self, name, rows, cols, dtype=None, zero=False):
        key = (name, rows, cols, dtype or self.dt)
        b = self.bufs.get(key)
        if b is None:
            b = (torch.zeros if zero else torch.empty)((rows, cols), dtype=dtype or self.dt, device=self.device)
            self.bufs[key] = b
        return b

    def release_buffers(self):
        """Drop the activation buffers (a 512x512 batch of 8 holds ~5 GB of them)."""
        self.bufs.clear()

    # ---- weights -----------------------------------------------------------------------------------------------
    def _res(self, pk: Packer, key: str, cin: int, cout: int) -> _VRes:
        r = _VRes()
        r.cin, r.cout, r.has_skip = cin, cout, cin != cout
        r.gn1, r.gn2 = pk.norm(key + ".norm1"), pk.norm(key + ".norm2")
        r.conv1 = pk.conv(key + ".conv1")
        r.conv2 = pk.conv(key + ".conv2", skip_key=(key + ".nin_shortcut") if r.has_skip else None)
        return r

    def load_state_dict(self, sd: Mapping[str, torch.Tensor], prefix: str = VAE_PREFIX, scale_factor: float = 1.0):
        """``sd``: reference keys ``first_stage_model.{post_quant_conv,decoder}.*`` (encoder / loss entries ignored).
        ``scale_factor`` is the latent scaling of ``decode_first_stage`` (ddpm.py:827); ``decode(z, scaled=True)``
        applies ``z / scale_factor`` by using a copy of ``post_quant_conv``'s weight divided by it."""
        with torch.cuda.device(self.device):
            pk = Packer(sd, prefix, self.dt, self.device)
            # post_quant_conv (1x1, embed_dim -> z_channels): tiny; SIMT engine, output into conv_in's padded buffer
            wpq = pk.t("post_quant_conv.weight")
            self.pq_w = torch.empty((self.z_channels, self.embed_dim), dtype=self.dt, device=self.device)
            ops.repack_conv_weight(wpq, self.pq_w)
            self.pq_w_scaled = torch.empty_like(self.pq_w)       # decode_first_stage's z / scale_factor, folded
            ops.repack_conv_weight(wpq * (1.0 / scale_factor), self.pq_w_scaled)
            self.scale_factor = float(scale_factor)
            self.pq_b = pk.vec("post_quant_conv.bias")
            self.conv_in = pk.conv("decoder.conv_in", tc_small=True)
            block_in = self.ch * self.ch_mult[-1]
            self.mid1 = self._res(pk, "decoder.mid.block_1", block_in, block_in)
            a = "decoder.mid.attn_1"
            self.attn_gn = pk.norm(a + ".norm")
            self.attn_qk = pk.stacked_linear([a + ".q", a + ".k"], with_bias=True)      # one [2C, C] GEMM
            self.attn_wv = pk.unblocked(a + ".v")                                       # used as the "activation" of v^T
            self.attn_out = pk.conv(a + ".proj_out")
            # P (v0 + 1 b_v^T) = P v0 + b_v^T  ->  proj_out bias' = W_o b_v + b_o
            wo = pk.t(a + ".proj_out.weight").reshape(block_in, block_in)
            self.attn_out.bias = (wo @ pk.vec(a + ".v.bias") + pk.vec(a + ".proj_out.bias")).contiguous()
            self.attn_ch = block_in
            self.mid2 = self._res(pk, "decoder.mid.block_2", block_in, block_in)
            self.levels = []
            for i_level in reversed(range(len(self.ch_mult))):
                block_out = self.ch * self.ch_mult[i_level]
                blocks = []
                for i_block in range(self.num_res_blocks + 1):
                    blocks.append(self._res(pk, f"decoder.up.{i_level}.block.{i_block}", block_in, block_out))
                    block_in = block_out
                upc = pk.conv(f"decoder.up.{i_level}.upsample.conv") if i_level != 0 else None
                self.levels.append((blocks, upc))
            self.norm_out = pk.norm("decoder.norm_out")
            self.conv_out = pk.conv("decoder.conv_out", tc_small=True)
        self.loaded = True
        return self

    # ---- blocks ------------------------------------------------------------------------------------------------
    def _conv(self, pc: PConv, x, out, B, H, W, **kw):
        return ops.conv2d(x, pc.w, out, B, H, W, ksize=pc.ksize, stride=pc.stride, bias=pc.bias, **kw)

    def _gn(self, n: PNorm, x, out, B, HW, act):
        # this decoder's OWN barrier / partial-sum scratch: a decode on a side stream must not share the cooperative
        # GroupNorm kernel's barrier words with a sampler running on the same device
        from ._lib import lib
        scr = self.buf("gn.coop", 1, int(lib.pd_group_norm_scratch_floats(B)), torch.float32, zero=True)
        return ops.group_norm(x, out, n.gamma, n.beta, B, HW, eps=1e-6, act=act, scratch=scr.view(-1))

    def _res_block(self, r: _VRes, x, out, B, H, W):
        """ResnetBlock.forward with temb=None (model.py:123-145)."""
        M = B * H * W
        g1 = self.buf("t_gn", M, r.cin)
        self._gn(r.gn1, x, g1, B, H * W, PD_ACT_SILU)
        h1 = self.buf("t_h1", M, r.cout)
        self._conv(r.conv1, g1, h1, B, H, W)
        g2 = self.buf("t_gn", M, r.cout)
        self._gn(r.gn2, h1, g2, B, H * W, PD_ACT_SILU)
        if r.has_skip:
            self._conv(r.conv2, g2, out, B, H, W, x2=x)
        else:
            self._conv(r.conv2, g2, out, B, H, W, res=x)
        return out

    def _attn_block(self, x, out, B, H, W):
        """AttnBlock.forward (model.py:176-203)."""
        N, Cc = H * W, self.attn_ch
        M = B * N
        g = self.buf("t_gn", M, Cc)
        self._gn(self.attn_gn, x, g, B, N, PD_ACT_NONE)
        qk = self.buf("a_qk", M, 2 * Cc)
        ops.linear(g, self.attn_qk.w, qk, bias=self.attn_qk.bias)
        q = self.buf("a_q", M, Cc)
        k = self.buf("a_k", M, Cc)
        ops.cast2d(qk[:, :Cc], q)            # the score GEMM wants k contiguous (it plays the weight matrix)
        ops.cast2d(qk[:, Cc:], k)
        att = self.buf("a_att", M, Cc)
        s = self.buf("a_s", N, N)
        vt = self.buf("a_vt", Cc, N)
        scale = float(Cc) ** -0.5
        for b in range(B):
            rows = slice(b * N, (b + 1) * N)
            ops.linear(q[rows], k[rows], s)                                   # S = q k^T            [N, N]
            ops.softmax_rows(s, s, scale)                                     # softmax(S * c^-1/2)
            ops.linear(self.attn_wv, g[rows], vt)                             # v0^T = W_v g^T       [C, N]
            ops.linear(s, vt, att[rows])                                      # O = P v0             [N, C]
        self._conv(self.attn_out, att, out, 1, 1, M, res=x)
        return out

    def _up(self, pc: PConv, x, out, B, H, W):
        """Upsample.forward (model.py:60-64): nearest x2 then conv3x3."""
        if self.dt == torch.bfloat16:
            u = self.buf("t_up", B * 4 * H * W, x.shape[1])
            ops.upsample2x(x, u, B, H, W)
            return self._conv(pc, u, out, B, 2 * H, 2 * W)
        return self._conv(pc, x, out, B, H, W, upsample=True)

    # ---- public surface ----------------------------------------------------------------------------------------
    @torch.no_grad()
    @ops.on_device
    def decode(self, z: torch.Tensor, scaled: bool = False) -> torch.Tensor:
        """z [B, embed_dim, h, w] -> image [B, out_ch, 8h, 8w] fp32 (autoencoder.py:88-91).
        ``scaled=True`` decodes ``z / scale_factor`` (what ``decode_first_stage`` passes in)."""
        if not self.loaded:
            raise RuntimeError("AutoencoderKLDecoder: load_state_dict() has not been called")
        B, Cz, H, W = z.shape
        if Cz != self.embed_dim:
            raise ValueError(f"latent has {Cz} channels, the decoder expects {self.embed_dim}")
        z = z.to(device=self.device, dtype=torch.float32).contiguous()
        M = B * H * W
        zin = self.buf("z_in", M, Cz)
        ops.nchw_to_nhwc(z, zin)
        zq_full = self.buf("z_pq", M, self.conv_in.cin_pad, zero=True)       # only the first z_channels columns are written
        zq = zq_full[:, :self.z_channels]
        ops.conv2d(zin, self.pq_w_scaled if scaled else self.pq_w, zq, 1, 1, M, bias=self.pq_b)
        xin = zq_full if self.conv_in.cin_pad != self.z_channels else zq
        cur = self.buf("h_a", M, self.conv_in.cout)
        self._conv(self.conv_in, xin, cur, B, H, W)
        nxt = self.buf("h_b", M, self.mid1.cout)
        cur = self._res_block(self.mid1, cur, nxt, B, H, W)
        nxt = self.buf("h_a", M, self.attn_ch)
        cur = self._attn_block(cur, nxt, B, H, W)
        nxt = self.buf("h_b", M, self.mid2.cout)
        cur = self._res_block(self.mid2, cur, nxt, B, H, W)
        flip = 0
        for blocks, upc in self.levels:
            for r in blocks:
                nxt = self.buf("h_a" if flip == 0 else "h_b", B * H * W, r.cout)
                if nxt.data_ptr() == cur.data_ptr():
                    nxt = self.buf("h_c", B * H * W, r.cout)
                cur = self._res_block(r, cur, nxt, B, H, W)
                flip ^= 1
            if upc is not None:
                nxt = self.buf("h_up", B * 4 * H * W, upc.cout)
                cur = self._up(upc, cur, nxt, B, H, W)
                H, W = 2 * H, 2 * W
        M = B * H * W
        g = self.buf("t_gn", M, cur.shape[1])
        self._gn(self.norm_out, cur, g, B, H * W, PD_ACT_SILU)
        img_full = self.buf("img", M, self.conv_out.cout_pad, torch.float32)
        self._conv(self.conv_out, g, img_full, B, H, W)
        return ops.nhwc_to_nchw(img_full[:, :self.out_ch], B, self.out_ch, H, W)

    __call__ = decode
