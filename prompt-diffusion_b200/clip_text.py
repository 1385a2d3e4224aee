"""Text conditioning on the B200 kernels — the step right before the denoising loop.

Mirrors ``FrozenCLIPEmbedder.forward`` / ``.encode`` (ldm/modules/encoders/modules.py:117-131, ``layer="last"``) FROM
THE TOKEN IDS ON: the reference tokenises with ``CLIPTokenizer`` (vocabulary files, host-side string work — out of
scope) and calls Hugging Face ``CLIPTextModel(input_ids=tokens).last_hidden_state``; this class runs that transformer
(ViT-L/14 text tower: 12 pre-LayerNorm blocks, 12 heads of 64, causal self-attention over 77 tokens, quick-GELU MLP,
final LayerNorm) from the checkpoint's ``cond_stage_model.transformer.text_model.*`` tensors.

Ops: ``pd_embedding_lookup`` -> per block [``pd_layer_norm`` -> fused q|k|v GEMM (the d^-1/2 query scale is applied by
the attention kernel) -> ``pd_attention_causal`` -> out_proj GEMM (+residual) -> ``pd_layer_norm`` -> fc1 GEMM ->
``pd_quick_gelu`` -> fc2 GEMM (+residual)] -> ``pd_layer_norm``.  M = 77 * batch rows, run once per ``sample()``.
"""
from __future__ import annotations

from typing import Dict, List, Mapping

import torch

from . import ops
from .packing import PConv, PNorm, Packer

_MODES = {"bf16": torch.bfloat16, "fp32": torch.float32}
CLIP_PREFIX = "cond_stage_model.transformer."


class _Block:
    __slots__ = ("ln1", "qkv", "out", "ln2", "fc1", "fc2")


class FrozenCLIPTextEncoder:
    def __init__(self, mode: str = "bf16", device="cuda", heads: int = 12, max_length: int = 77):
        if mode not in _MODES:
            raise ValueError(f"mode must be one of {list(_MODES)}")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("prompt_diffusion_b200.FrozenCLIPTextEncoder runs on CUDA only (no CPU fallback)")
        self.mode, self.dt, self.heads, self.max_length = mode, _MODES[mode], heads, max_length
        self.bufs: Dict[tuple, torch.Tensor] = {}
        self.blocks: List[_Block] = []
        self.loaded = False

    def buf(self, name, rows, cols, dtype=None):
        key = (name, rows, cols, dtype or self.dt)
        b = self.bufs.get(key)
        if b is None:
            b = torch.empty((rows, cols), dtype=dtype or self.dt, device=self.device)
            self.bufs[key] = b
        return b

    def load_state_dict(self, sd: Mapping[str, torch.Tensor], prefix: str = CLIP_PREFIX):
        """``sd``: LDM-checkpoint keys ``cond_stage_model.transformer.text_model.*`` (or a bare HF ``CLIPTextModel`` state
        dict with ``prefix=""``)."""
        with torch.cuda.device(self.device):
            pk = Packer(sd, prefix + "text_model.", self.dt, self.device)
            self.tok = pk.vec("embeddings.token_embedding.weight")          # fp32 tables
            self.pos = pk.vec("embeddings.position_embedding.weight")
            self.width = self.tok.shape[1]
            self.blocks = []
            i = 0
            while (prefix + f"text_model.encoder.layers.{i}.layer_norm1.weight") in sd:
                L = f"encoder.layers.{i}."
                b = _Block()
                b.ln1, b.ln2 = pk.norm(L + "layer_norm1"), pk.norm(L + "layer_norm2")
                b.qkv = pk.stacked_linear([L + "self_attn.q_proj", L + "self_attn.k_proj", L + "self_attn.v_proj"], True)
                b.out = pk.stacked_linear([L + "self_attn.out_proj"], True)
                b.fc1 = pk.stacked_linear([L + "mlp.fc1"], True)
                b.fc2 = pk.stacked_linear([L + "mlp.fc2"], True)
                self.blocks.append(b)
                i += 1
            if not self.blocks:
                raise KeyError(f"checkpoint has no '{prefix}text_model.encoder.layers.*' tensors")
            self.final_ln = pk.norm("final_layer_norm")
        self.loaded = True
        return self

    @torch.no_grad()
    @ops.on_device
    def forward(self, tokens: torch.Tensor) -> torch.Tensor:
        """tokens int64 [B, L <= 77] (the tokenizer's ``input_ids``) -> last_hidden_state fp32 [B, L, width]."""
        if not self.loaded:
            raise RuntimeError("FrozenCLIPTextEncoder: load_state_dict() has not been called")
        if tokens.dim() != 2 or tokens.shape[1] > self.pos.shape[0]:
            raise ValueError(f"tokens must be [B, L <= {self.pos.shape[0]}], got {tuple(tokens.shape)}")
        B, L = tokens.shape
        C, H = self.width, self.heads
        d = C // H
        M = B * L
        ids = tokens.to(device=self.device, dtype=torch.int64).contiguous().reshape(-1)
        x = self.buf("x", M, C)
        ops.embedding_lookup(ids, self.tok, self.pos, x, L)
        h = self.buf("h", M, C)
        qkv = self.buf("qkv", M, 3 * C)
        att = self.buf("att", M, C)
        y = self.buf("y", M, C)
        ff = self.buf("ff", M, self.blocks[0].fc1.cout)
        for b in self.blocks:
            ops.layer_norm(x, h, b.ln1.gamma, b.ln1.beta)
            ops.linear(h, b.qkv.w, qkv, bias=b.qkv.bias)
            # HF scales q by d^-1/2 after the bias (modeling_clip CLIPAttention); the kernel's score scale is the same thing
            ops.attention_causal(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], att, B, H, L, d, scale=d ** -0.5)
            ops.linear(att, b.out.w, y, bias=b.out.bias, res=x)
            ops.layer_norm(y, h, b.ln2.gamma, b.ln2.beta)
            ops.linear(h, b.fc1.w, ff, bias=b.fc1.bias)
            ops.quick_gelu(ff, ff)
            ops.linear(ff, b.fc2.w, x, bias=b.fc2.bias, res=y)
        ops.layer_norm(x, h, self.final_ln.gamma, self.final_ln.beta)
        out = torch.empty((M, C), dtype=torch.float32, device=self.device)
        ops.cast2d(h, out)
        return out.reshape(B, L, C)

    encode = forward          # modules.py:130-131
    __call__ = forward
