"""Thin torch-tensor wrappers over the C-ABI (``include/pd_b200.h``).

PyTorch is plumbing only: it owns device memory and the current CUDA stream; every op
below passes raw pointers / pitches to ``libpd_b200.so``.  Activations are 2-D views
``[pixels, channels]`` with unit inner stride and an arbitrary row pitch (so a column
slice of a wider concat buffer is a legal input or output).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import (PD_ACT_GEGLU, PD_ACT_NONE, PD_ACT_SILU, PD_BF16, PD_ENGINE_AUTO, PD_ENGINE_SIMT, PD_ENGINE_TC,
                   PD_F32, ConvParams, check, lib)

_DT = {torch.float32: PD_F32, torch.bfloat16: PD_BF16}


def on_device(fn):
    """Decorator for public entry points of the model classes (anything with ``self.device``): run with the model's
    device current.  The C-ABI launches on the current device's current stream and keeps per-device state keyed by
    ``cudaGetDevice()``, so a ``cuda:1`` model must not be driven while ``cuda:0`` is current."""
    import functools

    @functools.wraps(fn)
    def wrapped(self, *a, **kw):
        dev = torch.device(self.device)
        if dev.type != "cuda" or dev.index is None or torch.cuda.current_device() == dev.index:
            return fn(self, *a, **kw)
        with torch.cuda.device(dev):
            return fn(self, *a, **kw)
    return wrapped


def dt_code(t: torch.Tensor) -> int:
    try:
        return _DT[t.dtype]
    except KeyError:
        raise TypeError(f"unsupported dtype {t.dtype}") from None


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _ld(t: torch.Tensor) -> int:
    """Row pitch (elements) of a 2-D pixel-major view."""
    if t.dim() != 2 or (t.shape[1] > 1 and t.stride(1) != 1):
        raise ValueError(f"expected a 2-D view with unit inner stride, got shape {tuple(t.shape)} "
                         f"strides {t.stride()}")
    return t.stride(0) if t.shape[0] > 1 else max(t.stride(0), t.shape[1])


def _cuda(*ts):
    """Every tensor must live on the CURRENT CUDA device: the library launches on that device's current stream and
    keeps per-device state keyed by ``cudaGetDevice()``.  The model classes enter ``torch.cuda.device(self.device)``
    around their public entry points; direct callers of these ops on a second GPU must do the same."""
    cur = torch.cuda.current_device()
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise ValueError("prompt_diffusion_b200 ops need CUDA tensors (there is no CPU path)")
        if t.device.index != cur:
            raise ValueError(f"tensor on cuda:{t.device.index} but the current device is cuda:{cur}: wrap the call in "
                             f"torch.cuda.device({t.device.index}) (the model classes do this themselves)")


def conv2d(x, w, out, B, H, W, *, ksize=1, stride=1, upsample=False, bias=None, rowvec=None, res=None,
           x2=None, act=PD_ACT_NONE, alpha=1.0, engine=PD_ENGINE_AUTO, ln_stats=None, ln_colsum=None,
           ln_parts=None, ln_parts_out=None, ln_rows=0, ln_eps=1e-5, gn_stats_out=None, gn_ld=0, gn_rec_off=0,
           gn_recs_per_image=0, pad=None, out_strides=None):
    """out[B*Ho*Wo, Cout] = act(alpha*(conv(x) [+ x2 @ Wskip] + bias) + rowvec[b] + res).
    ``act=PD_ACT_GEGLU``: w / bias hold 2F rows (see ``geglu_interleave``), out has F columns.
    ``ln_stats`` / ``ln_colsum``: LayerNorm folded into the layer (see ``fold_layer_norm``); ``ln_parts`` instead of
    ``ln_stats`` takes the row partials a producing launch wrote through ``ln_parts_out``.
    ``gn_stats_out`` (fp32 view at this launch's first channel) receives GroupNorm column records, see
    pd_conv_params.  ``ksize=2`` with ``pad=(py, px)`` and ``out_strides=(sx, sy, sb)`` is one phase of
    Upsample's nearest-x2 + conv3x3 (``out`` then points at the phase's first pixel)."""
    _cuda(x, w, out, bias, rowvec, res, x2, ln_stats, ln_colsum, ln_parts, ln_parts_out, gn_stats_out)
    p = ConvParams()
    p.x, p.x2, p.w, p.bias, p.rowvec, p.res, p.out = _p(x), _p(x2), _p(w), _p(bias), _p(rowvec), _p(res), _p(out)
    p.B, p.H, p.W, p.C = B, H, W, x.shape[1]
    p.C2 = 0 if x2 is None else x2.shape[1]
    # GEGLU epilogue: the GEMM has 2F columns (value | gate, interleaved in blocks of 32), the output F
    p.Cout = out.shape[1] * (2 if act == PD_ACT_GEGLU else 1)
    p.ksize, p.stride, p.upsample = ksize, stride, int(bool(upsample))
    p.ldx = _ld(x)
    p.ldx2 = 0 if x2 is None else _ld(x2)
    p.ldr = 0 if res is None else _ld(res)
    p.ldo = _ld(out)
    p.ldrv = 0 if rowvec is None else _ld(rowvec)
    p.act, p.dtype, p.out_dtype, p.engine, p.alpha = act, dt_code(x), dt_code(out), engine, float(alpha)
    p.w_blocked = int(bool(getattr(w, "_pd_blocked", False)))
    p.ln_stats, p.ln_colsum = _p(ln_stats), _p(ln_colsum)
    p.ln_parts, p.ln_parts_out, p.ln_rows, p.ln_eps = _p(ln_parts), _p(ln_parts_out), int(ln_rows), float(ln_eps)
    p.gn_stats_out, p.gn_ld, p.gn_rec_off, p.gn_recs_per_image = _p(gn_stats_out), int(gn_ld), int(gn_rec_off), int(gn_recs_per_image)
    if pad is not None:
        p.pad_y, p.pad_x = int(pad[0]), int(pad[1])
    if out_strides is not None:
        p.out_sx, p.out_sy, p.out_sb = (int(v) for v in out_strides)
    for t in (ln_parts, ln_parts_out, gn_stats_out):
        if t is not None and t.dtype != torch.float32:
            raise TypeError("ln_parts / ln_parts_out / gn_stats_out are fp32 buffers")
    if ln_stats is not None and (ln_stats.dtype != torch.float32 or ln_stats.shape != (x.shape[0], 2) or
                                 not ln_stats.is_contiguous() or ln_colsum is None or ln_colsum.dtype != torch.float32 or
                                 ln_colsum.numel() != p.Cout):
        raise ValueError("ln_stats must be contiguous fp32 [rows, 2] and ln_colsum fp32 [Cout]")
    if w.dtype != x.dtype or (x2 is not None and x2.dtype != x.dtype):
        raise TypeError("x, x2 and w must share a dtype")
    if res is not None and res.dtype != out.dtype:
        raise TypeError("res must have the output dtype")
    ktot = ksize * ksize * p.C + p.C2
    if w.dim() != 2 or w.shape[0] != p.Cout or w.shape[1] != ktot or not w.is_contiguous():
        raise ValueError(f"weight must be contiguous [{p.Cout}, {ktot}], got {tuple(w.shape)}")
    if x.shape[0] != B * H * W:
        raise ValueError(f"x has {x.shape[0]} pixels, expected {B * H * W}")
    check(lib.pd_conv2d(C.byref(p), _stream()), "pd_conv2d")
    return out


def geglu_interleave(w: torch.Tensor) -> torch.Tensor:
    """Reorder the rows of FeedForward's ``net.0.proj`` weight / bias ([2F, ...]: F value rows then F gate rows,
    attention.py:49-56) into blocks of 32 value rows followed by their 32 gate rows — the layout
    ``conv2d(..., act=PD_ACT_GEGLU)`` expects."""
    F = w.shape[0] // 2
    if w.shape[0] != 2 * F or F % 32 != 0:
        raise ValueError("GEGLU interleave needs 2F rows with F a multiple of 32")
    idx = torch.arange(F, device=w.device).reshape(F // 32, 32)
    perm = torch.cat([idx, idx + F], dim=1).reshape(-1)
    return w.index_select(0, perm).contiguous()


def block_weight(w: torch.Tensor) -> torch.Tensor:
    """[Cout, K] (K contiguous, K % 64 == 0) -> the same storage size laid out k-block-major [K/64][Cout][64] and
    tagged so that ``conv2d`` passes ``w_blocked=1`` (see pd_conv_params).  The returned tensor keeps the logical
    shape [Cout, K]; only the tcgen05 engine can read it."""
    cout, k = w.shape
    if k % 64 != 0:
        raise ValueError("block_weight needs K % 64 == 0")
    out = w.reshape(cout, k // 64, 64).permute(1, 0, 2).contiguous().reshape(cout, k)
    out._pd_blocked = True
    return out


def linear(x, w, out, **kw):
    return conv2d(x, w, out, 1, 1, x.shape[0], **kw)


def repack_conv_weight(w_oihw: torch.Tensor, out: torch.Tensor, cin_pad: Optional[int] = None, k_offset: int = 0):
    """fp32 OIHW (or [out,in]) -> rows of `out` ([Cout, ldk], K-major, tap-major), see pd_repack_conv_weight."""
    _cuda(w_oihw, out)
    if w_oihw.dim() == 2:
        w_oihw = w_oihw[:, :, None, None]
    cout, cin, kh, kw = w_oihw.shape
    cin_pad = cin if cin_pad is None else cin_pad
    w_oihw = w_oihw.contiguous().float()
    check(lib.pd_repack_conv_weight(w_oihw.data_ptr(), out.data_ptr(), cout, cin, kh, kw, cin_pad,
                                    out.stride(0), k_offset, dt_code(out), _stream()), "pd_repack_conv_weight")
    return out


_gn_scratch = {}


def group_norm(x, out, gamma, beta, B, HW, *, groups=32, eps=1e-5, act=PD_ACT_NONE, scratch=None):
    """``scratch`` (zero-filled once, pd_group_norm_scratch_floats(B) floats) holds the cooperative kernel's barrier words:
    pass one per model / stream when several may run GroupNorm concurrently on one device (the model classes pass
    their pool's); the default is one per (device, B), which is only safe for single-stream use."""
    _cuda(x, out, gamma, beta)
    key = (x.device.index, B)
    if scratch is None:
        scratch = _gn_scratch.get(key)
    if scratch is None:
        scratch = torch.zeros(int(lib.pd_group_norm_scratch_floats(B)), dtype=torch.float32, device=x.device)
        _gn_scratch[key] = scratch
    check(lib.pd_group_norm(x.data_ptr(), _ld(x), out.data_ptr(), _ld(out), gamma.data_ptr(), beta.data_ptr(),
                            scratch.data_ptr(), B, HW, x.shape[1], groups, float(eps), act, dt_code(x),
                            dt_code(out), _stream()), "pd_group_norm")
    return out


def ln_parts_floats(rows: int) -> int:
    """Size (floats) of a LayerNorm row-partial buffer behind ``conv2d(..., ln_parts_out=)`` for ``rows`` rows."""
    return int(lib.pd_conv2d_ln_parts_floats(int(rows)))


_gn_sup = {}


def gn_stats_supported(B, Ho, Wo, ksize, stride) -> bool:
    key = (B, Ho, Wo, ksize, stride)
    v = _gn_sup.get(key)
    if v is None:
        v = bool(lib.pd_conv2d_gn_stats_supported(B, Ho, Wo, ksize, stride))
        _gn_sup[key] = v
    return v


def group_norm_apply(x, out, gamma, beta, colstats, stats_ld, recs_per_image, scratch, B, HW, *, groups=32, eps=1e-5,
                     act=PD_ACT_NONE, split=False):
    """GroupNorm(+SiLU) from the (sum, sumsq) records the producing GEMM epilogues emitted: x is streamed once.
    ``split``: out has 2C columns (bf16(y) | bf16(y - bf16(y)))."""
    _cuda(x, out, gamma, beta, colstats, scratch)
    if x.dtype != torch.bfloat16 or out.dtype != torch.bfloat16 or colstats.dtype != torch.float32 or scratch.dtype != torch.float32:
        raise TypeError("group_norm_apply: bf16 tensors, fp32 statistics")
    if scratch.numel() < B * 64:
        raise ValueError("group_norm_apply: scratch needs B * 64 floats")
    check(lib.pd_group_norm_apply(x.data_ptr(), _ld(x), out.data_ptr(), _ld(out), gamma.data_ptr(), beta.data_ptr(),
                                  colstats.data_ptr(), int(stats_ld), int(recs_per_image), scratch.data_ptr(), B, HW,
                                  x.shape[1], groups, float(eps), act, int(bool(split)), _stream()), "pd_group_norm_apply")
    return out


def layer_norm_stats(x, stats, eps=1e-5):
    """stats[row] = (mean, rstd) of LayerNorm over the last dim; consumed by ``conv2d(..., ln_stats=stats)``."""
    _cuda(x, stats)
    if stats.dtype != torch.float32 or stats.shape != (x.shape[0], 2) or not stats.is_contiguous():
        raise ValueError("stats must be contiguous fp32 [rows, 2]")
    check(lib.pd_layer_norm_stats(x.data_ptr(), _ld(x), stats.data_ptr(), x.shape[0], x.shape[1], float(eps),
                                  dt_code(x), _stream()), "pd_layer_norm_stats")
    return stats


def fold_layer_norm(w, bias, gamma, beta, dtype):
    """Fold LayerNorm's affine into the linear layer that consumes it (attention.py:271-275):
        Linear(LN(x))[m, n] = rstd[m] * (sum_k x[m,k] * (W[n,k] gamma[k]) - mean[m] * sum_k W[n,k] gamma[k])
                              + (bias[n] + sum_k W[n,k] beta[k])
    Returns (w_scaled in ``dtype``, bias', colsum) where colsum sums the ROUNDED scaled weights (so the mean
    correction cancels exactly what the tensor cores accumulate).  w: fp32 [N, K]."""
    w32 = w.float()
    ws = (w32 * gamma.float()[None, :]).to(dtype).contiguous()
    colsum = ws.float().sum(dim=1).contiguous()
    b = w32 @ beta.float()
    if bias is not None:
        b = b + bias.float()
    return ws, b.contiguous(), colsum


def layer_norm(x, out, gamma, beta, eps=1e-5):
    _cuda(x, out, gamma, beta)
    check(lib.pd_layer_norm(x.data_ptr(), _ld(x), out.data_ptr(), _ld(out), gamma.data_ptr(), beta.data_ptr(),
                            x.shape[0], x.shape[1], float(eps), dt_code(x), _stream()), "pd_layer_norm")
    return out


def geglu(x, out):
    _cuda(x, out)
    check(lib.pd_geglu(x.data_ptr(), _ld(x), out.data_ptr(), _ld(out), x.shape[0], out.shape[1], dt_code(x),
                       _stream()), "pd_geglu")
    return out


def attention_causal(q, k, v, out, B, heads, N, d, scale=None):
    """Causal self-attention over N tokens (CLIP text tower): query i attends keys <= i."""
    _cuda(q, k, v, out)
    scale = d ** -0.5 if scale is None else scale
    check(lib.pd_attention_causal(q.data_ptr(), _ld(q), k.data_ptr(), _ld(k), v.data_ptr(), _ld(v), out.data_ptr(),
                                  _ld(out), B, heads, N, d, float(scale), dt_code(q), _stream()), "pd_attention_causal")
    return out


def embedding_lookup(ids, tok, pos, out, L):
    """out[r] = tok[ids[r]] + pos[r % L]; ids int64 [rows], tok / pos fp32 tables, out [rows, C] in its own dtype."""
    _cuda(ids, tok, pos, out)
    if ids.dtype != torch.int64 or not ids.is_contiguous() or tok.dtype != torch.float32 or pos.dtype != torch.float32 \
            or not tok.is_contiguous() or not pos.is_contiguous() or tok.shape[1] != out.shape[1] or pos.shape[0] < L:
        raise TypeError("embedding_lookup wants contiguous int64 ids and contiguous fp32 [vocab, C] / [>=L, C] tables")
    check(lib.pd_embedding_lookup(ids.data_ptr(), tok.data_ptr(), pos.data_ptr(), out.data_ptr(), _ld(out),
                                  ids.numel(), L, out.shape[1], tok.shape[0], dt_code(out), _stream()),
          "pd_embedding_lookup")
    return out


def quick_gelu(x, out):
    _cuda(x, out)
    check(lib.pd_quick_gelu(x.data_ptr(), _ld(x), out.data_ptr(), _ld(out), x.shape[0], x.shape[1], dt_code(x),
                            _stream()), "pd_quick_gelu")
    return out


def softmax_rows(x, out, scale=1.0):
    """out[r, :] = softmax(x[r, :] * scale) over a 2-D (possibly pitched) score matrix (model.py:190-193)."""
    _cuda(x, out)
    check(lib.pd_softmax_rows(x.data_ptr(), _ld(x), out.data_ptr(), _ld(out), x.shape[0], x.shape[1], float(scale),
                              dt_code(x), _stream()), "pd_softmax_rows")
    return out


def attention(q, k, v, out, B, heads, Nq, Nk, d, scale=None, engine=0):
    _cuda(q, k, v, out)
    scale = d ** -0.5 if scale is None else scale
    check(lib.pd_attention_ex(q.data_ptr(), _ld(q), k.data_ptr(), _ld(k), v.data_ptr(), _ld(v), out.data_ptr(),
                              _ld(out), B, heads, Nq, Nk, d, float(scale), dt_code(q), engine, _stream()),
          "pd_attention")
    return out


_freq_tables = {}


def _freqs(dim: int, max_period: float, device) -> torch.Tensor:
    """Host-built frequency table, the way the reference builds it (util.py:165-167), uploaded once."""
    import math
    key = (dim, float(max_period), str(device))
    f = _freq_tables.get(key)
    if f is None:
        half = dim // 2
        f = torch.exp(-math.log(max_period) * torch.arange(start=0, end=half, dtype=torch.float32) / half).to(device)
        _freq_tables[key] = f
    return f


def timestep_embedding(t: torch.Tensor, out: torch.Tensor, max_period: float = 10000.0, host_freqs: bool = True):
    _cuda(t, out)
    if t.dtype != torch.int64:
        raise TypeError("timesteps must be int64")
    fr = _freqs(out.shape[1], max_period, out.device) if host_freqs else None
    check(lib.pd_timestep_embedding(t.data_ptr(), _p(fr), out.data_ptr(), _ld(out), t.shape[0], out.shape[1],
                                    float(max_period), dt_code(out), _stream()), "pd_timestep_embedding")
    return out


def silu(x, out):
    _cuda(x, out)
    check(lib.pd_silu(x.data_ptr(), out.data_ptr(), x.numel(), dt_code(x), _stream()), "pd_silu")
    return out


def nchw_to_nhwc(x: torch.Tensor, out: torch.Tensor, accumulate: bool = False):
    """x fp32 [B,C,H,W] contiguous -> out [B*H*W, >=C] (dtype of out)."""
    _cuda(x, out)
    if x.dtype != torch.float32 or not x.is_contiguous():
        raise TypeError("nchw_to_nhwc wants a contiguous fp32 NCHW tensor")
    B, Cc, H, W = x.shape
    check(lib.pd_nchw_to_nhwc(x.data_ptr(), out.data_ptr(), _ld(out), B, Cc, H, W, dt_code(out),
                              int(accumulate), _stream()), "pd_nchw_to_nhwc")
    return out


def nchw_to_nhwc_split(x: torch.Tensor, out: torch.Tensor):
    """x fp32 [B,C,H,W] -> bf16 out[:, 0:C] = hi, [C:2C] = lo, [2C:3C] = hi (see pd_nchw_to_nhwc_split)."""
    _cuda(x, out)
    if x.dtype != torch.float32 or not x.is_contiguous() or out.dtype != torch.bfloat16:
        raise TypeError("nchw_to_nhwc_split wants a contiguous fp32 NCHW tensor and a bf16 output")
    B, Cc, H, W = x.shape
    check(lib.pd_nchw_to_nhwc_split(x.data_ptr(), out.data_ptr(), _ld(out), B, Cc, H, W, _stream()),
          "pd_nchw_to_nhwc_split")
    return out


def add2d(x, y, out):
    _cuda(x, y, out)
    if not (x.dtype == y.dtype == out.dtype == torch.float32):
        raise TypeError("add2d is fp32")
    check(lib.pd_add2d(x.data_ptr(), _ld(x), y.data_ptr(), _ld(y), out.data_ptr(), _ld(out), x.shape[0], x.shape[1],
                       _stream()), "pd_add2d")
    return out


def nhwc_to_nchw(x: torch.Tensor, B, Cc, H, W, scale: float = 1.0, out: Optional[torch.Tensor] = None):
    _cuda(x)
    if out is None:
        out = torch.empty((B, Cc, H, W), dtype=torch.float32, device=x.device)
    check(lib.pd_nhwc_to_nchw(x.data_ptr(), _ld(x), out.data_ptr(), B, Cc, H, W, dt_code(x), float(scale),
                              _stream()), "pd_nhwc_to_nchw")
    return out


def cast2d(x, out):
    _cuda(x, out)
    check(lib.pd_cast2d(x.data_ptr(), _ld(x), dt_code(x), out.data_ptr(), _ld(out), dt_code(out), x.shape[0],
                        x.shape[1], _stream()), "pd_cast2d")
    return out


def upsample2x(x, out, B, H, W):
    _cuda(x, out)
    check(lib.pd_upsample2x(x.data_ptr(), _ld(x), out.data_ptr(), _ld(out), B, H, W, x.shape[1], dt_code(x),
                            _stream()), "pd_upsample2x")
    return out


def cfg_ddim_step(eps_uncond, eps_cond, x, noise, coef, x_prev, pred_x0=None, e_t=None):
    """Fused CFG + DDIM update; all fp32, same layout; coef = 6 device floats."""
    _cuda(eps_cond, x, coef, x_prev)
    for t in (eps_uncond, eps_cond, x, noise, x_prev, pred_x0, e_t):
        if t is not None and (t.dtype != torch.float32 or not t.is_contiguous()):
            raise TypeError("cfg_ddim_step wants contiguous fp32 tensors")
    check(lib.pd_cfg_ddim_step(_p(eps_uncond), eps_cond.data_ptr(), x.data_ptr(), _p(noise), coef.data_ptr(),
                               x_prev.data_ptr(), _p(pred_x0), _p(e_t), x.numel(), _stream()), "pd_cfg_ddim_step")
    return x_prev
