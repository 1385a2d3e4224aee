"""ORACLE — test infrastructure, not product code.

A plain-PyTorch fp32 *restatement* of the reference's denoising hot path
(Prompt-Diffusion ``ControlLDM.apply_model`` inside ``DDIMSampler``), written
functionally over a reference-format ``state_dict``.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import this module; the product package never does.

Parity status: PINNED.  ``tests/golden/make_golden.py`` imports the reference
itself (``/root/reference``) in the dev container, loads the same procedural
checkpoint into the reference's own ``ControlLDM`` and stores its outputs in
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks this file against
those fixtures (the reference ships no tests or golden vectors of its own —
SURVEY.md section 4).

Every function cites the reference file:line it follows (paths relative to the
reference root).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

UNET = "model.diffusion_model."
CTRL = "control_model."

# (cout, stride) of the 7 non-final hint-encoder convs — cldm/cldm.py:147-181
_HINT_STACK = ((16, 1), (16, 1), (32, 2), (32, 1), (96, 2), (96, 1), (256, 2))


# --------------------------------------------------------------------------
# schedule (host side)
# --------------------------------------------------------------------------
def make_beta_schedule_linear(n_timestep: int, linear_start: float, linear_end: float) -> np.ndarray:
    """ldm/modules/diffusionmodules/util.py:21-25 ("linear")."""
    betas = torch.linspace(linear_start ** 0.5, linear_end ** 0.5, n_timestep, dtype=torch.float64) ** 2
    return betas.numpy()


def register_schedule(timesteps=1000, linear_start=0.00085, linear_end=0.012) -> Dict[str, torch.Tensor]:
    """ldm/models/diffusion/ddpm.py:138-178 — the buffers the sampler reads."""
    betas = make_beta_schedule_linear(timesteps, linear_start, linear_end)
    alphas = 1.0 - betas
    alphas_cumprod = np.cumprod(alphas, axis=0)
    alphas_cumprod_prev = np.append(1.0, alphas_cumprod[:-1])
    f32 = lambda a: torch.tensor(a, dtype=torch.float32)
    return {"betas": f32(betas), "alphas_cumprod": f32(alphas_cumprod),
            "alphas_cumprod_prev": f32(alphas_cumprod_prev),
            "sqrt_one_minus_alphas_cumprod": f32(np.sqrt(1.0 - alphas_cumprod))}


def make_ddim_timesteps(num_ddim_timesteps: int, num_ddpm_timesteps: int) -> np.ndarray:
    """ldm/modules/diffusionmodules/util.py:46-60 ("uniform")."""
    c = num_ddpm_timesteps // num_ddim_timesteps
    return np.asarray(list(range(0, num_ddpm_timesteps, c))) + 1


def make_ddim_sampling_parameters(alphacums: torch.Tensor, ddim_timesteps: np.ndarray, eta: float):
    """ldm/modules/diffusionmodules/util.py:63-74.  ``alphas``/``sigmas`` come out
    as torch fp32, ``alphas_prev`` as a numpy array (fp64 holding fp32 values)."""
    alphas = alphacums[ddim_timesteps]
    alphas_prev = np.asarray([alphacums[0]] + alphacums[ddim_timesteps[:-1]].tolist())
    sigmas = eta * np.sqrt((1 - alphas_prev) / (1 - alphas) * (1 - alphas / alphas_prev))
    return sigmas, alphas, alphas_prev


def make_schedule(sched: Dict[str, torch.Tensor], S: int, eta: float = 0.0):
    """cldm/ddim_hacked.py:23-52."""
    n = sched["alphas_cumprod"].shape[0]
    ts = make_ddim_timesteps(S, n)
    sigmas, alphas, alphas_prev = make_ddim_sampling_parameters(sched["alphas_cumprod"].cpu(), ts, eta)
    return {"ddim_timesteps": ts, "ddim_sigmas": sigmas, "ddim_alphas": alphas,
            "ddim_alphas_prev": alphas_prev, "ddim_sqrt_one_minus_alphas": np.sqrt(1.0 - alphas)}


def timestep_embedding(timesteps: torch.Tensor, dim: int, max_period: int = 10000) -> torch.Tensor:
    """ldm/modules/diffusionmodules/util.py:154-174."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half
                      ).to(timesteps.device)
    args = timesteps[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


# --------------------------------------------------------------------------
# blocks
# --------------------------------------------------------------------------
class _Net:
    """Key-prefixed view of a state dict plus the few hyper-parameters."""

    def __init__(self, sd, prefix, cfg, qdtype=None):
        self.sd, self.p, self.cfg, self.q = sd, prefix, cfg, qdtype

    def w(self, key):
        return self.sd[self.p + key]

    def has(self, key):
        return (self.p + key) in self.sd

    # optional operand quantisation (emulates a low-precision GEMM input for error
    # budgeting; None = exact fp32 restatement)
    def qz(self, t):
        return t if self.q is None else t.to(self.q).to(torch.float32)

    def conv(self, x, key, stride=1, padding=0):
        return F.conv2d(self.qz(x), self.qz(self.w(key + ".weight")), self.w(key + ".bias"),
                        stride=stride, padding=padding)

    def linear(self, x, key, bias=True):
        return F.linear(self.qz(x), self.qz(self.w(key + ".weight")),
                        self.w(key + ".bias") if bias else None)

    def gn(self, x, key, eps):
        return F.group_norm(x.float(), 32, self.w(key + ".weight"), self.w(key + ".bias"), eps)

    def ln(self, x, key):
        return F.layer_norm(x, (x.shape[-1],), self.w(key + ".weight"), self.w(key + ".bias"), 1e-5)


def time_embed(net: _Net, timesteps: torch.Tensor) -> torch.Tensor:
    """openaimodel.py:526-531 / cldm.py:131-136 applied as in cldm.py:26-27,303-304."""
    t_emb = timestep_embedding(timesteps, net.cfg.model_channels)
    return net.linear(F.silu(net.linear(t_emb, "time_embed.0")), "time_embed.2")


def res_block(net: _Net, key: str, x: torch.Tensor, emb: torch.Tensor) -> torch.Tensor:
    """ResBlock._forward, openaimodel.py:254-274 (use_scale_shift_norm=False, no up/down)."""
    h = net.conv(F.silu(net.gn(x, key + ".in_layers.0", 1e-5)), key + ".in_layers.2", padding=1)
    emb_out = net.linear(F.silu(emb), key + ".emb_layers.1")
    h = h + emb_out[..., None, None]
    h = net.conv(F.silu(net.gn(h, key + ".out_layers.0", 1e-5)), key + ".out_layers.3", padding=1)
    if net.has(key + ".skip_connection.weight"):
        x = net.conv(x, key + ".skip_connection")
    return x + h


def cross_attention(net: _Net, key: str, x: torch.Tensor, context: Optional[torch.Tensor], heads: int):
    """CrossAttention.forward, ldm/modules/attention.py:163-194 (fp32 sim, no mask)."""
    q = net.linear(x, key + ".to_q", bias=False)
    ctx = x if context is None else context
    k = net.linear(ctx, key + ".to_k", bias=False)
    v = net.linear(ctx, key + ".to_v", bias=False)
    b, n, c = q.shape
    d = c // heads
    split = lambda t: t.reshape(b, t.shape[1], heads, d).permute(0, 2, 1, 3).reshape(b * heads, t.shape[1], d)
    q, k, v = split(q), split(k), split(v)
    out = torch.empty_like(q)
    # per-(batch*head) chunks: same arithmetic, avoids materialising [b*h, n, n] at once
    step = max(1, min(q.shape[0], (1 << 28) // max(1, n * k.shape[1])))
    for s in range(0, q.shape[0], step):
        sim = torch.einsum("bid,bjd->bij", net.qz(q[s:s + step]), net.qz(k[s:s + step])) * (d ** -0.5)
        sim = sim.softmax(dim=-1)
        out[s:s + step] = torch.einsum("bij,bjd->bid", net.qz(sim), net.qz(v[s:s + step]))
    out = out.reshape(b, heads, n, d).permute(0, 2, 1, 3).reshape(b, n, c)
    return net.linear(out, key + ".to_out.0")


def feed_forward(net: _Net, key: str, x: torch.Tensor) -> torch.Tensor:
    """FeedForward with GEGLU, attention.py:49-76 (exact erf GELU)."""
    a, gate = net.linear(x, key + ".net.0.proj").chunk(2, dim=-1)
    return net.linear(a * F.gelu(gate), key + ".net.2")


def spatial_transformer(net: _Net, key: str, x: torch.Tensor, context: torch.Tensor) -> torch.Tensor:
    """SpatialTransformer.forward (use_linear=False, depth 1), attention.py:321-340 +
    BasicTransformerBlock._forward :271-275."""
    b, c, h, w = x.shape
    heads = net.cfg.num_heads
    x_in = x
    x = net.conv(net.gn(x, key + ".norm", 1e-6), key + ".proj_in")
    x = x.permute(0, 2, 3, 1).reshape(b, h * w, c)
    tb = key + ".transformer_blocks.0"
    x = cross_attention(net, tb + ".attn1", net.ln(x, tb + ".norm1"), None, heads) + x
    x = cross_attention(net, tb + ".attn2", net.ln(x, tb + ".norm2"), context, heads) + x
    x = feed_forward(net, tb + ".ff", net.ln(x, tb + ".norm3")) + x
    x = x.reshape(b, h, w, c).permute(0, 3, 1, 2)
    return net.conv(x, key + ".proj_out") + x_in


def _block(net: _Net, key: str, h, emb, context):
    """One ``TimestepEmbedSequential`` (openaimodel.py:79-87): children are found
    from the state-dict keys, dispatched by what they are."""
    i = 0
    while True:
        k = f"{key}.{i}"
        if net.has(k + ".in_layers.0.weight"):
            h = res_block(net, k, h, emb)
        elif net.has(k + ".transformer_blocks.0.norm1.weight"):
            h = spatial_transformer(net, k, h, context)
        elif net.has(k + ".op.weight"):                       # Downsample, openaimodel.py:157-159
            h = net.conv(h, k + ".op", stride=2, padding=1)
        elif net.has(k + ".conv.weight"):                     # Upsample, openaimodel.py:108-118
            h = net.conv(F.interpolate(h, scale_factor=2, mode="nearest"), k + ".conv", padding=1)
        elif net.has(k + ".weight"):                          # plain conv (input_blocks.0.0)
            kk = net.w(k + ".weight").shape[-1]
            h = net.conv(h, k, padding=kk // 2)
        else:
            break
        i += 1
    if i == 0:
        raise KeyError(f"no layers under {net.p}{key}")
    return h


def _num_blocks(net: _Net, stem: str) -> int:
    n = 0
    while any(k.startswith(f"{net.p}{stem}.{n}.") for k in net.sd):
        n += 1
    return n


def hint_encoder(net: _Net, stem: str, hint: torch.Tensor) -> torch.Tensor:
    """input_hint_block / input_cond_block, cldm/cldm.py:147-181: 8 convs, SiLU between."""
    h = hint
    for i, (_c, stride) in enumerate(_HINT_STACK):
        h = F.silu(net.conv(h, f"{stem}.{2 * i}", stride=stride, padding=1))
    return net.conv(h, f"{stem}.{2 * len(_HINT_STACK)}", padding=1)


def control_net_forward(sd, cfg, x, timesteps, example_pair, query, context, qdtype=None) -> List[torch.Tensor]:
    """ControlNet.forward, cldm/cldm.py:302-325 → 13 tensors."""
    net = _Net(sd, CTRL, cfg, qdtype)
    emb = time_embed(net, timesteps)
    guided_hint = hint_encoder(net, "input_hint_block", example_pair) + \
        hint_encoder(net, "input_cond_block", query)
    outs = []
    h = x.float()
    for i in range(_num_blocks(net, "input_blocks")):
        h = _block(net, f"input_blocks.{i}", h, emb, context)
        if guided_hint is not None:
            h = h + guided_hint
            guided_hint = None
        outs.append(net.conv(h, f"zero_convs.{i}.0"))
    h = _block(net, "middle_block", h, emb, context)
    outs.append(net.conv(h, "middle_block_out.0"))
    return outs


def unet_forward(sd, cfg, x, timesteps, context, control: Optional[list], only_mid_control=False,
                 qdtype=None) -> torch.Tensor:
    """ControlledUnetModel.forward, cldm/cldm.py:23-45 (consumes ``control`` by pop)."""
    net = _Net(sd, UNET, cfg, qdtype)
    emb = time_embed(net, timesteps)
    hs = []
    h = x.float()
    for i in range(_num_blocks(net, "input_blocks")):
        h = _block(net, f"input_blocks.{i}", h, emb, context)
        hs.append(h)
    h = _block(net, "middle_block", h, emb, context)
    if control is not None:
        h = h + control.pop()
    for i in range(_num_blocks(net, "output_blocks")):
        if only_mid_control or control is None:
            h = torch.cat([h, hs.pop()], dim=1)
        else:
            h = torch.cat([h, hs.pop() + control.pop()], dim=1)
        h = _block(net, f"output_blocks.{i}", h, emb, context)
    h = F.silu(net.gn(h, "out.0", 1e-5))
    return net.conv(h, "out.2", padding=1)


def apply_model(sd, cfg, x_noisy, t, cond: dict, control_scales: Optional[Sequence[float]] = None,
                only_mid_control: bool = False, qdtype=None) -> torch.Tensor:
    """ControlLDM.apply_model, cldm/cldm.py:369-382."""
    assert isinstance(cond, dict)
    cond_txt = torch.cat(cond["c_crossattn"], 1)
    assert cond["example_pair"] is not None
    control = control_net_forward(sd, cfg, x_noisy, t, torch.cat(cond["example_pair"], 1),
                                  cond["query"][0], cond_txt, qdtype)
    scales = [1.0] * 13 if control_scales is None else control_scales
    control = [c * s for c, s in zip(control, scales)]
    return unet_forward(sd, cfg, x_noisy, t, cond_txt, control, only_mid_control, qdtype)


# --------------------------------------------------------------------------
# sampler
# --------------------------------------------------------------------------
def p_sample_ddim(sd, cfg, ddim, x, c, t, index, unconditional_guidance_scale=1.0,
                  unconditional_conditioning=None, temperature=1.0, noise=None,
                  control_scales=None, only_mid_control=False, qdtype=None):
    """DDIMSampler.p_sample_ddim, cldm/ddim_hacked.py:180-234 (eps-parameterisation)."""
    b = x.shape[0]
    if unconditional_conditioning is not None:
        x_in, t_in = torch.cat([x] * 2), torch.cat([t] * 2)
        c_in = {k: [torch.cat([unconditional_conditioning[k][i], c[k][i]]) for i in range(len(c[k]))]
                for k in c}
        e_u, e_c = apply_model(sd, cfg, x_in, t_in, c_in, control_scales, only_mid_control, qdtype).chunk(2)
        e_t = e_u + unconditional_guidance_scale * (e_c - e_u)
    else:
        e_t = apply_model(sd, cfg, x, t, c, control_scales, only_mid_control, qdtype)
    full = lambda v: torch.full((b, 1, 1, 1), float(v), device=x.device)
    a_t, a_prev = full(ddim["ddim_alphas"][index]), full(ddim["ddim_alphas_prev"][index])
    sigma_t = full(ddim["ddim_sigmas"][index])
    sqrt_one_minus_at = full(ddim["ddim_sqrt_one_minus_alphas"][index])
    pred_x0 = (x - sqrt_one_minus_at * e_t) / a_t.sqrt()
    dir_xt = (1.0 - a_prev - sigma_t ** 2).sqrt() * e_t
    if noise is None:
        noise = torch.randn(x.shape, device=x.device)
    x_prev = a_prev.sqrt() * pred_x0 + dir_xt + sigma_t * noise * temperature
    return x_prev, pred_x0, e_t


def ddim_sample(sd, cfg, S, shape, conditioning, eta=0.0, x_T=None, unconditional_guidance_scale=1.0,
                unconditional_conditioning=None, log_every_t=100, control_scales=None,
                only_mid_control=False, qdtype=None, return_eps=False, max_steps=None):
    """DDIMSampler.sample + ddim_sampling, cldm/ddim_hacked.py:55-178."""
    sched = register_schedule(getattr(cfg, "timesteps", 1000), cfg.linear_start, cfg.linear_end)
    ddim = make_schedule(sched, S, eta)
    device = x_T.device if x_T is not None else "cpu"
    b = shape[0]
    img = torch.randn(shape, device=device) if x_T is None else x_T
    intermediates = {"x_inter": [img], "pred_x0": [img]}
    eps_trace = []
    time_range = np.flip(ddim["ddim_timesteps"])
    total = time_range.shape[0]
    for i, step in enumerate(time_range):
        if max_steps is not None and i >= max_steps:
            break
        index = total - i - 1
        ts = torch.full((b,), int(step), device=device, dtype=torch.long)
        img, pred_x0, e_t = p_sample_ddim(sd, cfg, ddim, img, conditioning, ts, index,
                                          unconditional_guidance_scale, unconditional_conditioning,
                                          control_scales=control_scales,
                                          only_mid_control=only_mid_control, qdtype=qdtype)
        if return_eps:
            eps_trace.append(e_t)
        if index % log_every_t == 0 or index == total - 1:
            intermediates["x_inter"].append(img)
            intermediates["pred_x0"].append(pred_x0)
    if return_eps:
        intermediates["eps"] = eps_trace
    return img, intermediates
