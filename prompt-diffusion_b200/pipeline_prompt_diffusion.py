"""diffusers-style denoising loop over the B200 kernels: the part of
``PromptDiffusionPipeline.__call__`` between prompt encoding and VAE decode
(pipeline_prompt_diffusion.py:1195-1290) with the call contracts of its three collaborators:

* ``controlnet(sample, t, encoder_hidden_states=, controlnet_query_cond=, controlnet_cond=, conditioning_scale=,
  guess_mode=, return_dict=False)``  -> ``prompt_diffusion_b200.PromptDiffusionControlNetModel`` (:1237-1246);
* ``unet(sample, t, encoder_hidden_states=, down_block_additional_residuals=, mid_block_additional_residual=,
  return_dict=False)[0]`` (:1257-1266) -> ``UNet2DConditionShim`` over ``ControlledUnetModel``;
* ``scheduler.set_timesteps / scale_model_input / step`` (:1214-1215, :1274) -> ``DDIMScheduler`` below.

``diffusers`` is a third-party dependency that is neither vendored in the reference nor installed here (floor
``0.33.0.dev0``, train_promptdiffusion_sd15.py:64), so ``DDIMScheduler`` is RESTATED from its published algorithm
(Song et al. DDIM, eq. 12, with the Stable-Diffusion-1.5 scheduler config: scaled-linear betas 0.00085..0.012,
``steps_offset=1``, ``timestep_spacing="leading"``, ``clip_sample=False``, ``set_alpha_to_one=False``) — parity
against diffusers itself is UNPINNED.  What IS pinned: with that config the scheduler visits the timesteps
981, 961, ..., 1 and uses alpha_cumprod[0] after the last one, i.e. exactly ``cldm/ddim_hacked.py``'s schedule, so the
loop must reproduce ``DDIMSampler.sample`` (whose parity is pinned to the reference) — tests/test_pipeline_gpu.py.

Prompt encoding and image decoding are the neighbours in ``clip_text.py`` / ``autoencoder.py``; tokenisation and PIL
pre/post-processing are host-side work outside this path: the loop takes tensors.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence, Union

import numpy as np
import torch

from .cldm.cldm import ControlledUnetModel
from .promptdiffusioncontrolnet import PromptDiffusionControlNetModel


_FUSED_STEP = True


def fused_step_enabled() -> bool:
    return _FUSED_STEP


def set_fused_step(on: bool) -> bool:
    """Switch the pipeline's fused step (taken when both shims share a buffer pool) on/off; off = the call-by-call
    route through ``controlnet(...)`` / ``unet(...)`` / ``scheduler.step(...)``.  Returns the previous setting."""
    global _FUSED_STEP
    prev, _FUSED_STEP = _FUSED_STEP, bool(on)
    return prev


class DDIMScheduler:
    """Restatement of ``diffusers.DDIMScheduler`` for epsilon prediction (see the module docstring)."""

    order = 1
    init_noise_sigma = 1.0

    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.00085, beta_end: float = 0.012,
                 beta_schedule: str = "scaled_linear", clip_sample: bool = False, set_alpha_to_one: bool = False,
                 steps_offset: int = 1, timestep_spacing: str = "leading", prediction_type: str = "epsilon"):
        if beta_schedule != "scaled_linear" or clip_sample or prediction_type != "epsilon":
            raise NotImplementedError("only the Stable-Diffusion-1.5 DDIM configuration is restated")
        if timestep_spacing not in ("leading", "trailing"):
            raise NotImplementedError(f"timestep_spacing={timestep_spacing}")
        betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.num_train_timesteps, self.steps_offset, self.timestep_spacing = num_train_timesteps, steps_offset, timestep_spacing
        self.num_inference_steps: Optional[int] = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy().astype(np.int64))

    def set_timesteps(self, num_inference_steps: int, device=None):
        if num_inference_steps > self.num_train_timesteps:
            raise ValueError("num_inference_steps cannot exceed num_train_timesteps")
        self.num_inference_steps = num_inference_steps
        if self.timestep_spacing == "leading":
            ratio = self.num_train_timesteps // num_inference_steps
            ts = (np.arange(0, num_inference_steps) * ratio).round()[::-1].copy().astype(np.int64) + self.steps_offset
        else:
            ratio = self.num_train_timesteps / num_inference_steps
            ts = np.round(np.arange(self.num_train_timesteps, 0, -ratio)).astype(np.int64) - 1
        self.timesteps = torch.from_numpy(ts).to(device)

    def scale_model_input(self, sample: torch.Tensor, timestep=None) -> torch.Tensor:
        return sample

    def step(self, model_output: torch.Tensor, timestep: Union[int, torch.Tensor], sample: torch.Tensor,
             eta: float = 0.0, generator: Optional[torch.Generator] = None, return_dict: bool = False):
        if self.num_inference_steps is None:
            raise ValueError("set_timesteps() has not been called")
        t = int(timestep)
        prev_t = t - self.num_train_timesteps // self.num_inference_steps
        a_t = float(self.alphas_cumprod[t])
        a_prev = float(self.alphas_cumprod[prev_t]) if prev_t >= 0 else float(self.final_alpha_cumprod)
        pred_x0 = (sample - (1.0 - a_t) ** 0.5 * model_output) / a_t ** 0.5
        var = (1.0 - a_prev) / (1.0 - a_t) * (1.0 - a_t / a_prev)
        std = eta * var ** 0.5
        prev = a_prev ** 0.5 * pred_x0 + (1.0 - a_prev - std ** 2) ** 0.5 * model_output
        if eta > 0:
            noise = torch.randn(model_output.shape, generator=generator, device=model_output.device,
                                dtype=model_output.dtype)
            prev = prev + std * noise
        return (prev,) if not return_dict else {"prev_sample": prev, "pred_original_sample": pred_x0}


    def step_coefficients(self, eta: float = 0.0):
        """Per timestep of ``self.timesteps``: (alpha_t, alpha_prev, sigma_t, sqrt(1 - alpha_t)) — the four scalars ``step``
        uses, in the row format of ``pd_cfg_ddim_step`` (the fused CFG + DDIM update kernel)."""
        rows = []
        for t in [int(v) for v in self.timesteps.cpu()]:
            prev_t = t - self.num_train_timesteps // self.num_inference_steps
            a_t = float(self.alphas_cumprod[t])
            a_prev = float(self.alphas_cumprod[prev_t]) if prev_t >= 0 else float(self.final_alpha_cumprod)
            std = eta * ((1.0 - a_prev) / (1.0 - a_t) * (1.0 - a_t / a_prev)) ** 0.5
            rows.append((a_t, a_prev, std, (1.0 - a_t) ** 0.5))
        return rows


class UNet2DConditionShim:
    """The slice of ``UNet2DConditionModel.__call__`` the pipeline uses (:1257-1266), over ``ControlledUnetModel``
    (ldm ``forward(x, timesteps, context, control)``: control = 12 down residuals + the mid residual, consumed by pop)."""

    def __init__(self, unet: ControlledUnetModel):
        self.unet = unet
        self.device = unet.device
        self.config = type("Config", (), {"in_channels": unet.cfg.in_channels, "time_cond_proj_dim": None})()

    @torch.no_grad()
    def __call__(self, sample, timestep, encoder_hidden_states, timestep_cond=None, cross_attention_kwargs=None,
                 down_block_additional_residuals: Optional[Sequence[torch.Tensor]] = None,
                 mid_block_additional_residual: Optional[torch.Tensor] = None, return_dict: bool = False):
        if timestep_cond is not None or cross_attention_kwargs is not None:
            raise NotImplementedError("timestep_cond / cross_attention_kwargs are not part of the SD1.5 prompt-diffusion path")
        t = timestep if torch.is_tensor(timestep) else torch.tensor([timestep], dtype=torch.int64, device=sample.device)
        t = t.reshape(-1).to(torch.int64).expand(sample.shape[0])
        control = None
        if down_block_additional_residuals is not None:
            control = list(down_block_additional_residuals) + [mid_block_additional_residual]
        out = self.unet.forward(sample, t, encoder_hidden_states, control)
        return (out,) if not return_dict else {"sample": out}


class PromptDiffusionPipeline:
    """Denoising loop of the reference pipeline (steps 5-8 of ``__call__``) on tensors."""

    def __init__(self, unet: Union[ControlledUnetModel, UNet2DConditionShim], controlnet: PromptDiffusionControlNetModel,
                 scheduler: Optional[DDIMScheduler] = None, vae=None, vae_scale_factor: int = 8):
        self.unet = unet if isinstance(unet, UNet2DConditionShim) else UNet2DConditionShim(unet)
        self.controlnet = controlnet
        self.scheduler = scheduler if scheduler is not None else DDIMScheduler()
        self.vae = vae                       # optional AutoencoderKLDecoder (output_type="pt")
        self.vae_scale_factor = vae_scale_factor
        self.device = self.unet.device
        # Both shims over nets of ONE buffer pool (``from_ldm``, or nets built with the same ``pool=``): the loop runs the
        # fused step of the ldm path — zero-conv epilogues adding onto the UNet's skips in place, no NCHW round trips of the
        # 13 residuals, CFG + scheduler step in one kernel, the whole step replayed as a CUDA graph.
        self._fused = self._sampler = None
        u, c = self.unet.unet, self.controlnet.net
        if u.pool is c.pool and u.mode == c.mode and isinstance(self.scheduler, DDIMScheduler):
            from .cldm.cldm import ControlLDM
            from .cldm.ddim_hacked import DDIMSampler
            self._fused = ControlLDM.from_nets(u, c)
            self._sampler = DDIMSampler(self._fused)

    @classmethod
    def from_ldm(cls, model, scheduler: Optional[DDIMScheduler] = None, vae=None) -> "PromptDiffusionPipeline":
        """Pipeline over the two nets of a loaded ``ControlLDM`` (shared weights and buffer pool -> fused step)."""
        cn = PromptDiffusionControlNetModel(model.cfg, model.mode, model.device, net=model.control_model)
        return cls(model.model.diffusion_model, cn, scheduler, vae if vae is not None else model.first_stage_model)

    def _fused_loop(self, latents, embeds, image, image_pair, do_cfg, guidance_scale, eta, generator, keep,
                    controlnet_conditioning_scale, callback_on_step_end, callback, callback_steps):
        """Steps 8 of ``__call__`` (:1209-1290) for the plain case (no guess_mode): per step ONE graph replay of
        controlnet -> unet -> ``e_u + s (e_c - e_u)`` -> ``DDIMScheduler.step`` (cldm/ddim_hacked.py's fused step, fed
        with this scheduler's coefficients)."""
        sch, dev, B = self.scheduler, self.device, latents.shape[0]
        ts = [int(v) for v in sch.timesteps.cpu()]
        rows = [list(r) + [guidance_scale if do_cfg else 1.0, 1.0] for r in sch.step_coefficients(eta)]
        coef = torch.tensor(rows, dtype=torch.float32).to(dev)
        conds = {"c_crossattn": [embeds], "example_pair": [image_pair], "query": [image]}
        old_scales = list(self._fused.control_scales)
        try:
            for i, t in enumerate(ts):
                self._fused.control_scales = [controlnet_conditioning_scale * keep[i]] * 13
                noise = None
                if eta > 0:
                    noise = torch.randn(latents.shape, generator=generator, device=latents.device, dtype=latents.dtype)
                t_vec = torch.full((B,), t, device=dev, dtype=torch.int64)
                latents, _ = self._sampler._plain_step(latents, t_vec, conds, conds if do_cfg else None, do_cfg,
                                                       coef[i], noise)
                if callback_on_step_end is not None:
                    outs = callback_on_step_end(self, i, sch.timesteps[i], {"latents": latents})
                    if isinstance(outs, dict):
                        latents = outs.pop("latents", latents)
                if callback is not None and i % callback_steps == 0:
                    callback(i, sch.timesteps[i], latents)
        finally:
            self._fused.control_scales = old_scales
        return latents

    @torch.no_grad()
    def __call__(self, prompt_embeds: torch.Tensor, image: torch.Tensor, image_pair: torch.Tensor,
                 negative_prompt_embeds: Optional[torch.Tensor] = None, num_inference_steps: int = 50,
                 guidance_scale: float = 7.5, eta: float = 0.0, generator: Optional[torch.Generator] = None,
                 latents: Optional[torch.Tensor] = None, controlnet_conditioning_scale: float = 1.0,
                 guess_mode: bool = False, control_guidance_start: float = 0.0, control_guidance_end: float = 1.0,
                 output_type: str = "latent", callback_on_step_end: Optional[Callable] = None,
                 callback: Optional[Callable] = None, callback_steps: int = 1):
        """``image`` = query condition [B,3,H,W], ``image_pair`` = example pair [B,6,H,W] (already pre-processed to
        [0,1] tensors, prepare_image :781-812), ``prompt_embeds`` / ``negative_prompt_embeds`` = [B,77,768].
        Returns latents (``output_type="latent"``) or the decoded image tensor (``"pt"``, needs ``vae``)."""
        dev = self.device
        do_cfg = guidance_scale > 1.0                                                   # :825-826
        if do_cfg and negative_prompt_embeds is None:
            raise ValueError("classifier-free guidance needs negative_prompt_embeds")
        B = prompt_embeds.shape[0]
        H, W = image.shape[-2:]
        image, image_pair = image.to(dev, torch.float32), image_pair.to(dev, torch.float32)
        fused = self._fused is not None and not guess_mode and fused_step_enabled()
        if do_cfg and not guess_mode and not fused:                                     # prepare_image :808-810
            image, image_pair = torch.cat([image] * 2), torch.cat([image_pair] * 2)     # (fused: B hints, tiled by the net)
        embeds = prompt_embeds.to(dev, torch.float32)
        if do_cfg:
            embeds = torch.cat([negative_prompt_embeds.to(dev, torch.float32), embeds])  # [uncond, cond] :1079-1080
        self.scheduler.set_timesteps(num_inference_steps, device=dev)                   # :1163
        timesteps = self.scheduler.timesteps
        shape = (B, self.unet.config.in_channels, H // self.vae_scale_factor, W // self.vae_scale_factor)
        if latents is None:
            latents = torch.randn(shape, generator=generator, device=generator.device if generator is not None else dev)
        latents = latents.to(dev, torch.float32) * self.scheduler.init_noise_sigma      # prepare_latents :690-709
        keep = [1.0 - float(i / len(timesteps) < control_guidance_start or (i + 1) / len(timesteps) > control_guidance_end)
                for i in range(len(timesteps))]                                         # :1196-1202
        if fused:
            latents = self._fused_loop(latents, embeds, image, image_pair, do_cfg, guidance_scale, eta, generator, keep,
                                       controlnet_conditioning_scale, callback_on_step_end, callback, callback_steps)
            timesteps = []
        for i, t in enumerate(timesteps):
            x_in = torch.cat([latents] * 2) if do_cfg else latents                      # :1220-1221
            x_in = self.scheduler.scale_model_input(x_in, t)
            if guess_mode and do_cfg:                                                   # :1224-1231
                c_x, c_emb = self.scheduler.scale_model_input(latents, t), embeds.chunk(2)[1]
            else:
                c_x, c_emb = x_in, embeds
            cond_scale = controlnet_conditioning_scale * keep[i]                        # :1233-1240
            down, mid = self.controlnet(c_x, t, encoder_hidden_states=c_emb, controlnet_query_cond=image,
                                        controlnet_cond=image_pair, conditioning_scale=cond_scale,
                                        guess_mode=guess_mode, return_dict=False)
            if guess_mode and do_cfg:                                                   # :1248-1254
                down = [torch.cat([torch.zeros_like(d), d]) for d in down]
                mid = torch.cat([torch.zeros_like(mid), mid])
            noise_pred = self.unet(x_in, t, encoder_hidden_states=embeds, down_block_additional_residuals=down,
                                   mid_block_additional_residual=mid, return_dict=False)[0]
            if do_cfg:                                                                  # :1269-1271
                e_u, e_c = noise_pred.chunk(2)
                noise_pred = e_u + guidance_scale * (e_c - e_u)
            latents = self.scheduler.step(noise_pred, t, latents, eta=eta, generator=generator, return_dict=False)[0]
            if callback_on_step_end is not None:                                        # :1276-1284
                outs = callback_on_step_end(self, i, t, {"latents": latents})
                if isinstance(outs, dict):
                    latents = outs.pop("latents", latents)
            if callback is not None and i % callback_steps == 0:                        # :1287-1290
                callback(i, t, latents)
        if output_type == "latent":
            return latents
        if self.vae is None:
            raise ValueError("output_type != 'latent' needs a first-stage decoder (vae=AutoencoderKLDecoder)")
        img = self.vae.decode(latents, scaled=True)                                     # vae.decode(latents / scaling_factor)
        return (img / 2 + 0.5).clamp(0, 1) if output_type == "pt" else img
