"""DDIM sampler with classifier-free guidance for dict conditionings — same public surface as the
reference's ``cldm/ddim_hacked.py`` ``DDIMSampler`` (``sample``, ``make_schedule``, ``ddim_sampling``,
``p_sample_ddim``, ``encode``, ``stochastic_encode``, ``decode``; signatures at :11-15, :55-79, :123-129,
:181-184, :237-238, :284, :300-301), driving ``model.apply_model`` once per step.

Differences in mechanism, not in results:
* the CFG combine and the DDIM update (:193, :211-233) are ONE fused kernel (``pd_cfg_ddim_step``) fed by
  a per-step coefficient row that already lives on the device — no per-step device->host scalar reads;
* the [uncond, cond] conditioning batch (:190-191) is concatenated once per ``sample()`` instead of every
  step (it does not depend on the step), which also lets the model reuse its hint / context-K/V caches;
* rarely used branches (v-parameterisation, score corrector, x0 quantisation, noise dropout, inpainting
  mask) keep the reference's unfused arithmetic in torch elementwise ops.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from .. import _lib, ops
from ..schedule import make_ddim_sampling_parameters, make_ddim_timesteps


_STEP_GRAPHS = os.environ.get("PD_B200_STEP_GRAPH", "1") != "0"


def step_graphs_enabled() -> bool:
    return _STEP_GRAPHS


def set_step_graphs(on: bool) -> bool:
    """Switch CUDA-graph replay of the denoising step on/off (off: every kernel is launched eagerly, which
    per-launch profiling needs).  Returns the previous setting."""
    global _STEP_GRAPHS
    prev, _STEP_GRAPHS = _STEP_GRAPHS, bool(on)
    return prev


def _cond_key(conds):
    """Shapes only: the captured step reads the model's hint / context-K/V cache buffers (static addresses),
    which ``prepare_conditioning`` refreshes eagerly whenever the conditioning tensors change."""
    return tuple((k, tuple(t_.shape)) for k in sorted(conds) for t_ in conds[k])


def _graph_key(model, x, t, conds, guided, with_noise):
    return (id(model), tuple(x.shape), str(x.device), bool(guided), bool(with_noise), _cond_key(conds),
            tuple(float(v) for v in model.control_scales), bool(model.only_mid_control),
            model.control_model.w is not None and id(model.control_model.w), id(model.model.diffusion_model.w))


class _StepGraph:
    """One denoising step (apply_model at [uncond, cond] + fused CFG/DDIM update) captured into a CUDA graph
    with static input/output buffers.  Step-invariant caches (hint encoders, context K/V) are filled by the
    eager warm-up run, so the captured graph holds only per-step work."""

    def __init__(self, sampler, x, t, c, c_in, guided, coef, noise):
        self.model = sampler.model
        self.x, self.t = x.clone(), t.clone()
        self.coef = coef.clone()
        self.noise = noise.clone() if noise is not None else None
        self.x_prev, self.pred_x0 = torch.empty_like(x), torch.empty_like(x)
        run = lambda: sampler._eager_step(self.x, self.t, c, c_in, guided, self.coef, self.noise, self.x_prev,
                                          self.pred_x0)
        side = torch.cuda.Stream(device=x.device)
        side.wait_stream(torch.cuda.current_stream(x.device))
        with torch.cuda.stream(side):
            run()                                  # allocates every pool buffer, fills the caches
            run()
        torch.cuda.current_stream(x.device).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        n0 = _lib.lib.pd_launch_count()
        with torch.cuda.graph(self.graph):
            run()
        self.n_kernels = int(_lib.lib.pd_launch_count() - n0)     # library kernels inside the captured step

    def replay(self, x, t, coef, noise, conds):
        self.model.prepare_conditioning(conds)     # no-op unless the conditioning tensors changed
        self.x.copy_(x)
        self.t.copy_(t)
        self.coef.copy_(coef)
        if self.noise is not None:
            self.noise.copy_(noise)
        self.graph.replay()
        _lib.note_graph_replay(self.n_kernels)
        return self.x_prev.clone(), self.pred_x0.clone()


def _noise_like(shape, device, repeat=False):
    """ldm/modules/diffusionmodules/util.py:267-270."""
    if repeat:
        return torch.randn((1, *shape[1:]), device=device).repeat(shape[0], *((1,) * (len(shape) - 1)))
    return torch.randn(shape, device=device)


class DDIMSampler(object):
    def __init__(self, model, schedule="linear", **kwargs):
        super().__init__()
        self.model = model
        self.ddpm_num_timesteps = model.num_timesteps
        self.schedule = schedule
        self._coef_cache = None
        self._graphs = None

    def register_buffer(self, name, attr):
        # the reference moves tensors to "cuda" here (:17-21); buffers follow the model's device instead
        if isinstance(attr, torch.Tensor) and attr.device != torch.device(self.model.device):
            attr = attr.to(self.model.device)
        setattr(self, name, attr)

    def make_schedule(self, ddim_num_steps, ddim_discretize="uniform", ddim_eta=0., verbose=True):
        self.ddim_timesteps = make_ddim_timesteps(ddim_discr_method=ddim_discretize,
                                                  num_ddim_timesteps=ddim_num_steps,
                                                  num_ddpm_timesteps=self.ddpm_num_timesteps, verbose=verbose)
        alphas_cumprod = self.model.alphas_cumprod
        assert alphas_cumprod.shape[0] == self.ddpm_num_timesteps, 'alphas have to be defined for each timestep'
        f32 = lambda x: x.clone().detach().to(torch.float32).to(self.model.device)
        acp_cpu = alphas_cumprod.detach().cpu()
        self.register_buffer('betas', f32(self.model.betas))
        self.register_buffer('alphas_cumprod', f32(alphas_cumprod))
        self.register_buffer('alphas_cumprod_prev', f32(self.model.alphas_cumprod_prev))
        self.register_buffer('sqrt_alphas_cumprod', f32(np.sqrt(acp_cpu)))
        self.register_buffer('sqrt_one_minus_alphas_cumprod', f32(np.sqrt(1. - acp_cpu)))
        self.register_buffer('log_one_minus_alphas_cumprod', f32(np.log(1. - acp_cpu)))
        self.register_buffer('sqrt_recip_alphas_cumprod', f32(np.sqrt(1. / acp_cpu)))
        self.register_buffer('sqrt_recipm1_alphas_cumprod', f32(np.sqrt(1. / acp_cpu - 1)))

        ddim_sigmas, ddim_alphas, ddim_alphas_prev = make_ddim_sampling_parameters(
            alphacums=acp_cpu, ddim_timesteps=self.ddim_timesteps, eta=ddim_eta, verbose=verbose)
        # same container types as the reference: torch, torch, numpy, torch
        self.register_buffer('ddim_sigmas', ddim_sigmas)
        self.register_buffer('ddim_alphas', ddim_alphas)
        self.register_buffer('ddim_alphas_prev', ddim_alphas_prev)
        self.register_buffer('ddim_sqrt_one_minus_alphas', np.sqrt(1. - ddim_alphas))
        sigmas_orig = ddim_eta * torch.sqrt(
            (1 - self.alphas_cumprod_prev) / (1 - self.alphas_cumprod) *
            (1 - self.alphas_cumprod / self.alphas_cumprod_prev))
        self.register_buffer('ddim_sigmas_for_original_num_steps', sigmas_orig)
        self._coef_cache = None

    # ---- per-step coefficient rows on the device ------------------------------------------------------
    def _step_coefs(self, use_original_steps: bool):
        """fp32 host table [T, 4] = (a_t, a_prev, sigma_t, sqrt_one_minus_at), the values p_sample_ddim
        indexes at :206-214."""
        if use_original_steps:
            alphas = self.model.alphas_cumprod
            alphas_prev = self.model.alphas_cumprod_prev
            s1m = self.model.sqrt_one_minus_alphas_cumprod
            # mirrored as written (:209): the reference reads this table from the MODEL — it only exists on the sampler,
            # so a model that does not provide it raises AttributeError here exactly like the reference does
            sigmas = self.model.ddim_sigmas_for_original_num_steps
        else:
            alphas, alphas_prev = self.ddim_alphas, self.ddim_alphas_prev
            s1m, sigmas = self.ddim_sqrt_one_minus_alphas, self.ddim_sigmas
        to_np = lambda a: (a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)).astype(np.float32)
        return np.stack([to_np(alphas), to_np(alphas_prev), to_np(sigmas), to_np(s1m)], axis=1)

    def _coef_row(self, index, use_original_steps, scale, temperature):
        """Device tensor of 6 floats for ``pd_cfg_ddim_step``; the table is uploaded once per schedule and
        (scale, temperature) pair."""
        key = (bool(use_original_steps), float(scale), float(temperature))
        if self._coef_cache is None or self._coef_cache[0] != key:
            tab = self._step_coefs(use_original_steps)
            full = np.concatenate([tab, np.full((tab.shape[0], 1), scale, np.float32),
                                   np.full((tab.shape[0], 1), temperature, np.float32)], axis=1)
            self._coef_cache = (key, torch.from_numpy(np.ascontiguousarray(full)).to(self.model.device), tab)
        return self._coef_cache[1][index], self._coef_cache[2][index]

    @torch.no_grad()
    def sample(self, S, batch_size, shape, conditioning=None, callback=None, normals_sequence=None,
               img_callback=None, quantize_x0=False, eta=0., mask=None, x0=None, temperature=1.,
               noise_dropout=0., score_corrector=None, corrector_kwargs=None, verbose=True, x_T=None,
               log_every_t=100, unconditional_guidance_scale=1., unconditional_conditioning=None,
               dynamic_threshold=None, ucg_schedule=None, **kwargs):
        if conditioning is not None:
            if isinstance(conditioning, dict):
                first = conditioning[list(conditioning.keys())[0]]
                while isinstance(first, list):
                    first = first[0]
                cbs = first.shape[0]
                if cbs != batch_size:
                    print(f"Warning: Got {cbs} conditionings but batch-size is {batch_size}")
            elif isinstance(conditioning, list):
                for ctmp in conditioning:
                    if ctmp.shape[0] != batch_size:
                        print(f"Warning: Got {ctmp.shape[0]} conditionings but batch-size is {batch_size}")
            else:
                if conditioning.shape[0] != batch_size:
                    print(f"Warning: Got {conditioning.shape[0]} conditionings but batch-size is {batch_size}")

        self.make_schedule(ddim_num_steps=S, ddim_eta=eta, verbose=verbose)
        C, H, W = shape
        size = (batch_size, C, H, W)
        if verbose:
            print(f'Data shape for DDIM sampling is {size}, eta {eta}')
        return self.ddim_sampling(conditioning, size, callback=callback, img_callback=img_callback,
                                  quantize_denoised=quantize_x0, mask=mask, x0=x0,
                                  ddim_use_original_steps=False, noise_dropout=noise_dropout,
                                  temperature=temperature, score_corrector=score_corrector,
                                  corrector_kwargs=corrector_kwargs, x_T=x_T, log_every_t=log_every_t,
                                  unconditional_guidance_scale=unconditional_guidance_scale,
                                  unconditional_conditioning=unconditional_conditioning,
                                  dynamic_threshold=dynamic_threshold, ucg_schedule=ucg_schedule)

    _SHARED_KEYS = ("example_pair", "query")

    def _concat_conds(self, c, uc):
        """[uncond, cond] batch for every entry of the conditioning dict (:189-191), built ON THE DEVICE: host
        tensors are uploaded first (one async copy each from pinned memory) and concatenated there — concatenating
        2 x 75 MB of hints on the host cost 70-90 ms per ``sample()`` under torchrun's single OMP thread.

        When the unconditional branch passes the very same ``example_pair`` / ``query`` tensor objects as the
        conditional one (the notebook always does, run_prompt_diffusion.ipynb cell 5:27-33), the [B] tensor is handed
        to the model once instead of as a [2B] concatenation of two identical halves: ``ControlLDM`` tiles a hint batch
        that divides the latent batch, so the two hint encoders run on B images, not 2B."""
        dev = torch.device(self.model.device)
        to_dev = lambda t: t if t.device == dev else t.to(dev, non_blocking=True)
        tiles = bool(getattr(self.model, "tiles_hint_batch", False))
        out = {}
        for k in c:
            out[k] = []
            for i in range(len(c[k])):
                if tiles and k in self._SHARED_KEYS and uc[k][i] is c[k][i]:
                    out[k].append(to_dev(c[k][i]))
                else:
                    out[k].append(torch.cat([to_dev(uc[k][i]), to_dev(c[k][i])]))
        return out

    @torch.no_grad()
    def ddim_sampling(self, cond, shape, x_T=None, ddim_use_original_steps=False, callback=None,
                      timesteps=None, quantize_denoised=False, mask=None, x0=None, img_callback=None,
                      log_every_t=100, temperature=1., noise_dropout=0., score_corrector=None,
                      corrector_kwargs=None, unconditional_guidance_scale=1.,
                      unconditional_conditioning=None, dynamic_threshold=None, ucg_schedule=None):
        device = self.model.betas.device
        b = shape[0]
        img = torch.randn(shape, device=device) if x_T is None else x_T.to(device)

        if timesteps is None:
            timesteps = self.ddpm_num_timesteps if ddim_use_original_steps else self.ddim_timesteps
        elif not ddim_use_original_steps:
            subset_end = int(min(timesteps / self.ddim_timesteps.shape[0], 1) * self.ddim_timesteps.shape[0]) - 1
            timesteps = self.ddim_timesteps[:subset_end]

        intermediates = {'x_inter': [img], 'pred_x0': [img]}
        time_range = reversed(range(0, timesteps)) if ddim_use_original_steps else np.flip(timesteps)
        total_steps = timesteps if ddim_use_original_steps else timesteps.shape[0]

        # the [uncond, cond] batch does not depend on the step: build it once
        c_in = None
        if unconditional_conditioning is not None and isinstance(cond, dict):
            c_in = self._concat_conds(cond, unconditional_conditioning)

        for i, step in enumerate(time_range):
            index = total_steps - i - 1
            ts = torch.full((b,), int(step), device=device, dtype=torch.long)

            if mask is not None:
                assert x0 is not None
                img_orig = self.model.q_sample(x0, ts)
                img = img_orig * mask + (1. - mask) * img

            if ucg_schedule is not None:
                assert len(ucg_schedule) == len(time_range)
                unconditional_guidance_scale = ucg_schedule[i]

            img, pred_x0 = self.p_sample_ddim(img, cond, ts, index=index,
                                              use_original_steps=ddim_use_original_steps,
                                              quantize_denoised=quantize_denoised, temperature=temperature,
                                              noise_dropout=noise_dropout, score_corrector=score_corrector,
                                              corrector_kwargs=corrector_kwargs,
                                              unconditional_guidance_scale=unconditional_guidance_scale,
                                              unconditional_conditioning=unconditional_conditioning,
                                              dynamic_threshold=dynamic_threshold, _c_in=c_in)
            if callback:
                callback(i)
            if img_callback:
                img_callback(pred_x0, i)

            if index % log_every_t == 0 or index == total_steps - 1:
                intermediates['x_inter'].append(img)
                intermediates['pred_x0'].append(pred_x0)

        return img, intermediates

    @torch.no_grad()
    def p_sample_ddim(self, x, c, t, index, repeat_noise=False, use_original_steps=False,
                      quantize_denoised=False, temperature=1., noise_dropout=0., score_corrector=None,
                      corrector_kwargs=None, unconditional_guidance_scale=1., unconditional_conditioning=None,
                      dynamic_threshold=None, _c_in=None):
        b, device = x.shape[0], x.device
        model = self.model
        x = x.contiguous()
        if x.dtype != torch.float32:
            x = x.float()

        # CFG branch is taken whenever an unconditional conditioning is given — even at scale 1 (:188)
        guided = unconditional_conditioning is not None

        if dynamic_threshold is not None:
            raise NotImplementedError()

        plain = (model.parameterization == "eps" and score_corrector is None and not quantize_denoised
                 and noise_dropout == 0.)
        coef, host_row = self._coef_row(index, use_original_steps, unconditional_guidance_scale, temperature)
        if plain:
            c_in = None
            if guided:
                c_in = _c_in if _c_in is not None else self._concat_conds(c, unconditional_conditioning)
            # the reference draws the noise every step, also when sigma_t == 0 (:230): keep the RNG stream
            noise = _noise_like(x.shape, device, repeat_noise)
            noise = noise if host_row[2] != 0.0 else None
            return self._plain_step(x, t, c, c_in, guided, coef, noise)

        if guided:
            x_in, t_in = torch.cat([x] * 2), torch.cat([t] * 2)
            c_in = _c_in if _c_in is not None else self._concat_conds(c, unconditional_conditioning)
            out = model.apply_model(x_in, t_in, c_in)
            e_u, e_c = out[:b], out[b:]          # == .chunk(2): contiguous halves of a fresh tensor
        else:
            e_u, e_c = None, model.apply_model(x, t, c)
        noise = _noise_like(x.shape, device, repeat_noise)

        # ---- rarely used branches: the reference's unfused arithmetic (:193-233) ---------------------------
        model_output = e_c if e_u is None else e_u + unconditional_guidance_scale * (e_c - e_u)
        if model.parameterization == "v":
            e_t = model.predict_eps_from_z_and_v(x, t, model_output)
        else:
            e_t = model_output
        if score_corrector is not None:
            assert model.parameterization == "eps", 'not implemented'
            e_t = score_corrector.modify_score(model, e_t, x, t, c, **corrector_kwargs)
        full = lambda v: torch.full((b, 1, 1, 1), float(v), device=device)
        a_t, a_prev, sigma_t, s1m = (full(host_row[0]), full(host_row[1]), full(host_row[2]), full(host_row[3]))
        if model.parameterization != "v":
            pred_x0 = (x - s1m * e_t) / a_t.sqrt()
        else:
            pred_x0 = model.predict_start_from_z_and_v(x, t, model_output)
        if quantize_denoised:
            pred_x0, _, *_ = model.first_stage_model.quantize(pred_x0)
        dir_xt = (1. - a_prev - sigma_t ** 2).sqrt() * e_t
        noise = sigma_t * noise * temperature
        if noise_dropout > 0.:
            noise = torch.nn.functional.dropout(noise, p=noise_dropout)
        x_prev = a_prev.sqrt() * pred_x0 + dir_xt + noise
        return x_prev, pred_x0

    # ---- one denoising step on the fused path: apply_model + CFG combine + DDIM update --------------------
    def _eager_step(self, x, t, c, c_in, guided, coef, noise, x_prev, pred_x0):
        b = x.shape[0]
        if guided:
            cfg_call = getattr(self.model, "apply_model_cfg", None)
            if cfg_call is not None and x.dtype == torch.float32:
                out = cfg_call(x, t, c_in)          # same eps, no duplicated latent / timestep tensors
            else:
                out = self.model.apply_model(torch.cat([x] * 2), torch.cat([t] * 2), c_in)
            e_u, e_c = out[:b], out[b:]
        else:
            e_u, e_c = None, self.model.apply_model(x, t, c)
        ops.cfg_ddim_step(e_u, e_c.contiguous(), x, noise, coef, x_prev, pred_x0)

    def _plain_step(self, x, t, c, c_in, guided, coef, noise):
        """Runs the step either eagerly or — for this package's ControlLDM — as ONE replayed CUDA graph
        (~600 kernel launches per step otherwise cost more host time than a fast GPU step takes)."""
        if x.is_cuda and torch.cuda.current_device() != x.device.index:
            with torch.cuda.device(x.device):        # the C-ABI launches on the current device's stream
                return self._plain_step(x, t, c, c_in, guided, coef, noise)
        if step_graphs_enabled() and getattr(self.model, "supports_step_graph", False) and x.is_cuda:
            conds = c_in if guided else c
            key = _graph_key(self.model, x, t, conds, guided, noise is not None)
            g = self._graphs.get(key) if self._graphs is not None else None
            if g is None:
                if self._graphs is None or len(self._graphs) >= 4:
                    self._graphs = {}
                g = _StepGraph(self, x, t, c, c_in, guided, coef, noise)
                self._graphs[key] = g
            return g.replay(x, t, coef, noise, conds)
        x_prev, pred_x0 = torch.empty_like(x), torch.empty_like(x)
        self._eager_step(x, t, c, c_in, guided, coef, noise, x_prev, pred_x0)
        return x_prev, pred_x0

    @torch.no_grad()
    def encode(self, x0, c, t_enc, use_original_steps=False, return_intermediates=None,
               unconditional_guidance_scale=1.0, unconditional_conditioning=None, callback=None):
        """DDIM inversion (:237-281).  Like the reference, the CFG branch concatenates the conditionings with
        ``torch.cat`` and therefore only works for tensor conditionings, not ControlLDM's dicts."""
        num_reference_steps = self.ddpm_num_timesteps if use_original_steps else self.ddim_timesteps.shape[0]
        assert t_enc <= num_reference_steps
        num_steps = t_enc
        if use_original_steps:
            alphas_next = self.alphas_cumprod[:num_steps]
            alphas = self.alphas_cumprod_prev[:num_steps]
        else:
            alphas_next = self.ddim_alphas[:num_steps]
            alphas = torch.tensor(self.ddim_alphas_prev[:num_steps])
        x_next = x0
        intermediates, inter_steps = [], []
        for i in range(num_steps):
            t = torch.full((x0.shape[0],), i, device=self.model.device, dtype=torch.long)
            if unconditional_guidance_scale == 1.:
                noise_pred = self.model.apply_model(x_next, t, c)
            else:
                assert unconditional_conditioning is not None
                e_t_uncond, noise_pred = torch.chunk(
                    self.model.apply_model(torch.cat((x_next, x_next)), torch.cat((t, t)),
                                           torch.cat((unconditional_conditioning, c))), 2)
                noise_pred = e_t_uncond + unconditional_guidance_scale * (noise_pred - e_t_uncond)
            xt_weighted = (alphas_next[i] / alphas[i]).sqrt() * x_next
            weighted_noise_pred = alphas_next[i].sqrt() * (
                (1 / alphas_next[i] - 1).sqrt() - (1 / alphas[i] - 1).sqrt()) * noise_pred
            x_next = xt_weighted + weighted_noise_pred
            if return_intermediates and i % (num_steps // return_intermediates) == 0 and i < num_steps - 1:
                intermediates.append(x_next)
                inter_steps.append(i)
            elif return_intermediates and i >= num_steps - 2:
                intermediates.append(x_next)
                inter_steps.append(i)
            if callback:
                callback(i)
        out = {'x_encoded': x_next, 'intermediate_steps': inter_steps}
        if return_intermediates:
            out.update({'intermediates': intermediates})
        return x_next, out

    @torch.no_grad()
    def stochastic_encode(self, x0, t, use_original_steps=False, noise=None):
        """q(x_t | x_0) at DDIM index t (:284-297)."""
        if use_original_steps:
            sqrt_ac = self.sqrt_alphas_cumprod
            sqrt_1mac = self.sqrt_one_minus_alphas_cumprod
        else:
            sqrt_ac = torch.sqrt(self.ddim_alphas)
            sqrt_1mac = self.ddim_sqrt_one_minus_alphas
        if noise is None:
            noise = torch.randn_like(x0)
        gather = lambda a: a.to(x0.device).gather(-1, t).reshape(t.shape[0], *((1,) * (x0.dim() - 1)))
        return gather(sqrt_ac) * x0 + gather(sqrt_1mac) * noise

    @torch.no_grad()
    def decode(self, x_latent, cond, t_start, unconditional_guidance_scale=1.0,
               unconditional_conditioning=None, use_original_steps=False, callback=None):
        timesteps = np.arange(self.ddpm_num_timesteps) if use_original_steps else self.ddim_timesteps
        timesteps = timesteps[:t_start]
        time_range = np.flip(timesteps)
        total_steps = timesteps.shape[0]
        x_dec = x_latent
        # the [uncond, cond] batch does not depend on the step: build it once so the model's identity-keyed hint /
        # context-K/V caches hit on every step (rebuilding it per step recomputed both hint encoders 50 times)
        c_in = None
        if unconditional_conditioning is not None and isinstance(cond, dict):
            c_in = self._concat_conds(cond, unconditional_conditioning)
        for i, step in enumerate(time_range):
            index = total_steps - i - 1
            ts = torch.full((x_latent.shape[0],), int(step), device=x_latent.device, dtype=torch.long)
            x_dec, _ = self.p_sample_ddim(x_dec, cond, ts, index=index, use_original_steps=use_original_steps,
                                          unconditional_guidance_scale=unconditional_guidance_scale,
                                          unconditional_conditioning=unconditional_conditioning, _c_in=c_in)
            if callback:
                callback(i)
        return x_dec
