"""System test of the widened path: token ids -> CLIP text tower -> 50% of a DDIM+CFG sampling loop -> first-stage
decode, every stage on the CUDA kernels through the reference-facing methods of ONE ControlLDM loaded from ONE
checkpoint dict (control_model.* + model.diffusion_model.* + cond_stage_model.* + first_stage_model.*), against the
three oracles chained the same way (notebook cell 5: get_learned_conditioning -> DDIMSampler.sample ->
decode_first_stage)."""
import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("mode,tol", [("fp32", 2e-4), ("bf16", 6e-2)])
def test_tokens_to_image(cfg, state_dict_cpu, vae_state_dict_cpu, clip_state_dict_cpu, mode, tol):
    from oracle import cldm_oracle as O, clip_oracle as C, vae_oracle as V
    from prompt_diffusion_b200 import ControlLDM, DDIMSampler
    from prompt_diffusion_b200.synth import synthetic_inputs, synthetic_tokens
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.set_grad_enabled(False)
    full = {**state_dict_cpu, **vae_state_dict_cpu, **clip_state_dict_cpu}
    model = ControlLDM(cfg, mode=mode, device=DEV).load_state_dict(full)
    assert model.cond_stage_model is not None and model.first_stage_model is not None
    S, steps_run, scale = 4, 4, 7.5
    inp = synthetic_inputs(cfg, 1, 64, 64, seed=4, device=DEV)              # 8x8 latent, 64x64 hints / image
    tok = synthetic_tokens(2, seed=6).to(DEV)                                 # [prompt, negative prompt]
    # ---- product path -------------------------------------------------------------------------------------------
    c = model.get_learned_conditioning(tok[:1])
    uc = model.get_learned_conditioning(tok[1:])
    cond = {"c_crossattn": [c], "example_pair": [inp["example_pair"]], "query": [inp["query"]]}
    un = {"c_crossattn": [uc], "example_pair": [inp["example_pair"]], "query": [inp["query"]]}
    z, _ = DDIMSampler(model).sample(S, 1, (4, 8, 8), cond, verbose=False, eta=0.0, x_T=inp["x_T"],
                                     unconditional_guidance_scale=scale, unconditional_conditioning=un)
    img = model.decode_first_stage(z)
    # ---- oracle chain (fp32, on the GPU) ----------------------------------------------------------------------------
    sd = {k: v.to(DEV) for k, v in full.items()}
    rc, ruc = C.clip_text_forward(sd, tok[:1]), C.clip_text_forward(sd, tok[1:])
    rcond = {"c_crossattn": [rc], "example_pair": [inp["example_pair"]], "query": [inp["query"]]}
    run = {"c_crossattn": [ruc], "example_pair": [inp["example_pair"]], "query": [inp["query"]]}
    rz, _ = O.ddim_sample(sd, cfg, S, (1, 4, 8, 8), rcond, eta=0.0, x_T=inp["x_T"], unconditional_guidance_scale=scale,
                          unconditional_conditioning=run, max_steps=steps_run)
    rimg = V.decode_first_stage(sd, rz, cfg.scale_factor)
    e_c, e_z, e_i = rel_l2(c.cpu(), rc.cpu()), rel_l2(z.cpu(), rz.cpu()), rel_l2(img.cpu(), rimg.cpu())
    print(f"[parity] tokens->image {mode}: conditioning {e_c:.3e}, latents after {S} DDIM steps {e_z:.3e}, image {e_i:.3e}")
    assert img.shape == (1, 3, 64, 64) and bool(torch.isfinite(img).all())
    assert e_c <= tol and e_z <= tol and e_i <= tol


def test_diffusers_style_loop_equals_ldm_sampler(cfg, state_dict_cpu):
    """PromptDiffusionPipeline's denoising loop (pipeline_prompt_diffusion.py:1195-1290: controlnet -> unet with
    additional residuals -> CFG -> DDIMScheduler.step) over the diffusers-signature shims must land on the latents of
    DDIMSampler.sample (cldm/ddim_hacked.py, parity pinned to the reference) for the same weights, noise and
    conditioning: same schedule, same arithmetic, different call structure.  fp32 mode, 5 steps (a divisor of 1000: for
    other step counts ldm's range(0, 1000, 1000 // S) yields an extra timestep), CFG 7.5; plus the
    guess_mode and control_guidance_end branches run and stay finite."""
    from prompt_diffusion_b200 import ControlLDM, DDIMSampler, PromptDiffusionControlNetModel, PromptDiffusionPipeline
    from prompt_diffusion_b200.synth import make_conds, synthetic_inputs
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.set_grad_enabled(False)
    model = ControlLDM(cfg, mode="fp32", device=DEV).load_state_dict(state_dict_cpu)
    inp = synthetic_inputs(cfg, 2, 64, 64, seed=8, device=DEV)
    cond, un = make_conds(inp)
    S, scale = 5, 7.5
    ref, _ = DDIMSampler(model).sample(S, 2, (4, 8, 8), cond, verbose=False, eta=0.0, x_T=inp["x_T"],
                                       unconditional_guidance_scale=scale, unconditional_conditioning=un)
    cn = PromptDiffusionControlNetModel(cfg, mode="fp32", device=DEV).load_state_dict(state_dict_cpu)
    pipe = PromptDiffusionPipeline(model.model.diffusion_model, cn)
    seen = []
    z = pipe(prompt_embeds=inp["c_crossattn"], negative_prompt_embeds=inp["uc_crossattn"], image=inp["query"],
             image_pair=inp["example_pair"], num_inference_steps=S, guidance_scale=scale, latents=inp["x_T"],
             output_type="latent", callback=lambda i, t, l: seen.append(int(t)))
    err = rel_l2(z.cpu(), ref.cpu())
    print(f"[parity] diffusers-style loop vs DDIMSampler.sample (fp32, {S} steps): rel-L2 = {err:.3e}")
    assert seen == [int(t) for t in pipe.scheduler.timesteps] and seen[0] == 1 + (S - 1) * (1000 // S) and seen[-1] == 1
    assert err <= 1e-4
    z2 = pipe(prompt_embeds=inp["c_crossattn"], negative_prompt_embeds=inp["uc_crossattn"], image=inp["query"],
              image_pair=inp["example_pair"], num_inference_steps=3, guidance_scale=scale, latents=inp["x_T"],
              guess_mode=True, control_guidance_end=0.5)
    assert z2.shape == ref.shape and bool(torch.isfinite(z2).all()) and rel_l2(z2.cpu(), ref.cpu()) > 1e-3
