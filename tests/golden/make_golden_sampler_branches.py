"""Golden vectors for the sampler branches the headline run does not take, produced by the UNMODIFIED reference
``DDIMSampler`` (``/root/reference/cldm/ddim_hacked.py``) on the CPU:

* ``ddim_sampling`` with ``mask`` / ``x0`` (:154-157) — the only random input, ``q_sample``'s noise, is handed to the
  reference's own ``q_sample(x_start, t, noise=)`` argument from a stored list, so the GPU test can replay it;
* ``ucg_schedule`` (:159-161);
* ``decode`` (:299-318) with CFG, ``stochastic_encode`` (:283-297) with given noise,
  ``encode`` (:236-281) at guidance scale 1 with ``return_intermediates``;
* ``ddim_sampling(timesteps=...)`` subset (:138-140); ``p_sample_ddim(use_original_steps=True)`` (:206-216) raises in the
  reference (see below) — recorded as behaviour.

Run in the dev container only (the reference does not travel to the GPU box):

    python tests/golden/make_golden_sampler_branches.py

All cases: 128x128 image = 16x16 latent, batch 1, 4 DDIM steps, procedural checkpoint / inputs of
``prompt_diffusion_b200.synth`` (seed 0 / seed 2) — the same bits wherever they are regenerated.
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import REPO, build_reference_model, load_synthetic  # noqa: E402

S, B, H, W = 4, 1, 128, 128
SHAPE = (4, H // 8, W // 8)


def main():
    torch.set_grad_enabled(False)
    t0 = time.time()
    model = build_reference_model()
    from cldm.ddim_hacked import DDIMSampler
    from prompt_diffusion_b200.config import CLDM_V15 as cfg
    from prompt_diffusion_b200.synth import make_conds, synthetic_inputs
    load_synthetic(model, cfg, seed=0)
    DDIMSampler.register_buffer = lambda self, name, attr: setattr(self, name, attr)   # CPU (see make_golden.py)

    inp = synthetic_inputs(cfg, B, H, W, seed=2)
    cond, un = make_conds(inp)
    g = torch.Generator().manual_seed(77)
    x0 = torch.randn((B,) + SHAPE, generator=g)
    mask = (torch.rand((B, 1) + SHAPE[1:], generator=g) > 0.5).float()
    q_noise = torch.randn((S, B) + SHAPE, generator=g)
    enc_noise = torch.randn((B,) + SHAPE, generator=g)
    out = {"x0": x0.numpy(), "mask": mask.numpy(), "q_noise": q_noise.numpy(), "enc_noise": enc_noise.numpy()}

    smp = DDIMSampler(model)

    # ---- mask / x0 blend ------------------------------------------------------------------------------------
    ref_q_sample = model.q_sample
    calls = []

    def q_sample_with_stored_noise(x_start, t, noise=None):
        calls.append(int(t[0]))
        return ref_q_sample(x_start, t, noise=q_noise[len(calls) - 1])
    model.q_sample = q_sample_with_stored_noise
    z, inter = smp.sample(S, B, SHAPE, cond, verbose=False, eta=0.0, x_T=inp["x_T"], mask=mask, x0=x0,
                          unconditional_guidance_scale=5.0, unconditional_conditioning=un, log_every_t=1)
    model.q_sample = ref_q_sample
    assert len(calls) == S
    out["mask_final"] = z.numpy()
    out["mask_pred_x0"] = torch.stack(inter["pred_x0"]).numpy()
    out["mask_q_timesteps"] = np.asarray(calls)

    # ---- ucg_schedule ---------------------------------------------------------------------------------------
    ucg = [7.0, 1.0, 3.5, 0.0]
    z, inter = smp.sample(S, B, SHAPE, cond, verbose=False, eta=0.0, x_T=inp["x_T"],
                          unconditional_guidance_scale=9.0, unconditional_conditioning=un, ucg_schedule=ucg,
                          log_every_t=1)
    out["ucg_schedule"] = np.asarray(ucg)
    out["ucg_final"] = z.numpy()
    out["ucg_x_inter"] = torch.stack(inter["x_inter"]).numpy()

    # ---- stochastic_encode -> decode (the img2img pair) -----------------------------------------------------
    smp.make_schedule(S, ddim_eta=0.0, verbose=False)
    t_enc = 3
    z_enc = smp.stochastic_encode(x0, torch.tensor([t_enc - 1] * B), noise=enc_noise)
    out["stoch_encoded"] = z_enc.numpy()
    out["stoch_encoded_orig_steps"] = smp.stochastic_encode(x0, torch.tensor([500] * B), use_original_steps=True,
                                                            noise=enc_noise).numpy()
    dec_calls = []
    z_dec = smp.decode(z_enc, cond, t_enc, unconditional_guidance_scale=5.0, unconditional_conditioning=un,
                       callback=dec_calls.append)
    out["decode_final"] = z_dec.numpy()
    out["decode_t_start"] = np.asarray(t_enc)
    assert dec_calls == list(range(t_enc))

    # ---- encode (DDIM inversion), guidance scale 1 (the only form that works with dict conditionings) -------
    z_inv, info = smp.encode(x0, cond, 3, return_intermediates=3)
    out["encode_final"] = z_inv.numpy()
    out["encode_intermediates"] = torch.stack(info["intermediates"]).numpy()
    out["encode_intermediate_steps"] = np.asarray(info["intermediate_steps"])

    # ---- ddim_sampling(timesteps=...) subset ----------------------------------------------------------------
    z, inter = smp.ddim_sampling(cond, (B,) + SHAPE, x_T=inp["x_T"], timesteps=3, log_every_t=1,
                                 unconditional_guidance_scale=5.0, unconditional_conditioning=un)
    out["subset_final"] = z.numpy()
    out["subset_n_inter"] = np.asarray(len(inter["x_inter"]))

    # ---- p_sample_ddim(use_original_steps=True): the reference reads `self.model.ddim_sigmas_for_original_num_steps`
    # (:209) — an attribute of the SAMPLER, not of the model — and raises AttributeError; recorded as behaviour
    ts = torch.full((B,), 700, dtype=torch.long)
    try:
        smp.p_sample_ddim(inp["x_T"], cond, ts, index=700, use_original_steps=True,
                          unconditional_guidance_scale=5.0, unconditional_conditioning=un)
        out["orig_steps_raises_attribute_error"] = np.asarray(0)
    except AttributeError as e:
        print("use_original_steps=True:", e)
        out["orig_steps_raises_attribute_error"] = np.asarray(1)

    path = os.path.join(REPO, "tests", "golden", "sampler_branches_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, f"{os.path.getsize(path) / 1e3:.1f} kB, total {time.time() - t0:.1f}s")
    for k, v in out.items():
        print(f"  {k}: {v.shape}")


if __name__ == "__main__":
    main()
