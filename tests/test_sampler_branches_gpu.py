"""The sampler branches the headline run does not take — mask / x0 blend, ``ucg_schedule``, ``decode``,
``stochastic_encode``, ``encode``, the ``timesteps=`` subset and ``use_original_steps`` — against outputs of the
UNMODIFIED reference ``DDIMSampler`` (tests/golden/sampler_branches_golden.npz, written by
tests/golden/make_golden_sampler_branches.py; reference cldm/ddim_hacked.py:122-318).

Gates: 3-4 step trajectories at 16x16 latents; fp32 mode rel-L2 <= 1e-3 (per-step eps parity is ~3e-6, the rest is
trajectory amplification), bf16 mode rel-L2 <= 5e-2 and cosine >= 0.999 (the north-star trajectory gate).
"""
import os

import numpy as np
import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda"
S, B, H, W = 4, 1, 128, 128
SHAPE = (4, H // 8, W // 8)
TOL = {"fp32": 1e-3, "bf16": 5e-2}


@pytest.fixture(scope="module")
def gold():
    path = os.path.join(os.path.dirname(__file__), "golden", "sampler_branches_golden.npz")
    return {k: v for k, v in np.load(path).items()}


@pytest.fixture(scope="module", autouse=True)
def _no_tf32():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.set_grad_enabled(False)
    yield
    torch.set_grad_enabled(True)


@pytest.fixture(scope="module")
def models(cfg, state_dict_cpu):
    from prompt_diffusion_b200 import ControlLDM
    out = {m: ControlLDM(cfg, mode=m, device=DEV).load_state_dict(state_dict_cpu) for m in ("fp32", "bf16")}
    torch.cuda.synchronize()
    return out


@pytest.fixture(scope="module")
def inputs(cfg):
    from prompt_diffusion_b200.synth import make_conds, synthetic_inputs
    inp = {k: v.to(DEV) for k, v in synthetic_inputs(cfg, B, H, W, seed=2).items()}
    cond, un = make_conds(inp)
    return inp, cond, un


def _check(tag, mode, got, want):
    want = torch.as_tensor(want)
    err = rel_l2(got.cpu(), want)
    cos = float(torch.nn.functional.cosine_similarity(got.cpu().flatten().double(), want.flatten().double(), dim=0))
    print(f"[parity] sampler branch {tag} {mode}: rel-L2 = {err:.3e}, cosine = {cos:.6f}")
    assert cos >= 0.999
    assert err <= TOL[mode]


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_mask_blend_vs_reference(models, inputs, gold, mode):
    """ddim_hacked.py:154-157: img = q_sample(x0, ts) * mask + (1 - mask) * img before every step; q_sample's noise
    is replayed from the golden file through q_sample's own ``noise=`` argument."""
    from prompt_diffusion_b200 import DDIMSampler
    model = models[mode]
    inp, cond, un = inputs
    q_noise = torch.as_tensor(gold["q_noise"]).to(DEV)
    seen = []
    own_q_sample = model.q_sample

    def q_sample_with_stored_noise(x_start, t, noise=None):
        seen.append(int(t[0]))
        return own_q_sample(x_start, t, noise=q_noise[len(seen) - 1])
    model.q_sample = q_sample_with_stored_noise
    try:
        z, inter = DDIMSampler(model).sample(
            S, B, SHAPE, cond, verbose=False, eta=0.0, x_T=inp["x_T"], mask=torch.as_tensor(gold["mask"]).to(DEV),
            x0=torch.as_tensor(gold["x0"]).to(DEV), unconditional_guidance_scale=5.0,
            unconditional_conditioning=un, log_every_t=1)
    finally:
        del model.q_sample
    assert seen == gold["mask_q_timesteps"].tolist()
    assert len(inter["pred_x0"]) == gold["mask_pred_x0"].shape[0]
    _check("mask", mode, z, gold["mask_final"])
    _check("mask pred_x0[1]", mode, inter["pred_x0"][1], gold["mask_pred_x0"][1])


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_ucg_schedule_vs_reference(models, inputs, gold, mode):
    """ddim_hacked.py:159-161: the guidance scale of step i is ucg_schedule[i] (here 7, 1, 3.5, 0 — the CFG batch is
    still evaluated at scale 1 and 0, :188)."""
    from prompt_diffusion_b200 import DDIMSampler
    inp, cond, un = inputs
    z, inter = DDIMSampler(models[mode]).sample(
        S, B, SHAPE, cond, verbose=False, eta=0.0, x_T=inp["x_T"], unconditional_guidance_scale=9.0,
        unconditional_conditioning=un, ucg_schedule=gold["ucg_schedule"].tolist(), log_every_t=1)
    _check("ucg_schedule", mode, z, gold["ucg_final"])
    for i, (a, b) in enumerate(zip(inter["x_inter"], gold["ucg_x_inter"])):
        assert rel_l2(a.cpu(), b) <= TOL[mode], i
    with pytest.raises(AssertionError):          # :160 assert len(ucg_schedule) == len(time_range)
        DDIMSampler(models[mode]).sample(S, B, SHAPE, cond, verbose=False, x_T=inp["x_T"],
                                         unconditional_conditioning=un, ucg_schedule=[1.0, 2.0])


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_stochastic_encode_then_decode_vs_reference(models, inputs, gold, mode):
    """ddim_hacked.py:283-318, the img2img pair: q(x_t | x_0) at DDIM index t_enc - 1, then t_enc guided steps."""
    from prompt_diffusion_b200 import DDIMSampler
    inp, cond, un = inputs
    smp = DDIMSampler(models[mode])
    smp.make_schedule(S, ddim_eta=0.0, verbose=False)
    x0 = torch.as_tensor(gold["x0"]).to(DEV)
    noise = torch.as_tensor(gold["enc_noise"]).to(DEV)
    t_enc = int(gold["decode_t_start"])
    z_enc = smp.stochastic_encode(x0, torch.tensor([t_enc - 1] * B, device=DEV), noise=noise)
    assert rel_l2(z_enc.cpu(), gold["stoch_encoded"]) <= 1e-6
    z_orig = smp.stochastic_encode(x0, torch.tensor([500] * B, device=DEV), use_original_steps=True, noise=noise)
    assert rel_l2(z_orig.cpu(), gold["stoch_encoded_orig_steps"]) <= 1e-6
    calls = []
    z = smp.decode(z_enc, cond, t_enc, unconditional_guidance_scale=5.0, unconditional_conditioning=un,
                   callback=calls.append)
    assert calls == list(range(t_enc))
    _check("decode", mode, z, gold["decode_final"])


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_encode_inversion_vs_reference(models, inputs, gold, mode):
    """ddim_hacked.py:236-281 at guidance scale 1 (the form that works with ControlLDM's dict conditionings; the CFG
    branch torch.cat's the conditionings and fails for dicts in the reference too — mirrored, not fixed)."""
    from prompt_diffusion_b200 import DDIMSampler
    inp, cond, un = inputs
    smp = DDIMSampler(models[mode])
    smp.make_schedule(S, ddim_eta=0.0, verbose=False)
    x0 = torch.as_tensor(gold["x0"]).to(DEV)
    z, info = smp.encode(x0, cond, 3, return_intermediates=3)
    assert info["intermediate_steps"] == gold["encode_intermediate_steps"].tolist()
    assert len(info["intermediates"]) == gold["encode_intermediates"].shape[0]
    _check("encode", mode, z, gold["encode_final"])
    _check("encode intermediates[0]", mode, info["intermediates"][0], gold["encode_intermediates"][0])
    with pytest.raises(TypeError):               # CFG inversion: torch.cat of two dicts (:260-262)
        smp.encode(x0, cond, 1, unconditional_guidance_scale=3.0, unconditional_conditioning=un)


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_timestep_subset_and_original_steps_vs_reference(models, inputs, gold, mode):
    """ddim_sampling(timesteps=3) keeps the first subset_end ddim timesteps (:138-140); p_sample_ddim with
    use_original_steps=True reads its coefficients from the 1000-step DDPM tables (:206-216) — and its sigmas from an
    attribute the reference model does not have (:209): AttributeError there, mirrored here."""
    from prompt_diffusion_b200 import DDIMSampler
    inp, cond, un = inputs
    smp = DDIMSampler(models[mode])
    smp.make_schedule(S, ddim_eta=0.0, verbose=False)
    z, inter = smp.ddim_sampling(cond, (B,) + SHAPE, x_T=inp["x_T"], timesteps=3, log_every_t=1,
                                 unconditional_guidance_scale=5.0, unconditional_conditioning=un)
    assert len(inter["x_inter"]) == int(gold["subset_n_inter"])
    _check("timesteps subset", mode, z, gold["subset_final"])
    assert int(gold["orig_steps_raises_attribute_error"]) == 1
    ts = torch.full((B,), 700, dtype=torch.long, device=DEV)
    with pytest.raises(AttributeError):          # no `ddim_sigmas_for_original_num_steps` on the model (:209)
        smp.p_sample_ddim(inp["x_T"], cond, ts, index=700, use_original_steps=True,
                          unconditional_guidance_scale=5.0, unconditional_conditioning=un)
    # a model that provides the table works: coefficients from the 1000-step DDPM tables, eta 0 => sigma 0
    model = models[mode]
    model.ddim_sigmas_for_original_num_steps = smp.ddim_sigmas_for_original_num_steps
    try:
        xp, px0 = smp.p_sample_ddim(inp["x_T"], cond, ts, index=700, use_original_steps=True,
                                    unconditional_guidance_scale=5.0, unconditional_conditioning=un)
    finally:
        del model.ddim_sigmas_for_original_num_steps
    a_t, a_prev = float(model.alphas_cumprod[700]), float(model.alphas_cumprod_prev[700])
    e_t = (inp["x_T"] - a_t ** 0.5 * px0) / (1.0 - a_t) ** 0.5           # invert pred_x0 (:218)
    want = a_prev ** 0.5 * px0 + (1.0 - a_prev) ** 0.5 * e_t               # :229-233 at sigma 0
    assert rel_l2(xp.cpu(), want.cpu()) <= 1e-5
