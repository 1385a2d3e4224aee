"""Pin the oracle (oracle/cldm_oracle.py) against outputs of the reference itself
(tests/golden/cldm_v15_golden.npz, made by tests/golden/make_golden.py)."""
import numpy as np
import torch

from conftest import rel_l2
from oracle import cldm_oracle as O
from prompt_diffusion_b200.synth import make_conds, synthetic_inputs


def test_schedule_known_answers(golden, cfg):
    sched = O.register_schedule(cfg.timesteps, cfg.linear_start, cfg.linear_end)
    for k in ("betas", "alphas_cumprod", "alphas_cumprod_prev"):
        assert np.array_equal(sched[k].numpy(), golden[k]), k
    for S, eta in ((20, 0.0), (50, 0.0), (50, 0.5)):
        d = O.make_schedule(sched, S, eta)
        tag = f"sched_S{S}_eta{eta}"
        assert np.array_equal(d["ddim_timesteps"], golden[tag + "_timesteps"])
        assert np.array_equal(np.asarray(d["ddim_alphas"]), golden[tag + "_alphas"])
        assert np.array_equal(np.asarray(d["ddim_alphas_prev"]), golden[tag + "_alphas_prev"])
        assert np.array_equal(np.asarray(d["ddim_sigmas"]), golden[tag + "_sigmas"])
        assert np.array_equal(np.asarray(d["ddim_sqrt_one_minus_alphas"]), golden[tag + "_sqrt_one_minus"])
    # SURVEY.md appendix C scalars
    assert float(sched["alphas_cumprod"][0]) == 0.9991499781608582
    d50 = O.make_schedule(sched, 50, 0.0)
    assert float(d50["ddim_alphas"][49]) == 0.00577550008893013
    assert list(d50["ddim_timesteps"][:3]) == [1, 21, 41] and d50["ddim_timesteps"][-1] == 981


def test_timestep_embedding(golden):
    emb = O.timestep_embedding(torch.tensor(golden["temb_t"]), 320).numpy()
    assert np.array_equal(emb, golden["temb"])


CASES = {
    "cfg1": (1, 256, 256, None, False),
    "lat8": (2, 64, 64, None, False),
    "rect": (1, 192, 128, [0.5 + 0.1 * i for i in range(13)], False),
    "midonly": (1, 128, 128, None, True),
}


def _cfg_inputs(cfg, b, H, W):
    inp = synthetic_inputs(cfg, b, H, W, seed=2)
    cond, un = make_conds(inp)
    x_in = torch.cat([inp["x_T"]] * 2)
    c_in = {k: [torch.cat([un[k][0], cond[k][0]])] for k in cond}
    return x_in, c_in


def test_apply_model_matches_reference(golden, cfg, state_dict_cpu):
    torch.set_grad_enabled(False)
    for name, (b, H, W, scales, only_mid) in CASES.items():
        x_in, c_in = _cfg_inputs(cfg, b, H, W)
        t = torch.tensor(golden[f"{name}_t"], dtype=torch.long)
        ctrl = O.control_net_forward(state_dict_cpu, cfg, x_in, t, c_in["example_pair"][0],
                                     c_in["query"][0], c_in["c_crossattn"][0])
        assert len(ctrl) == 13
        assert rel_l2(ctrl[0][:, :8], golden[f"{name}_ctrl0"]) < 1e-5
        assert rel_l2(ctrl[12][:, :8], golden[f"{name}_ctrl12"]) < 1e-5
        eps = O.apply_model(state_dict_cpu, cfg, x_in, t, c_in, scales, only_mid)
        err = rel_l2(eps, golden[f"{name}_eps"])
        assert err < 1e-5, (name, err)


def test_sampler_matches_reference(golden, cfg, state_dict_cpu):
    """BASELINE config 1 (256^2, batch 1, 20 steps, CFG 9) — first 3 steps against the
    reference's stored intermediates is enough to pin the loop (the full 20-step
    trajectory is checked on the GPU box against the CUDA path)."""
    torch.set_grad_enabled(False)
    inp = synthetic_inputs(cfg, 1, 256, 256, seed=2)
    cond, un = make_conds(inp)
    # log_every_t=5 in the golden run: x_inter = [x_T, idx19, idx15, idx10, idx5, idx0]
    img, inter = O.ddim_sample(state_dict_cpu, cfg, 20, (1, 4, 32, 32), cond, eta=0.0, x_T=inp["x_T"],
                               unconditional_guidance_scale=9.0, unconditional_conditioning=un,
                               log_every_t=5, max_steps=1)
    assert rel_l2(inter["x_inter"][1], golden["sample_cfg1_x_inter"][1]) < 1e-5
    assert rel_l2(inter["pred_x0"][1], golden["sample_cfg1_pred_x0"][1]) < 1e-5


def test_vae_decode_oracle_vs_reference_golden(golden_vae, vae_state_dict_cpu):
    """oracle/vae_oracle.py (decode_first_stage -> AutoencoderKL.decode -> Decoder.forward) against the reference's own
    Decoder output on the same procedural checkpoint; also pins the parameter census (49.5 M)."""
    from oracle import vae_oracle as V
    assert int(golden_vae["n_params"]) == sum(v.numel() for v in vae_state_dict_cpu.values())
    torch.set_grad_enabled(False)
    for name in ("z16", "z8x24"):
        z = torch.tensor(golden_vae[name + "_z"])
        img = V.decode_first_stage(vae_state_dict_cpu, z, float(golden_vae["scale_factor"]))
        assert img.shape == golden_vae[name + "_img"].shape
        assert rel_l2(img, golden_vae[name + "_img"]) < 1e-5, name


def test_clip_text_oracle_vs_transformers_golden(golden_clip, clip_state_dict_cpu):
    """oracle/clip_oracle.py against transformers.CLIPTextModel on the same procedural checkpoint and token ids; pins
    the parameter census (123 060 480) and the tokenizer-shaped synthetic ids."""
    from oracle import clip_oracle as C
    from prompt_diffusion_b200.synth import synthetic_tokens
    assert int(golden_clip["n_params"]) == sum(v.numel() for v in clip_state_dict_cpu.values()) == 123_060_480
    tokens = torch.tensor(golden_clip["tokens"])
    assert torch.equal(tokens, synthetic_tokens(3, seed=2))
    with torch.no_grad():
        z = C.clip_text_forward(clip_state_dict_cpu, tokens)
    assert z.shape == golden_clip["z"].shape and rel_l2(z, golden_clip["z"]) < 1e-5
