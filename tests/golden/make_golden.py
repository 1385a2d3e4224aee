"""Generate golden vectors by running the UNMODIFIED reference (``/root/reference``).

Run in the dev container only (the reference does not exist on the GPU box):

    python tests/golden/make_golden.py [--out tests/golden] [--skip-sample]

Recipe: stub ``omegaconf``, ``pytorch_lightning``, ``open_clip``;
build ``ControlLDM`` from ``models/cldm_v15.yaml`` with the first/cond stages
replaced; load the procedural checkpoint from ``prompt_diffusion_b200.synth``
(identical bits wherever it is regenerated); run the reference's own
``apply_model`` / ``DDIMSampler.sample`` on procedural inputs and store outputs.
Nothing from the reference is copied into the repo — only its numeric outputs.
"""
from __future__ import annotations

import argparse
import os
import sys
import time
import types

import numpy as np
import torch
import torch.nn as nn

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.environ.get("PD_REFERENCE_ROOT", "/root/reference")


def install_stubs():
    oc = types.ModuleType("omegaconf")
    lc = types.ModuleType("omegaconf.listconfig")

    class ListConfig(list):
        pass
    lc.ListConfig = ListConfig
    oc.listconfig = lc
    oc.ListConfig = ListConfig
    sys.modules["omegaconf"], sys.modules["omegaconf.listconfig"] = oc, lc

    pl = types.ModuleType("pytorch_lightning")

    class LightningModule(nn.Module):
        @property
        def device(self):
            return next(self.parameters()).device
    pl.LightningModule = LightningModule
    cb = types.ModuleType("pytorch_lightning.callbacks")
    cb.Callback = object
    ut = types.ModuleType("pytorch_lightning.utilities")
    rz = types.ModuleType("pytorch_lightning.utilities.rank_zero")
    rz.rank_zero_only = lambda f: f
    dist = types.ModuleType("pytorch_lightning.utilities.distributed")
    dist.rank_zero_only = rz.rank_zero_only
    ut.rank_zero, ut.distributed = rz, dist
    pl.callbacks, pl.utilities = cb, ut
    for n, m in (("pytorch_lightning", pl), ("pytorch_lightning.callbacks", cb),
                 ("pytorch_lightning.utilities", ut), ("pytorch_lightning.utilities.rank_zero", rz),
                 ("pytorch_lightning.utilities.distributed", dist)):
        sys.modules[n] = m
    sys.modules["open_clip"] = types.ModuleType("open_clip")


def build_reference_model():
    import yaml
    install_stubs()
    sys.path.insert(0, REF)
    sys.path.insert(0, REPO)
    from cldm.cldm import ControlLDM  # noqa: reference class
    with open(os.path.join(REF, "models", "cldm_v15.yaml")) as f:
        params = yaml.safe_load(f)["model"]["params"]
    params["first_stage_config"] = {"target": "torch.nn.Identity"}
    params["cond_stage_config"] = "__is_unconditional__"
    torch.manual_seed(0)
    model = ControlLDM(**params).eval()
    return model


def load_synthetic(model, cfg, seed=0):
    from prompt_diffusion_b200.synth import iter_synthetic_state_dict
    own = model.state_dict()
    n = 0
    for k, v in iter_synthetic_state_dict(cfg, seed):
        assert k in own, f"key {k} missing in reference state_dict"
        assert tuple(own[k].shape) == tuple(v.shape), (k, own[k].shape, v.shape)
        own[k].copy_(v)
        n += 1
    want = [k for k in own if k.startswith(("model.diffusion_model.", "control_model."))]
    assert n == len(want), (n, len(want))
    return n


def summary(t: torch.Tensor):
    t = t.double()
    return np.array([t.mean().item(), t.std().item(), t.abs().max().item()])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(REPO, "tests", "golden"))
    ap.add_argument("--skip-sample", action="store_true")
    args = ap.parse_args()
    torch.set_grad_enabled(False)

    t0 = time.time()
    model = build_reference_model()
    from cldm.ddim_hacked import DDIMSampler
    from ldm.modules.diffusionmodules.util import timestep_embedding
    from prompt_diffusion_b200.config import CLDM_V15 as cfg
    from prompt_diffusion_b200.synth import make_conds, synthetic_inputs
    n = load_synthetic(model, cfg, seed=0)
    print(f"reference built, {n} tensors loaded in {time.time() - t0:.1f}s")

    # CPU: register_buffer hard-codes .to('cuda') (cldm/ddim_hacked.py:17-21)
    DDIMSampler.register_buffer = lambda self, name, attr: setattr(self, name, attr)

    out = {}
    # ---- schedule known answers ------------------------------------------------
    for S, eta in ((20, 0.0), (50, 0.0), (50, 0.5)):
        smp = DDIMSampler(model)
        smp.make_schedule(S, ddim_eta=eta, verbose=False)
        tag = f"sched_S{S}_eta{eta}"
        out[tag + "_timesteps"] = np.asarray(smp.ddim_timesteps)
        out[tag + "_alphas"] = smp.ddim_alphas.numpy()
        out[tag + "_alphas_prev"] = np.asarray(smp.ddim_alphas_prev)
        out[tag + "_sigmas"] = np.asarray(smp.ddim_sigmas)
        out[tag + "_sqrt_one_minus"] = np.asarray(smp.ddim_sqrt_one_minus_alphas)
    out["betas"] = model.betas.numpy()
    out["alphas_cumprod"] = model.alphas_cumprod.numpy()
    out["alphas_cumprod_prev"] = model.alphas_cumprod_prev.numpy()
    out["temb_t"] = np.array([1, 21, 501, 981])
    out["temb"] = timestep_embedding(torch.tensor(out["temb_t"]), 320).numpy()

    # ---- apply_model cases ------------------------------------------------------
    # (name, batch, H, W, t values, control_scales, only_mid)
    cases = [
        ("cfg1", 1, 256, 256, [951, 951], None, False),            # BASELINE config 1 shape (B_eff 2)
        ("lat8", 2, 64, 64, [1, 501, 501, 981], None, False),      # 8x8 latent: multi-image tiles
        ("rect", 1, 192, 128, [301, 301], [0.5 + 0.1 * i for i in range(13)], False),
        ("midonly", 1, 128, 128, [701, 701], None, True),
    ]
    for name, b, H, W, tvals, scales, only_mid in cases:
        inp = synthetic_inputs(cfg, b, H, W, seed=2)
        cond, un = make_conds(inp)
        x_in = torch.cat([inp["x_T"]] * 2)
        c_in = {k: [torch.cat([un[k][0], cond[k][0]])] for k in cond}
        t = torch.tensor(tvals, dtype=torch.long)
        model.control_scales = [1.0] * 13 if scales is None else list(scales)
        model.only_mid_control = only_mid
        t1 = time.time()
        ctrl = model.control_model(x=x_in, timesteps=t, example_pair=c_in["example_pair"][0],
                                   query=c_in["query"][0], context=c_in["c_crossattn"][0])
        eps = model.apply_model(x_in, t, c_in)
        print(f"case {name}: apply_model {time.time() - t1:.1f}s eps std {eps.std():.4f}")
        out[f"{name}_eps"] = eps.numpy()
        out[f"{name}_t"] = np.asarray(tvals)
        out[f"{name}_ctrl_summary"] = np.stack([summary(c) for c in ctrl])
        out[f"{name}_ctrl0"] = ctrl[0][:, :8].numpy()       # first 8 channels of control 0
        out[f"{name}_ctrl12"] = ctrl[12][:, :8].numpy()     # first 8 channels of the mid control
    model.control_scales = [1.0] * 13
    model.only_mid_control = False

    # ---- full sampler, BASELINE config 1: 256^2, batch 1, 20 steps, CFG 9 -------
    if not args.skip_sample:
        inp = synthetic_inputs(cfg, 1, 256, 256, seed=2)
        cond, un = make_conds(inp)
        smp = DDIMSampler(model)
        t1 = time.time()
        samples, inter = smp.sample(20, 1, (4, 32, 32), cond, verbose=False, eta=0.0,
                                    x_T=inp["x_T"], unconditional_guidance_scale=9.0,
                                    unconditional_conditioning=un, log_every_t=5)
        dt = time.time() - t1
        print(f"sample cfg1: {dt:.1f}s")
        out["sample_cfg1_final"] = samples.numpy()
        out["sample_cfg1_x_inter"] = torch.stack(inter["x_inter"]).numpy()
        out["sample_cfg1_pred_x0"] = torch.stack(inter["pred_x0"]).numpy()
        out["sample_cfg1_seconds_ref_cpu"] = np.array([dt, torch.get_num_threads()])
        # eta > 0 path: 3 steps only (noise comes from torch's global CPU RNG, seeded)
        torch.manual_seed(1234)
        samples, inter = smp.sample(4, 1, (4, 16, 16), make_conds(synthetic_inputs(cfg, 1, 128, 128, 2))[0],
                                    verbose=False, eta=0.7, x_T=synthetic_inputs(cfg, 1, 128, 128, 2)["x_T"],
                                    unconditional_guidance_scale=1.0, log_every_t=1)
        out["sample_eta_final"] = samples.numpy()

    os.makedirs(args.out, exist_ok=True)
    path = os.path.join(args.out, "cldm_v15_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, f"{os.path.getsize(path) / 1e6:.2f} MB, total {time.time() - t0:.1f}s")


if __name__ == "__main__":
    main()
