import os, sys
sys.path.insert(0, "/root/repo")
import torch
from prompt_diffusion_b200 import CLDM_V15 as cfg, ControlLDM
from prompt_diffusion_b200.synth import make_conds, synthetic_inputs, synthetic_state_dict
torch.set_grad_enabled(False)
dev = "cuda"
rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm())
sd = synthetic_state_dict(cfg, 0, device=dev)
for mode, (bsz, size) in (("fp32", (2, 256)), ("bf16", (2, 256)), ("bf16", (8, 512))):
    m = ControlLDM(cfg, mode=mode, device=dev).load_state_dict(sd)
    inp = synthetic_inputs(cfg, bsz, size, size, seed=2, device=dev)
    cond, un = make_conds(inp)
    x = torch.cat([inp["x_T"]] * 2)
    c_in = {k: [torch.cat([un[k][0], cond[k][0]])] for k in cond}
    B = x.shape[0]
    t = torch.full((B,), 421, dtype=torch.long, device=dev)
    eps = m.apply_model(x, t, c_in)
    g = torch.Generator(device="cpu").manual_seed(12)
    perm = torch.randperm(B, generator=g).to(dev)
    c_perm = {k: [v[0][perm].contiguous()] for k, v in c_in.items()}
    eps_p = m.apply_model(x[perm].contiguous(), t, c_perm)
    xp = x * (1 + 1e-6 * torch.randn_like(x))
    eps_n = m.apply_model(xp, t, c_in)
    # per-sample permutation error
    per = [rel(eps_p[i], eps[perm[i]]) for i in range(B)]
    print(f"{mode} B_eff {B} {size}^2: permutation rel-L2 {rel(eps_p, eps[perm]):.3e} (per sample max {max(per):.3e} min {min(per):.3e}); "
          f"1e-6 input noise -> {rel(eps_n, eps):.3e}")
    del m
