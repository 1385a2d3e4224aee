"""Same-box A/B of the graph-replayed denoise step at config 2 with the three-group attention kernel on / off."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from prompt_diffusion_b200 import CLDM_V15 as cfg, ControlLDM, DDIMSampler, _lib
from prompt_diffusion_b200.synth import make_conds, synthetic_inputs, synthetic_state_dict
torch.set_grad_enabled(False)
dev = "cuda:0"
B, size = int(os.environ.get("AB_BATCH", 8)), int(os.environ.get("AB_SIZE", 512))
model = ControlLDM(cfg, mode="bf16", device=dev).load_state_dict(synthetic_state_dict(cfg, seed=0))
inp = {k: v.to(dev) for k, v in synthetic_inputs(cfg, B, size, size, seed=2).items()}
cond, un = make_conds(inp)
x = inp["x_T"]
ts = torch.full((B,), 501, device=dev, dtype=torch.long)
def measure(tag, reps=12):
    smp = DDIMSampler(model)
    smp.make_schedule(50, ddim_eta=0.0, verbose=False)
    c_in = smp._concat_conds(cond, un)
    run = lambda: smp.p_sample_ddim(x, cond, ts, index=25, unconditional_guidance_scale=9.0, unconditional_conditioning=un, _c_in=c_in)
    for _ in range(3): run()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): run()
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / reps)
    print(f"{tag:40s} {best:8.3f} ms / denoise step", flush=True)
for rnd in range(2):
    for on in (0, 2):
        _lib.lib.pd_debug_attention_tc3(on)
        measure(f"round {rnd}: three-group attention = {on}")
