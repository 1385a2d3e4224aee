"""Per-sample() fixed costs at BASELINE config 2: prepare_conditioning (hint encoders + context K/V) with fresh tensors."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from prompt_diffusion_b200 import CLDM_V15 as cfg, ControlLDM, DDIMSampler
from prompt_diffusion_b200.synth import make_conds, synthetic_inputs, synthetic_state_dict
torch.set_grad_enabled(False)
dev = "cuda:0"
model = ControlLDM(cfg, mode="bf16", device=dev).load_state_dict(synthetic_state_dict(cfg, 0, device=dev))
inp = synthetic_inputs(cfg, 8, 512, 512, seed=2, device=dev)
cond, un = make_conds(inp)
smp = DDIMSampler(model)
smp.make_schedule(50, ddim_eta=0.0, verbose=False)
for rep in range(4):
    c_in = smp._concat_conds({k: [v[0].clone()] for k, v in cond.items()}, {k: [v[0].clone()] for k, v in un.items()})
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    model.prepare_conditioning(c_in)
    e1.record(); torch.cuda.synchronize()
    print(f"prepare_conditioning (B_eff 16, 512^2 hints, fresh tensors): {e0.elapsed_time(e1):.2f} ms GPU, {(time.perf_counter()-t0)*1e3:.2f} ms wall")
