"""One denoise step (apply_model at B_eff = 2*batch + fused CFG/DDIM update) of BASELINE config 2 for ncu:
everything before torch.cuda.profiler.start() is warm-up (weights, caches).  Usage:
    python scripts/profile_step.py [--batch 8] [--size 512] [--reps 1]
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv ... python scripts/profile_step.py
"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from prompt_diffusion_b200 import CLDM_V15 as cfg, ControlLDM, DDIMSampler, _lib
from prompt_diffusion_b200.synth import make_conds, synthetic_inputs, synthetic_state_dict

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--size", type=int, default=512)
ap.add_argument("--reps", type=int, default=1)
ap.add_argument("--mode", default="bf16")
ap.add_argument("--graph", type=int, default=0, help="1: time the CUDA-graph replay of the step instead of eager launches")
a = ap.parse_args()
torch.set_grad_enabled(False)
if os.environ.get("PD_TIME_TABLE") == "0":          # A/B: direct evaluation of the timestep embedding every step
    from prompt_diffusion_b200.cldm import cldm as _M
    _M.TIME_EMBED_TABLE = False
from prompt_diffusion_b200.cldm.ddim_hacked import set_step_graphs
set_step_graphs(bool(a.graph))
dev = "cuda:0"
sd = synthetic_state_dict(cfg, 0, device=dev)
model = ControlLDM(cfg, mode=a.mode, device=dev).load_state_dict(sd)
del sd
inp = synthetic_inputs(cfg, a.batch, a.size, a.size, seed=2, device=dev)
cond, un = make_conds(inp)
smp = DDIMSampler(model)
smp.make_schedule(50, ddim_eta=0.0, verbose=False)
c_in = smp._concat_conds(cond, un)
x = inp["x_T"]
ts = torch.full((a.batch,), 501, device=dev, dtype=torch.long)
for _ in range(3):
    smp.p_sample_ddim(x, cond, ts, index=25, unconditional_guidance_scale=9.0, unconditional_conditioning=un, _c_in=c_in)
torch.cuda.synchronize()
n0 = _lib.launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.profiler.start()
t0 = time.perf_counter()
e0.record()
for _ in range(a.reps):
    smp.p_sample_ddim(x, cond, ts, index=25, unconditional_guidance_scale=9.0, unconditional_conditioning=un, _c_in=c_in)
e1.record()
host_ms = (time.perf_counter() - t0) * 1e3
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(f"denoise step: {e0.elapsed_time(e1) / a.reps:.3f} ms (GPU events), host issue {host_ms / a.reps:.3f} ms, "
      f"{(_lib.launch_count() - n0) // a.reps} pd_b200 launches/step, pool {model.pool.nbytes() / 2**30:.2f} GiB")
if os.environ.get("PD_DUMP"):
    set_step_graphs(False)
    _lib.lib.pd_prof_enable(1)
    smp.p_sample_ddim(x, cond, ts, index=25, unconditional_guidance_scale=9.0, unconditional_conditioning=un, _c_in=c_in)
    torch.cuda.synchronize()
    _lib.lib.pd_prof_dump(os.environ["PD_DUMP"].encode())
    _lib.lib.pd_prof_enable(0)
    print("dumped", os.environ["PD_DUMP"])
