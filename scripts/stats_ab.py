"""In-process A/B of the step time (CUDA-graph replay, CUDA events) at BASELINE config 2 under the feature switches:
GroupNorm statistics from GEMM epilogues (cldm.GN_STATS_MIN_K), LayerNorm partials (cldm.LN_PARTS), the four-group
attention kernel (pd_debug_attention_tc4).  One model load; a fresh sampler (= fresh step graph) per setting."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from prompt_diffusion_b200 import CLDM_V15 as cfg, ControlLDM, DDIMSampler, _lib
from prompt_diffusion_b200.cldm import cldm as M
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
    for _ in range(3):
        out = run()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            out = run()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / reps)
    print(f"{tag:58s} {best:8.3f} ms / denoise step", flush=True)
    return out[0]


ref = None
if os.environ.get("AB_QUICK"):
    # (GroupNorm records from launches with K >= min_k, LayerNorm partials), each setting twice, interleaved
    for rnd in range(2):
        for min_k, ln_parts in ((10 ** 9, False), (10 ** 9, True), (1024, False)):
            M.GN_STATS_MIN_K, M.LN_PARTS = min_k, ln_parts
            measure(f"round {rnd} gn_stats_min_k={min_k:<10d} ln_parts={ln_parts}")
    sys.exit(0)
ref = None
for tc4 in (0, 1):
    _lib.lib.pd_debug_attention_tc4(tc4)
    for min_k in (10 ** 9, 1024, 0):
        for ln_parts in (False, True):
            M.GN_STATS_MIN_K, M.LN_PARTS = min_k, ln_parts
            if tc4 == 1 and not (min_k == 1024 or (min_k == 0 and ln_parts)):
                continue
            z = measure(f"attn4={tc4} gn_stats_min_k={min_k:<10d} ln_parts={ln_parts}")
            if ref is None:
                ref = z
            else:
                print(f"    x_prev rel-L2 vs first setting: {float((z - ref).norm() / ref.norm()):.3e}", flush=True)
