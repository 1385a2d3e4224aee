"""Regenerate prompt-diffusion_b200/csrc/tune_table.inc (the committed per-shape launch-variant table of the tcgen05
GEMM engine).

On a B200 (gpurun):   PD_B200_AUTOTUNE=1 [PD_B200_RETUNE=1] python scripts/make_tune_table.py --dump gpurun_out/tune_dump.inc
    (PD_B200_RETUNE=1 ignores the committed rows, so every shape is timed again)
    records the GEMM shapes of one apply_model per workload (pd_prof), then launches every unique shape once on
    synthetic operands with the opt-in timing autotune on and writes the winners with pd_tune_dump.
Here (no GPU):        python scripts/make_tune_table.py --install gpurun_out/tune_dump.inc
    copies the dump into csrc/tune_table.inc (sorted, with a header); rebuild afterwards.
"""
import argparse
import csv
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
TABLE = os.path.join(REPO, "prompt-diffusion_b200", "csrc", "tune_table.inc")

ap = argparse.ArgumentParser()
ap.add_argument("--dump")
ap.add_argument("--install")
ap.add_argument("--workloads", default="8x512,16x768,1x512,1x768,1x256")
a = ap.parse_args()

if a.install:
    rows = sorted(set(l.strip() for l in open(a.install) if l.strip().startswith("{{")),
                  key=lambda l: [int(x) for x in l.replace("{", "").replace("}", "").split(",") if x.strip()])
    with open(TABLE, "w") as f:
        f.write("// {{M, N, K, ksize, stride, 0, 8 if GEGLU epilogue}, {cta_group, stream_k, BN override, resident B}} — generated on a B200 by\n"
                "// scripts/make_tune_table.py (PD_B200_AUTOTUNE=1 timing of every GEMM shape of the listed workloads); committed so\n"
                "// that the launch variant, and with it the fp32 summation order, never depends on timing noise.\n")
        for r in rows:
            f.write(r + "\n")
    print(f"installed {len(rows)} rows into {TABLE}")
    sys.exit(0)

import torch  # noqa: E402
from prompt_diffusion_b200 import CLDM_V15 as cfg, ControlLDM, _lib, ops  # noqa: E402
from prompt_diffusion_b200._lib import PD_ACT_GEGLU  # noqa: E402
from prompt_diffusion_b200.synth import make_conds, synthetic_inputs, synthetic_state_dict  # noqa: E402

assert os.environ.get("PD_B200_AUTOTUNE") == "1", "run with PD_B200_AUTOTUNE=1"
torch.set_grad_enabled(False)
dev = "cuda:0"
model = ControlLDM(cfg, mode="bf16", device=dev).load_state_dict(synthetic_state_dict(cfg, seed=0))
shapes = {}
tmp = a.dump + ".shapes.csv"
for wl in a.workloads.split(","):
    b, size = (int(v) for v in wl.split("x"))
    inp = {k: v.to(dev) for k, v in synthetic_inputs(cfg, b, size, size, seed=2).items()}
    cond, un = make_conds(inp)
    x_in = torch.cat([inp["x_T"]] * 2)
    c_in = {k: [torch.cat([un[k][0], cond[k][0]])] for k in cond}
    t = torch.full((2 * b,), 501, device=dev, dtype=torch.long)
    model.apply_model(x_in, t, c_in)                    # warm: caches, buffers (profiled calls never tune)
    _lib.lib.pd_prof_enable(1)
    model.apply_model(x_in, t, c_in)
    torch.cuda.synchronize()
    _lib.lib.pd_prof_dump(tmp.encode())
    _lib.lib.pd_prof_enable(0)
    for r in csv.DictReader(open(tmp)):
        key = tuple(int(r[k]) for k in ("B", "H", "W", "C", "C2", "N", "ksize", "stride", "act"))
        shapes[key] = shapes.get(key, 0) + 1
    print("recorded", wl, len(shapes), "unique shapes so far", flush=True)
del model
torch.cuda.empty_cache()
g = torch.Generator(device=dev).manual_seed(0)
for (B, H, W, C, C2, N, ks, st, act) in sorted(shapes):
    Ho, Wo = (H, W) if ks == 2 else ((H + 2 * (ks // 2) - ks) // st + 1, (W + 2 * (ks // 2) - ks) // st + 1)
    x = torch.randn(B * H * W, C, device=dev, generator=g).to(torch.bfloat16)
    w = (torch.randn(N, ks * ks * C + C2, device=dev, generator=g) * 0.02).to(torch.bfloat16)
    x2 = torch.randn(B * Ho * Wo, C2, device=dev, generator=g).to(torch.bfloat16) if C2 else None
    out = torch.empty(B * Ho * Wo, N // 2 if act == PD_ACT_GEGLU else N, device=dev, dtype=torch.bfloat16)
    bias = torch.zeros(N, device=dev)
    kw = dict(pad=(1, 1)) if ks == 2 else {}
    ops.conv2d(x, w, out, B, H, W, ksize=ks, stride=st, bias=bias, x2=x2, act=act if act == PD_ACT_GEGLU else 0, **kw)
    torch.cuda.synchronize()
_lib.check(_lib.lib.pd_tune_dump(a.dump.encode()), "pd_tune_dump")
print("tuned", len(shapes), "shapes; dumped", a.dump)
