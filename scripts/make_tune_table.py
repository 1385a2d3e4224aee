"""Regenerate prompt-diffusion_b200/csrc/tune_table.inc (the committed per-shape launch-variant table of the tcgen05
GEMM engine).

On a B200 (gpurun):   PD_B200_AUTOTUNE=1 python scripts/make_tune_table.py --dump gpurun_out/tune_dump.inc
    runs one apply_model of every workload shape the tests / bench use (and a VAE decode) with the opt-in timing
    autotune on, then writes the chosen variants with pd_tune_dump.
Here (no GPU):        python scripts/make_tune_table.py --install gpurun_out/tune_dump.inc
    copies the dump into csrc/tune_table.inc (sorted, with a header); rebuild afterwards.
"""
import argparse
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
TABLE = os.path.join(REPO, "prompt-diffusion_b200", "csrc", "tune_table.inc")

ap = argparse.ArgumentParser()
ap.add_argument("--dump")
ap.add_argument("--install")
ap.add_argument("--keep-default", action="store_true", help="also keep rows equal to the cost-model default")
a = ap.parse_args()

if a.install:
    rows = sorted(set(l.strip() for l in open(a.install) if l.strip().startswith("{{")),
                  key=lambda l: [int(x) for x in l.replace("{", "").replace("}", "").split(",") if x.strip()])
    with open(TABLE, "w") as f:
        f.write("// {{M, N, K, ksize, stride, has_x2, epilogue flags}, {cta_group, stream_k, BN override}} — generated on a B200 by\n"
                "// scripts/make_tune_table.py (PD_B200_AUTOTUNE=1 timing of every layer shape of configs 1/2/4, the test slices\n"
                "// and the first-stage decoder); committed so that the launch variant never depends on timing noise.\n")
        for r in rows:
            f.write(r + "\n")
    print(f"installed {len(rows)} rows into {TABLE}")
    sys.exit(0)

import torch  # noqa: E402
from prompt_diffusion_b200 import CLDM_V15 as cfg, ControlLDM, _lib  # noqa: E402
from prompt_diffusion_b200.synth import make_conds, synthetic_inputs, synthetic_state_dict  # noqa: E402

assert os.environ.get("PD_B200_AUTOTUNE") == "1", "run with PD_B200_AUTOTUNE=1"
torch.set_grad_enabled(False)
dev = "cuda:0"
sd = synthetic_state_dict(cfg, seed=0)
model = ControlLDM(cfg, mode="bf16", device=dev).load_state_dict(sd)
for b, size in ((8, 512), (1, 512), (2, 512), (16, 768), (1, 768), (1, 256), (2, 64), (1, 128), (2, 128), (1, 64)):
    inp = {k: v.to(dev) for k, v in synthetic_inputs(cfg, b, size, size, seed=2).items()}
    cond, un = make_conds(inp)
    x_in = torch.cat([inp["x_T"]] * 2)
    c_in = {k: [torch.cat([un[k][0], cond[k][0]])] for k in cond}
    t = torch.full((2 * b,), 501, device=dev, dtype=torch.long)
    model.apply_model(x_in, t, c_in)
    if size in (128, 64, 256):                       # unguided shapes of the tests (B_eff = b)
        model.apply_model(inp["x_T"], t[:b], cond)
    torch.cuda.synchronize()
    print("tuned", b, size, flush=True)
try:
    from prompt_diffusion_b200.autoencoder import AutoencoderKLDecoder
    from prompt_diffusion_b200.synth import synthetic_vae_state_dict
    vae = AutoencoderKLDecoder("bf16", dev).load_state_dict(synthetic_vae_state_dict(seed=0))
    for b, hw in ((8, 64), (1, 32), (2, 16)):
        vae.decode(torch.randn(b, 4, hw, hw, device=dev))
        torch.cuda.synchronize()
        print("tuned vae", b, hw, flush=True)
except Exception as e:                                   # the decoder is a neighbour of the path, not the path
    print("vae tuning skipped:", e)
_lib.check(_lib.lib.pd_tune_dump(a.dump.encode()), "pd_tune_dump")
print("dumped", a.dump)
