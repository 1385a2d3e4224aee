"""First-stage decode timing at BASELINE config 2 (8 images, 64x64 latent -> 512x512), CUDA events, bf16 mode.
    python scripts/vae_bench.py [--batch 8] [--lat 64] [--iters 5]"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from prompt_diffusion_b200 import AutoencoderKLDecoder, _lib
from prompt_diffusion_b200.synth import synthetic_vae_state_dict
ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8); ap.add_argument("--lat", type=int, default=64)
ap.add_argument("--iters", type=int, default=5); ap.add_argument("--mode", default="bf16")
a = ap.parse_args()
torch.set_grad_enabled(False)
dev = "cuda"
dec = AutoencoderKLDecoder(a.mode, dev).load_state_dict(synthetic_vae_state_dict(0, device=dev), scale_factor=0.18215)
z = torch.randn(a.batch, 4, a.lat, a.lat, device=dev)
for _ in range(2): img = dec.decode(z, scaled=True)
torch.cuda.synchronize()
n0 = _lib.launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.iters): img = dec.decode(z, scaled=True)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.iters
# algorithmic FLOPs of Decoder.forward (2*MAC of every conv + the middle attention), from the weight shapes
P = a.lat * a.lat
fl = 0.0
def res(cin, cout, px): return 2.0 * px * (9 * cin * cout + 9 * cout * cout + (cin * cout if cin != cout else 0))
fl += 2.0 * P * (4 * 4 + 9 * 4 * 512) + 2 * res(512, 512, P) + 2.0 * P * 4 * 512 * 512 + 4.0 * P * P * 512
px, cin = P, 512
for lvl, cout in enumerate((512, 512, 256, 128)):
    for _ in range(3):
        fl += res(cin, cout, px); cin = cout
    if lvl != 3:
        px *= 4; fl += 2.0 * px * 9 * cin * cin
fl += 2.0 * px * 9 * cin * 3
fl *= a.batch
print(f"vae decode {a.mode}: batch {a.batch} latent {a.lat}^2 -> {8*a.lat}^2: {ms:.2f} ms ({ms/a.batch:.2f} ms/image), "
      f"{fl/1e12:.2f} TFLOP -> {fl/ms/1e9:.0f} TFLOP/s, {(_lib.launch_count()-n0)//a.iters} launches, "
      f"buffers {sum(b.numel()*b.element_size() for b in dec.bufs.values())/2**30:.2f} GiB, |img| mean {float(img.abs().mean()):.4f}")
