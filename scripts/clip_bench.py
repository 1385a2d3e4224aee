"""CLIP text-tower timing (16 prompts x 77 tokens = cond + uncond of BASELINE config 2), CUDA events, bf16 mode."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from prompt_diffusion_b200 import FrozenCLIPTextEncoder, _lib
from prompt_diffusion_b200.synth import synthetic_clip_state_dict, synthetic_tokens
torch.set_grad_enabled(False)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
enc = FrozenCLIPTextEncoder("bf16", "cuda").load_state_dict(synthetic_clip_state_dict(0, device="cuda"))
tok = synthetic_tokens(B).cuda()
for _ in range(3): z = enc.encode(tok)
torch.cuda.synchronize()
n0 = _lib.launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): z = enc.encode(tok)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
fl = B * 77 * 12 * 2.0 * (4 * 768 * 768 + 2 * 768 * 3072) + B * 12 * 12 * 4.0 * 77 * 77 * 64
print(f"clip text encode bf16: {B} prompts x 77 tokens: {ms:.3f} ms, {fl/1e9:.1f} GFLOP -> {fl/ms/1e9:.1f} TFLOP/s, "
      f"{(_lib.launch_count()-n0)//10} launches (launch-bound: M = {B*77} rows)")
