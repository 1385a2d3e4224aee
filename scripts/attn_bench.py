"""CUDA-event micro-benchmark of pd_attention engines (2 = mma.sync, 3 = tcgen05) at the path's shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from prompt_diffusion_b200 import ops
dev = "cuda"
def run(B, heads, Nq, Nk, d, engine, iters=10):
    C = heads * d
    q = torch.randn(B * Nq, C, device=dev).to(torch.bfloat16)
    kv = torch.randn(B * Nk, 2 * C, device=dev).to(torch.bfloat16)
    out = torch.empty(B * Nq, C, device=dev, dtype=torch.bfloat16)
    for _ in range(2): ops.attention(q, kv[:, :C], kv[:, C:], out, B, heads, Nq, Nk, d, engine=engine)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): ops.attention(q, kv[:, :C], kv[:, C:], out, B, heads, Nq, Nk, d, engine=engine)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / iters * 1e3
    return us, 4.0 * B * heads * Nq * Nk * d / us / 1e6
print("   B  h    Nq    Nk    d eng |      us  TFLOP/s")
for (B, h, Nq, Nk, d) in [(16, 8, 4096, 4096, 40), (16, 8, 4096, 77, 40), (16, 8, 1024, 1024, 80), (16, 8, 1024, 77, 80), (32, 8, 9216, 9216, 40)]:
    for eng in (2, 3, 5, 6):
        if eng == 3 and d > 128: continue
        if eng == 5 and (d > 64 or Nk < 512): continue
        if eng == 6 and (d > 40 or Nk < 256): continue
        if Nq > 8000 and eng == 2: continue
        try:
            us, tf = run(B, h, Nq, Nk, d, eng, iters=5 if Nq > 8000 else 10)
            print("%4d %2d %5d %5d %4d %3d | %7.1f %8.0f" % (B, h, Nq, Nk, d, eng, us, tf))
        except Exception as e:
            print(B, h, Nq, Nk, d, eng, "FAILED", str(e)[:100])
