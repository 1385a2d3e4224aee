"""d = 160 attention (16x16 / 8x8 levels): tcgen05 kernel (engine 3) against the mma.sync kernel (engine 2), CUDA-graph timed."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from prompt_diffusion_b200 import ops
dev = "cuda"
def run(B, heads, Nq, Nk, d, engine, iters=20):
    C = heads * d
    q = torch.randn(B * Nq, C, device=dev).to(torch.bfloat16)
    kv = torch.randn(B * Nk, 2 * C, device=dev).to(torch.bfloat16)
    out = torch.empty(B * Nq, C, device=dev, dtype=torch.bfloat16)
    a = (q, kv[:, :C], kv[:, C:], out, B, heads, Nq, Nk, d)
    for _ in range(3): ops.attention(*a, engine=engine)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters): ops.attention(*a, engine=engine)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3
print("   B  h    Nq    Nk    d | mma.sync us | tcgen05 us")
for s_ in [(16, 8, 256, 256, 160), (16, 8, 256, 77, 160), (16, 8, 64, 64, 160), (16, 8, 64, 77, 160)]:
    print("%4d %2d %5d %5d %4d | %11.1f | %10.1f" % (*s_, run(*s_, 2), run(*s_, 3)))
print("77-key cross-attention:    | short-key (mma.sync) us | tcgen05 streaming us | tcgen05 persistent short-key us")
for s_ in [(16, 8, 4096, 77, 40), (16, 8, 1024, 77, 80), (32, 8, 9216, 77, 40), (32, 8, 2304, 77, 80), (2, 8, 4096, 77, 40)]:
    print("%4d %2d %5d %5d %4d | %11.1f | %10.1f | %10.1f" % (*s_, run(*s_, 4), run(*s_, 3), run(*s_, 7)))
print("self-attention d <= 64:      | streaming (engine 3) us | three groups (6) us | persistent (8) us")
for s_ in [(16, 8, 4096, 4096, 40), (32, 8, 9216, 9216, 40), (16, 8, 1024, 1024, 40)]:
    print("%4d %2d %5d %5d %4d | %11.1f | %10.1f | %10.1f" % (*s_, run(*s_, 3, iters=5), run(*s_, 6, iters=5), run(*s_, 8, iters=5)))
