"""Self-attention timing of the tcgen05 engine at the two hot shapes (CUDA events, warm): python scripts/attn_quick.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from prompt_diffusion_b200 import ops
dev = "cuda"
def run(B, heads, N, d, iters=10):
    C = heads * d
    qkv = torch.randn(B * N, 3 * C, device=dev).to(torch.bfloat16)
    out = torch.empty(B * N, C, device=dev, dtype=torch.bfloat16)
    a = (qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], out, B, heads, N, N, d)
    for _ in range(3): ops.attention(*a, engine=3)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): ops.attention(*a, engine=3)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / iters * 1e3
    # accuracy against fp32 softmax on a slice (batch 0, head 0)
    q = qkv[:N, :d].float(); k = qkv[:N, C:C + d].float(); v = qkv[:N, 2 * C:2 * C + d].float()
    ref = torch.softmax(q @ k.T * d ** -0.5, -1) @ v
    err = float((out[:N, :d].float() - ref).norm() / ref.norm())
    return us, err
def run_cross(B, heads, N, d, engine, iters=20):
    C = heads * d
    q = torch.randn(B * N, C, device=dev).to(torch.bfloat16)
    kv = torch.randn(B * 77, 2 * C, device=dev).to(torch.bfloat16)
    out = torch.empty(B * N, C, device=dev, dtype=torch.bfloat16)
    a = (q, kv[:, :C], kv[:, C:], out, B, heads, N, 77, d)
    for _ in range(3): ops.attention(*a, engine=engine)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): ops.attention(*a, engine=engine)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3
if os.environ.get("CROSS"):
    for (N, d) in ((4096, 40), (1024, 80)):
        print("cross Nk=77 N=%d d=%d: " % (N, d) + "  ".join("eng%d %.1f us" % (e, run_cross(16, 8, N, d, e)) for e in (2, 3, 4)))
tag = os.environ.get("PD_B200_LIB", "default").split("/")[-1]
r = [run(16, 8, 4096, 40), run(16, 8, 1024, 80)]
print("%-22s d40 N4096: %7.1f us (err %.2e) | d80 N1024: %6.1f us (err %.2e)" % (tag, r[0][0], r[0][1], r[1][0], r[1][1]))
