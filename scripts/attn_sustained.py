"""Attention engines 3 / 6 at B16 h8 N4096 d40 under SUSTAINED load (power-capped clocks), alone and alternating with a
large GEMM as in the model."""
import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from prompt_diffusion_b200 import ops
dev = "cuda"
B, h, N, d = 16, 8, 4096, 40
C = h * d
qkv = torch.randn(B * N, 3 * C, device=dev).to(torch.bfloat16)
out = torch.empty(B * N, C, device=dev, dtype=torch.bfloat16)
x = torch.randn(B * N, C, device=dev).to(torch.bfloat16)
w = (torch.randn(3 * C, C, device=dev) * 0.05).to(torch.bfloat16)
w2 = (torch.randn(C, C, device=dev) * 0.05).to(torch.bfloat16)
bias = torch.zeros(3 * C, device=dev); bias2 = torch.zeros(C, device=dev)
y = torch.empty(B * N, C, device=dev, dtype=torch.bfloat16)
def clocks():
    return subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
def run(eng, with_gemm, iters):
    def body():
        if with_gemm: ops.linear(x, w, qkv, bias=bias)
        ops.attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], out, B, h, N, N, d, engine=eng)
        if with_gemm: ops.linear(out, w2, y, bias=bias2, res=x)
    for _ in range(5): body()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(20): body()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters // 20): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (iters // 20 * 20) * 1e3
for with_gemm in (0, 1):
    for eng in (3, 6, 3, 6):
        us = run(eng, with_gemm, 600)
        print(f"engine {eng} {'qkv GEMM + attention + out GEMM' if with_gemm else 'attention alone':34s}: {us:8.1f} us per iteration   [{clocks()}]", flush=True)
